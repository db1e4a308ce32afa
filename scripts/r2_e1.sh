#!/bin/bash
# round 2, call e1: fp32 Box-Muller k-noise, gradient-line loads, computed line frequencies, spectrum-prefetch timing
set -u
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_tma_gpu.py tests/test_blocked_gpu.py tests/test_zz_jit_gpu.py -x -q > gpurun_out/e1_tests_a.log 2>&1
echo "rc=$?" >> gpurun_out/e1_tests_a.log; tail -4 gpurun_out/e1_tests_a.log
timeout -s KILL 500 python -m pytest tests/test_step_gpu.py -x -q -k "square or pfc or noise or 100_steps or tensorial or spectral or three_fields" > gpurun_out/e1_tests_b.log 2>&1
echo "rc=$?" >> gpurun_out/e1_tests_b.log; tail -4 gpurun_out/e1_tests_b.log
TUNE_VARIANTS="tma all, blocked s=7" timeout -s KILL 400 python scripts/tune_tma.py 1024 > gpurun_out/e1_tune.jsonl 2> gpurun_out/e1_tune.err
cat gpurun_out/e1_tune.jsonl; tail -3 gpurun_out/e1_tune.err
for w in pfc ch_sqgrad; do
  timeout -s KILL 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/e1_bench_$w.json 2> gpurun_out/e1_bench_$w.err
  echo "bench $w rc=$?"; python scripts/show_bench.py gpurun_out/e1_bench_$w.json; tail -3 gpurun_out/e1_bench_$w.err
done
timeout -s KILL 300 python bench.py --grid 256 --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-workloads --no-cfg2 > gpurun_out/e1_bench_256.json 2> gpurun_out/e1_bench_256.err
echo "bench 256 rc=$?"; python scripts/show_bench.py gpurun_out/e1_bench_256.json; tail -3 gpurun_out/e1_bench_256.err
