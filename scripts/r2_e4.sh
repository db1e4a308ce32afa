#!/bin/bash
# round 2, call e4: cfg 5 with the L2 prefetch of the tabulated factors; first device run of the in-pass
# specialisation (registered functions compiled into their first forward pass) and cfg 4 both ways; ImplicitEuler on
# cfg 2 + SquaredGradient (BASELINE.json configs[1] as worded)
set -u
mkdir -p gpurun_out
timeout -s KILL 200 python bench.py --workload pfc --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/e4_bench_pfc.json 2> gpurun_out/e4_bench_pfc.err
echo "bench pfc rc=$?"; python scripts/show_bench.py gpurun_out/e4_bench_pfc.json; tail -3 gpurun_out/e4_bench_pfc.err
GOPF_TEST_INPASS=1 timeout -s KILL 300 python -m pytest tests/test_zz_jit_gpu.py -x -q -k "compiled_into" > gpurun_out/e4_inpass_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e4_inpass_tests.log; tail -15 gpurun_out/e4_inpass_tests.log
for v in 0 1; do
  GOPF_JIT_INPASS=$v timeout -s KILL 300 python bench.py --workload precipitate --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/e4_bench_precip_inpass$v.json 2> gpurun_out/e4_bench_precip_inpass$v.err
  echo "bench precipitate inpass=$v rc=$?"; python scripts/show_bench.py gpurun_out/e4_bench_precip_inpass$v.json; tail -3 gpurun_out/e4_bench_precip_inpass$v.err
done
GOPF_JIT_INPASS=1 timeout -s KILL 200 python bench.py --workload ch_sqgrad --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/e4_bench_sqgrad_inpass.json 2> gpurun_out/e4_bench_sqgrad_inpass.err
echo "bench sqgrad inpass rc=$?"; python scripts/show_bench.py gpurun_out/e4_bench_sqgrad_inpass.json; tail -3 gpurun_out/e4_bench_sqgrad_inpass.err
timeout -s KILL 400 python bench.py --workload ch_sqgrad --stepper implicit_euler --steps 3 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/e4_bench_sqgrad_ie.json 2> gpurun_out/e4_bench_sqgrad_ie.err
echo "bench sqgrad implicit euler rc=$?"; python scripts/show_bench.py gpurun_out/e4_bench_sqgrad_ie.json; tail -3 gpurun_out/e4_bench_sqgrad_ie.err
