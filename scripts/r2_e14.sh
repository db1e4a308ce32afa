#!/bin/bash
# round 2, call e14: ImplicitEuler after the 32-bit Freq change (tests + the configs[1]-as-worded record), smoke()
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_implicit_euler_gpu.py -x -q > gpurun_out/e14_ie_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e14_ie_tests.log; tail -3 gpurun_out/e14_ie_tests.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/e14_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/e14_smoke.log
timeout -s KILL 300 python bench.py --workload ch_sqgrad --stepper implicit_euler --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/e14_bench_ie.json 2> gpurun_out/e14_bench_ie.err
echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/e14_bench_ie.json; tail -2 gpurun_out/e14_bench_ie.err
