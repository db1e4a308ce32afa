#!/bin/bash
# round 2, call e17: paired real-space kernel for the tabulated form and from 512-cell lines on
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_tma_gpu.py -x -q -k "real_pairs" > gpurun_out/e17_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e17_tests.log; tail -8 gpurun_out/e17_tests.log
timeout -s KILL 200 python bench.py --workload pfc --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/e17_bench_pfc.json 2> gpurun_out/e17_bench_pfc.err
echo "bench pfc rc=$?"; python scripts/show_bench.py gpurun_out/e17_bench_pfc.json; tail -2 gpurun_out/e17_bench_pfc.err
