#!/bin/bash
# round 2, call e10: the whole device suite with the paired real-space kernel in, then the default bench line
set -u
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/e10_suite.log 2>&1
echo "rc=$?" >> gpurun_out/e10_suite.log; tail -18 gpurun_out/e10_suite.log
timeout -s KILL 900 python bench.py > gpurun_out/e10_bench.json 2> gpurun_out/e10_bench.err
echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/e10_bench.json; tail -5 gpurun_out/e10_bench.err
