#!/bin/bash
# round 2, call e5: the whole device suite, then the default bench line (what the driver runs at round end)
set -u
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/e5_suite.log 2>&1
echo "rc=$?" >> gpurun_out/e5_suite.log; tail -25 gpurun_out/e5_suite.log
timeout -s KILL 900 python bench.py > gpurun_out/e5_bench.json 2> gpurun_out/e5_bench.err
echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/e5_bench.json; tail -5 gpurun_out/e5_bench.err
