#!/bin/bash
# round 2, call e8: the default bench line with the final library (what the driver runs at round end)
set -u
mkdir -p gpurun_out
timeout -s KILL 1200 python bench.py > gpurun_out/e8_bench.json 2> gpurun_out/e8_bench.err
echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/e8_bench.json; tail -5 gpurun_out/e8_bench.err
