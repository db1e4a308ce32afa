#!/bin/bash
# round 2, call e15: the whole device suite on the final tree
set -u
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/e15_suite.log 2>&1
echo "rc=$?" >> gpurun_out/e15_suite.log; tail -14 gpurun_out/e15_suite.log
