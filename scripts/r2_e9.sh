#!/bin/bash
# round 2, call e9: real fields, two lines per complex transform in the copy-engine real-space kernel
set -u
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_tma_gpu.py -x -q -k "real_pairs" > gpurun_out/e9_tests_a.log 2>&1
echo "rc=$?" >> gpurun_out/e9_tests_a.log; tail -15 gpurun_out/e9_tests_a.log
timeout -s KILL 600 python -m pytest tests/test_tma_gpu.py tests/test_blocked_gpu.py -x -q -k "not real_pairs and not 1024_cubed" > gpurun_out/e9_tests_b.log 2>&1
echo "rc=$?" >> gpurun_out/e9_tests_b.log; tail -5 gpurun_out/e9_tests_b.log
TUNE_VARIANTS="final" timeout -s KILL 400 python scripts/tune_tma.py 1024 512 > gpurun_out/e9_tune.jsonl 2> gpurun_out/e9_tune.err
cat gpurun_out/e9_tune.jsonl; tail -3 gpurun_out/e9_tune.err
