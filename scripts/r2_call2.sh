#!/bin/bash
# round 2, GPU call 2: line-worker copy-engine kernels (correctness, timing), then the new bench line
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_tma_gpu.py -x -q > gpurun_out/c2_tma_tests.log 2>&1
rc=$?; echo "tma tests rc=$rc" >> gpurun_out/c2_tma_tests.log
tail -5 gpurun_out/c2_tma_tests.log
if [ $rc -ne 0 ]; then
    for sel in pass real kspace strided split; do
        timeout -s KILL 200 python -m pytest tests/test_tma_gpu.py -q -k "$sel" > gpurun_out/c2_tma_tests_$sel.log 2>&1
        echo "rc=$?" >> gpurun_out/c2_tma_tests_$sel.log; tail -3 gpurun_out/c2_tma_tests_$sel.log
    done
fi
timeout -s KILL 600 python scripts/tune_tma.py 1024 512 > gpurun_out/c2_tune_tma.jsonl 2> gpurun_out/c2_tune_tma.err
cat gpurun_out/c2_tune_tma.jsonl; tail -3 gpurun_out/c2_tune_tma.err
[ $rc -ne 0 ] && export GOPF_TMA=0
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/c2_bench.json; tail -5 gpurun_out/c2_bench.err
