#!/bin/bash
# round 2, GPU call 1: copy-engine kernels (correctness, then timing), then the whole device suite with the
# round-2 defaults (NVRTC specialisation on, k-space noise compiled in)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout -s KILL 420 python -m pytest tests/test_tma_gpu.py -x -q > gpurun_out/c1_tma_tests.log 2>&1
rc=$?; echo "tma tests rc=$rc" >> gpurun_out/c1_tma_tests.log
if [ $rc -ne 0 ]; then
    # which of the three kernels is at fault?
    for k in PASS REAL KSPACE; do
        sel="pass"; [ $k = REAL ] && sel="real"; [ $k = KSPACE ] && sel="kspace"
        timeout -s KILL 200 python -m pytest tests/test_tma_gpu.py -q -k "$sel" > gpurun_out/c1_tma_tests_$sel.log 2>&1
        echo "rc=$?" >> gpurun_out/c1_tma_tests_$sel.log
    done
    export GOPF_TMA=0
fi
timeout -s KILL 900 python scripts/tune_tma.py 1024 512 > gpurun_out/c1_tune_tma.jsonl 2> gpurun_out/c1_tune_tma.err
timeout -s KILL 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_tma_gpu.py > gpurun_out/c1_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/c1_suite.log
tail -5 gpurun_out/c1_tma_tests.log; cat gpurun_out/c1_tune_tma.jsonl; tail -15 gpurun_out/c1_suite.log
