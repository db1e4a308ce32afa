#!/bin/bash
# round 2, call e3 (2 GPUs): sharded parity tests that need two ranks, then the sharded bench line at N = 2
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/e3_smi.txt 2>&1
timeout -s KILL 420 python -m pytest tests/test_dist_gpu.py -x -q -k "2-" > gpurun_out/e3_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e3_tests.log; tail -6 gpurun_out/e3_tests.log
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e3_bench_2gpu.json 2> gpurun_out/e3_bench_2gpu.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/e3_bench_2gpu.json; tail -5 gpurun_out/e3_bench_2gpu.err
