#!/bin/bash
# round 2, call e11 (2 GPUs): the sharded bench line with the final library (paired real-space kernel in the pipelined
# real-space side, 64 SMs for the peer-storing pass), and the two-rank parity tests at 256^3
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_dist_gpu.py -x -q -k "2-64 or 2-256" > gpurun_out/e11_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e11_tests.log; tail -4 gpurun_out/e11_tests.log
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e11_bench_2gpu.json 2> gpurun_out/e11_bench_2gpu.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/e11_bench_2gpu.json").read().strip().splitlines()[-1])
ph=d["roofline"]["phases"]
print(round(d["ms_per_step"],3), "ms/step", round(d["value"]/1e9,2), "G/s parity", d.get("parity",{}).get("rel_l2"), d["detail"].get("comm_ctas"), {k: round(v["avg_ms"]*v["launches"]/d["steps"],2) for k,v in ph.items()})
PY
tail -3 gpurun_out/e11_bench_2gpu.err
