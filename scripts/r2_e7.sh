#!/bin/bash
# round 2, call e7 (2 GPUs): SMs given to the peer-storing forward pass at N = 2 (the forward exchange was exposed for
# 2.6 ms per step with 48)
set -u
mkdir -p gpurun_out
for cc in 64 80 96; do
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$cc bench.py --gpus 2 --steps 10 --warmup 3 --no-parity --comm-ctas $cc > gpurun_out/e7_bench_2gpu_cc$cc.json 2> gpurun_out/e7_bench_2gpu_cc$cc.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/e7_bench_2gpu_cc$cc.json").read().strip().splitlines()[-1])
ph=d["roofline"]["phases"]
print("comm_ctas $cc:", round(d["ms_per_step"],3), "ms/step", {k: round(v["avg_ms"]*v["launches"]/d["steps"],2) for k,v in ph.items()})
PY
done
