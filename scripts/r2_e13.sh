#!/bin/bash
# round 2, call e13 (8 GPUs): SMs given to the peer-storing forward pass at N = 8 now that the compute side is faster
set -u
mkdir -p gpurun_out
for cc in 64 76; do
  timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 296$cc bench.py --gpus 8 --steps 20 --warmup 5 --no-parity --comm-ctas $cc > gpurun_out/e13_bench_8gpu_cc$cc.json 2> gpurun_out/e13_bench_8gpu_cc$cc.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/e13_bench_8gpu_cc$cc.json").read().strip().splitlines()[-1])
ph=d["roofline"]["phases"]
print("comm_ctas $cc:", round(d["ms_per_step"],3), "ms/step", {k: round(v["avg_ms"]*v["launches"]/d["steps"],2) for k,v in ph.items()})
PY
done
