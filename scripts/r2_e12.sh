#!/bin/bash
# round 2, call e12 (8 GPUs): the sharded bench line at N = 8 with the final library
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/e12_bench_8gpu.json 2> gpurun_out/e12_bench_8gpu.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/e12_bench_8gpu.json").read().strip().splitlines()[-1])
ph=d["roofline"]["phases"]
print(round(d["ms_per_step"],3), "ms/step", round(d["value"]/1e9,2), "G/s parity", d.get("parity",{}).get("rel_l2"), d["detail"].get("comm_ctas"), {k: round(v["avg_ms"]*v["launches"]/d["steps"],2) for k,v in ph.items()})
PY
tail -3 gpurun_out/e12_bench_8gpu.err
