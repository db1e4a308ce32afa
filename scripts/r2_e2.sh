#!/bin/bash
# round 2, call e2: k-noise restructure check, same-box A/B of the 256^3 step against the round-1 library, ncu of
# the tabulated k-space kernel (cfg 5) and of the 256-cell k-space kernel
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_zz_jit_gpu.py tests/test_step_gpu.py -x -q -k "noise or pfc" > gpurun_out/e2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e2_tests.log; tail -4 gpurun_out/e2_tests.log
for rep in 1 2; do
  (cd scripts/ab_r1 && timeout -s KILL 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-scaling-base > ../../gpurun_out/e2_ab_r1_$rep.json 2> ../../gpurun_out/e2_ab_r1_$rep.err)
  python - <<PY
import json
d=json.loads(open("gpurun_out/e2_ab_r1_$rep.json").read().strip().splitlines()[-1])
print("r1  lib:", round(d["ms_per_step"],4), "ms/step", [(k["kernel"], round(k["avg_ms"],4)) for k in d["roofline"]["kernels"]])
PY
  timeout -s KILL 200 python bench.py --grid 256 --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-workloads --no-cfg2 > gpurun_out/e2_ab_now_$rep.json 2> gpurun_out/e2_ab_now_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/e2_ab_now_$rep.json").read().strip().splitlines()[-1])
print("now lib:", round(d["ms_per_step"],4), "ms/step", [(k["kernel"], round(k["avg_ms"],4)) for k in d["roofline"]["kernels"]])
PY
done
timeout -s KILL 300 python bench.py --workload pfc --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/e2_bench_pfc.json 2> gpurun_out/e2_bench_pfc.err
echo "bench pfc rc=$?"; python scripts/show_bench.py gpurun_out/e2_bench_pfc.json; tail -3 gpurun_out/e2_bench_pfc.err
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_kspace -s 2 -c 1 -f -o gpurun_out/r2_prof_pfc_kspace python scripts/profile_pfc.py 512 > gpurun_out/e2_ncu_pfc.log 2>&1
tail -2 gpurun_out/e2_ncu_pfc.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_kspace -s 2 -c 1 -f -o gpurun_out/r2_prof_256_kspace python scripts/profile_workload.py 256 > gpurun_out/e2_ncu_256.log 2>&1
tail -2 gpurun_out/e2_ncu_256.log
