#!/bin/bash
# round 2, call e6: ncu evidence with the final library -- full capture of one 1024^3 step and one 256^3 step,
# full capture of the cfg 5 k-space kernel, and the launch list of the bench command itself
set -u
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'k_pass|k_fused' -s 8 -c 4 -f -o gpurun_out/r2_prof_1024 python scripts/profile_1024.py > gpurun_out/e6_ncu_1024.log 2>&1
tail -2 gpurun_out/e6_ncu_1024.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:'k_pass|k_fused' -s 12 -c 4 -f -o gpurun_out/r2_prof_256 python scripts/profile_workload.py 256 > gpurun_out/e6_ncu_256.log 2>&1
tail -2 gpurun_out/e6_ncu_256.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_kspace -s 2 -c 1 -f -o gpurun_out/r2_prof_pfc_kspace_v2 python scripts/profile_pfc.py 512 > gpurun_out/e6_ncu_pfc.log 2>&1
tail -2 gpurun_out/e6_ncu_pfc.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --blocks 1 --no-cpu-baseline --no-parity > gpurun_out/e6_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/e6_bench_under_ncu.log; wc -l gpurun_out/r2_launches_bench.csv
for tx in 4 16; do
  GOPF_KSPACE_TX=$tx timeout -s KILL 200 python bench.py --workload pfc --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/e6_bench_pfc_tx$tx.json 2> gpurun_out/e6_bench_pfc_tx$tx.err
  echo "bench pfc TX=$tx rc=$?"; python scripts/show_bench.py gpurun_out/e6_bench_pfc_tx$tx.json; tail -2 gpurun_out/e6_bench_pfc_tx$tx.err
done
