#!/bin/bash
# round 2, call e6b: gradient-operator tests; ncu evidence with the final library -- full capture of one 1024^3 step
# and one 256^3 step and of the cfg 5 k-space kernel (raw pages exported to CSV here: the reports themselves exceed
# what gpurun copies back), and the launch list of the bench command itself
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_terms_gradient_gpu.py -x -q > gpurun_out/e6_grad_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e6_grad_tests.log; tail -12 gpurun_out/e6_grad_tests.log
cap() {  # name, skip, count, regex, script args...
  local name=$1 skip=$2 count=$3 regex=$4; shift 4
  timeout -s KILL 600 ncu --set full --clock-control none -k regex:"$regex" -s $skip -c $count -f -o /tmp/$name python "$@" > gpurun_out/e6_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  tail -1 gpurun_out/e6_ncu_$name.log; ls -la gpurun_out/$name.raw.csv
}
cap r2_prof_1024 8 4 'k_pass|k_fused' scripts/profile_1024.py
cap r2_prof_256 12 4 'k_pass|k_fused' scripts/profile_workload.py 256
cap r2_prof_pfc_kspace 2 1 'k_fused_kspace' scripts/profile_pfc.py 512
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --blocks 1 --no-cpu-baseline --no-parity > gpurun_out/e6_bench_under_ncu.log 2>&1
tail -c 200 gpurun_out/e6_bench_under_ncu.log; wc -l gpurun_out/r2_launches_bench.csv
for tx in 4 16; do
  GOPF_KSPACE_TX=$tx timeout -s KILL 200 python bench.py --workload pfc --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/e6_bench_pfc_tx$tx.json 2> gpurun_out/e6_bench_pfc_tx$tx.err
  echo "bench pfc TX=$tx rc=$?"; python scripts/show_bench.py gpurun_out/e6_bench_pfc_tx$tx.json; tail -2 gpurun_out/e6_bench_pfc_tx$tx.err
done
du -sh gpurun_out
