#!/bin/bash
# round 2, call e16: ncu --set full of the paired real-space kernel at 1024^3 (raw page as CSV)
set -u
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none -k regex:'k_fused_real_pair' -s 1 -c 1 -f -o /tmp/r2_prof_1024_pair python scripts/profile_1024.py > gpurun_out/e16_ncu.log 2>&1
ncu -i /tmp/r2_prof_1024_pair.ncu-rep --page raw --csv > gpurun_out/r2_prof_1024_pair.raw.csv 2>/dev/null
tail -2 gpurun_out/e16_ncu.log; ls -la gpurun_out/r2_prof_1024_pair.raw.csv
