#!/bin/bash
# GPU runbook for the work that ended round 1 without GPU time.  Each stage is one
# gpurun call; outputs land in gpurun_out/ and the lines worth keeping are copied to profiles/ by hand.
#
#   gpurun --timeout 900 -- 'bash scripts/round2_gpu_checks.sh suite'
#   gpurun --timeout 900 -- 'bash scripts/round2_gpu_checks.sh jit'
#   gpurun --timeout 900 -- 'bash scripts/round2_gpu_checks.sh inpass'
#   gpurun --timeout 900 -- 'bash scripts/round2_gpu_checks.sh knoise'
set -u
mkdir -p gpurun_out
stage=${1:-suite}

case "$stage" in
suite)
    # 1. the device suite on the default kernels, then with the run-time specialisation switched on for
    #    every solver: the second run is the evidence needed to make GOPF_JIT the default
    python -m pytest tests -m gpu -x -q > gpurun_out/suite_default.log 2>&1; echo "default rc=$?" >> gpurun_out/suite_default.log
    GOPF_JIT=1 python -m pytest tests -m gpu -q > gpurun_out/suite_jit.log 2>&1; echo "jit rc=$?" >> gpurun_out/suite_jit.log
    tail -3 gpurun_out/suite_default.log gpurun_out/suite_jit.log
    ;;
jit)
    # 2. cfg 4 / cfg 5 at the bench size with the interpreter kernels and with the NVRTC images, per-kernel
    #    events; then one ncu pass over the specialised kernels (their cubins and sources are dumped for
    #    --import-source)
    python scripts/jit_check.py 512 5 > gpurun_out/jit_check_512.log 2>&1
    for w in precipitate pfc ch_sqgrad; do
        python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${w}_jit.json 2> gpurun_out/bench_${w}_jit.err
        python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-jit > gpurun_out/bench_${w}_nojit.json 2> gpurun_out/bench_${w}_nojit.err
    done
    mkdir -p gpurun_out/jit_dump
    GOPF_JIT_DUMP=gpurun_out/jit_dump ncu --set full --clock-control none --import-source on \
        -k regex:gopf_jit -c 6 -o gpurun_out/jit_kernels python scripts/jit_check.py 256 1 > gpurun_out/ncu_jit.log 2>&1
    tail -2 gpurun_out/jit_check_512.log
    ;;
inpass)
    # 2b. registered functions compiled into the load of their first forward pass (GOPF_JIT_INPASS)
    GOPF_TEST_INPASS=1 python -m pytest tests/test_zz_jit_gpu.py -q -k forward_pass > gpurun_out/inpass_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/inpass_pytest.log
    GOPF_JIT_INPASS=1 python scripts/jit_check.py 256 5 > gpurun_out/jit_check_inpass_256.log 2>&1
    GOPF_JIT_INPASS=1 python scripts/jit_check.py 512 5 > gpurun_out/jit_check_inpass_512.log 2>&1
    tail -3 gpurun_out/inpass_pytest.log; tail -2 gpurun_out/jit_check_inpass_256.log
    ;;
knoise)
    # 3. white noise drawn in k-space: rebuild the device code with the generator, check that the fast-form
    #    kernels did not move, run the device suite on that build, bench cfg 5 both ways
    make -C gopf_b200/csrc -j16 EXTRA=-DGOPF_KNOISE OBJDIR=../../build/obj_knoise > gpurun_out/knoise_build.log 2>&1
    python -c "from gopf_b200 import pf; assert pf.HasKSpaceNoise()" || exit 1
    python -m pytest tests -m gpu -x -q > gpurun_out/suite_knoise.log 2>&1; echo "knoise rc=$?" >> gpurun_out/suite_knoise.log
    python - > gpurun_out/knoise_cfg5.log 2>&1 <<'EOF'
import json, math, sys, time
sys.path.insert(0, ".")
import numpy as np
from gopf_b200 import pf as gpf, workloads

def build(n_edge, kspace):
    dims = [n_edge] * 3
    n = n_edge ** 3
    m = gpf.NewModel()
    f = gpf.NewField("density", n, workloads.pfc_initial(n))
    m.AddField(f)
    a = workloads.PFC_LATTICE
    peaks = [gpf.Peak(1.0, 2.0 * math.pi / a, workloads.PFC_PEAK_WIDTH, 4),
             gpf.Peak(1.0 / math.sqrt(2.0), 2.0 * math.pi / (a / math.sqrt(2.0)), workloads.PFC_PEAK_WIDTH, 4)]
    m.RegisterImplicitTerm("EXCESS", gpf.PairCorrlationTerm(gpf.ReciprocalSpacePairCorrelation(workloads.PFC_EFF_TEMP, peaks), "density", 1.0, True), None)
    ideal = gpf.IdealMixtureTerm(gpf.IdealMix(1.0, 1.0), "density", 1.0, True)
    m.RegisterMixedTerm("IDEAL", ideal, [ideal.DerivedField(n, m.Bricks)])
    m.RegisterFunction("NOISE", gpf.WhiteNoise(workloads.PFC_NOISE_STRENGTH, seed=7).Generate)
    m.AddEquation("ddensity/dt = IDEAL + EXCESS + NOISE")
    m.SetKSpaceNoise(kspace)
    s = gpf.NewSolver(m, dims, workloads.PFC_DT)
    s.Stepper.SetFilter(gpf.NewVandeven(5))
    return m, f, s

for edge in (256, 512):
    for kspace in (False, True):
        m, f, s = build(edge, kspace)
        s.Upload(); s.StepDevice(5); s.Synchronize()
        t0 = time.perf_counter(); s.StepDevice(20); s.Synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / 20
        print(json.dumps({"grid": edge, "kspace_noise": kspace, "fused": s.IsFused, "ms_per_step": round(ms, 4),
                          "g_cell_updates_per_s": round(edge ** 3 / ms / 1e6, 2)}), flush=True)
        s.close()
EOF
    tail -3 gpurun_out/suite_knoise.log; cat gpurun_out/knoise_cfg5.log
    ;;
*)
    echo "unknown stage $stage"; exit 2
    ;;
esac
