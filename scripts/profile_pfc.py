"""ncu driver: cfg 5 (phase-field crystal, k-space noise, Vandeven filter) on the tabulated fused kernels, a few steps.
  ncu --set full --clock-control none --import-source on -k regex:k_fused_kspace -s 2 -c 1 -o ... python scripts/profile_pfc.py [grid]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import workloads  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m, f, solver = workloads.build_pfc(gpf, gpf, [G, G, G], noise="device", kspace_noise=gpf.HasKSpaceNoise())
solver.Upload()
solver.StepDevice(4)
solver.Synchronize()
print("done", solver.FusedForm() if hasattr(solver, "FusedForm") else "")
