"""Driver for ncu captures of the 1024-cell-line kernels: a few fused Cahn-Hilliard steps at G^3
with a cheap initial field.  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:'k_pass|k_fused' -s 8 -c 4 \
      -o gpurun_out/prof1024 python scripts/profile_1024.py [grid]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import synthetic  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = G ** 3
model = gpf.NewModel()
conc = gpf.NewField("conc", n, None, pinned=True)
conc.Data[:] = 0.0
conc.Data[::7] = 0.5
model.AddScalar(gpf.NewScalar("gamma", 2.0))
model.AddScalar(gpf.NewScalar("m1", -1.0))
model.AddField(conc)
model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
solver = gpf.NewSolver(model, [G, G, G], 0.1)
solver.Upload()       # 3 plain forward passes
solver.StepDevice(4)  # 1 plain inverse pass, then 4 x (mid inverse, real, mid forward, k-space)
solver.Synchronize()
print("done")
