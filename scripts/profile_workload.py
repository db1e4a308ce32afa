"""Small driver for ncu captures: a few plain 256^3 transforms (axis 2, 1, 0 passes)
followed by a few fused Cahn-Hilliard steps.  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:'k_pass|k_fused' -s 12 -c 10 \
      -o gpurun_out/prof python scripts/profile_workload.py [grid]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import pfutil as gpfutil  # noqa: E402
from gopf_b200 import synthetic  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = G ** 3
model = gpf.NewModel()
conc = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
model.AddScalar(gpf.NewScalar("gamma", 2.0))
model.AddScalar(gpf.NewScalar("m1", -1.0))
model.AddField(conc)
model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
solver = gpf.NewSolver(model, [G, G, G], 0.1)
solver.Upload()          # 3 plain forward passes (axis 2, 1, 0)
solver.StepDevice(6)     # 1 plain inverse axis-0 pass, then 6 x (mid inv, real, mid fwd, kspace)
solver.Download()        # 3 plain inverse passes
print("done", float(np.abs(conc.Data).max()))
