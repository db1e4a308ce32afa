"""ncu driver: 2-D Cahn-Hilliard [1024, 8192] (k-space kernel along 1024-cell lines, 2048 tiles), a few steps.
  GOPF_TMA_KSPACE=1 ncu --set full --clock-control none --import-source on -k regex:k_fused_kspace -s 1 -c 1 -o ... python scripts/profile_kspace.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import pfutil as gpfutil  # noqa: E402
from gopf_b200 import synthetic  # noqa: E402

dims = [int(a) for a in sys.argv[1:]] or [1024, 8192]
n = int(np.prod(dims))
m = gpf.NewModel()
f = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
m.AddScalar(gpf.NewScalar("gamma", 2.0))
m.AddScalar(gpf.NewScalar("m1", -1.0))
m.AddField(f)
m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
s = gpf.NewSolver(m, dims, 0.1)
s.Upload()
s.StepDevice(3)
s.Synchronize()
print("done", gpfutil.TmaLaunchCount())
