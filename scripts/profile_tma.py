"""Small driver for ncu captures of the copy-engine kernels: one 3-D transform with a 1024-cell middle axis and a few
2-D 1024^2 Cahn-Hilliard steps (k_pass_strided_tma, k_fused_real_tma, k_fused_kspace_tma).
  ncu --set full --clock-control none --import-source on -k regex:tma -c 6 -o gpurun_out/prof_tma python scripts/profile_tma.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import pfutil as gpfutil  # noqa: E402
from gopf_b200 import synthetic  # noqa: E402

dims = [32, 1024, 1024]
x = np.zeros(int(np.prod(dims)), dtype=np.complex128)
x[::5] = 1.0
ft = gpfutil.NewFFTW(dims)
ft.FFT(x)
ft.FFT(x)
ft.close()
G = 1024
n = G * G
m = gpf.NewModel()
f = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
m.AddScalar(gpf.NewScalar("gamma", 2.0))
m.AddScalar(gpf.NewScalar("m1", -1.0))
m.AddField(f)
m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
s = gpf.NewSolver(m, [G, G], 0.1)
s.Upload()
s.StepDevice(3)
s.Synchronize()
print("done", gpfutil.TmaLaunchCount())
