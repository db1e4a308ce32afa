"""A/B of the next-wave L2 prefetch (GOPF_PREFETCH=0/1): per-kernel CUDA-event times of the fused
Cahn-Hilliard step on one GPU.  Usage: python scripts/tune_prefetch.py [256 512 1024]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import synthetic  # noqa: E402

for G in [int(a) for a in sys.argv[1:]] or [256, 512, 1024]:
    n = G ** 3
    model = gpf.NewModel()
    conc = gpf.NewField("conc", n, None, pinned=True)
    conc.Data[:] = 0.0
    conc.Data[::7] = 0.5
    model.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    model.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    model.AddField(conc)
    model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    solver = gpf.NewSolver(model, [G, G, G], synthetic.CAHN_HILLIARD_DT, device=0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    solver.SetStream(stream.cuda_stream)
    solver.Upload()
    for pf in (0, 1, 0, 1):
        os.environ["GOPF_PREFETCH"] = str(pf)
        solver.StepDevice(5)
        torch.cuda.synchronize()
        solver.ProfileBegin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        solver.StepDevice(10)
        e1.record(stream)
        torch.cuda.synchronize()
        prof = solver.ProfileEnd()
        ms = e0.elapsed_time(e1) / 10
        ks = ", ".join(f"{k['kernel']} {k['total_ms'] / k['launches']:.3f} ms ({k['bytes_per_launch'] / (k['total_ms'] / k['launches']) / 1e6:.0f} GB/s)"
                       for k in prof if k["launches"])
        print(f"{G}^3 prefetch={pf}: step {ms:.3f} ms = {n / ms / 1e6:.2f} G cell-updates/s | {ks}", flush=True)
    del solver, model, conc
    torch.cuda.empty_cache()
