"""Tuning of the fused k-space kernel variants (GOPF_KSPACE_LATE / GOPF_KSPACE_TX) on one GPU:
per-kernel CUDA-event times of the Cahn-Hilliard step at several grid sizes, and a bitwise
comparison of the variants' results.  Usage: python scripts/tune_kspace.py [256 512 1024]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import synthetic  # noqa: E402

grids = [int(a) for a in sys.argv[1:]] or [256, 512, 1024]
variants = {256: [(0, 0), (1, 4), (1, 8), (1, 16)], 512: [(0, 0), (1, 4), (1, 8), (1, 16)], 1024: [(0, 0), (1, 2), (1, 4), (1, 8)]}
for G in grids:
    n = G ** 3
    model = gpf.NewModel()
    conc = gpf.NewField("conc", n, None, pinned=True)
    if G <= 256:
        synthetic.cahn_hilliard_initial(n, 0, out=conc.Data)
    else:  # timing only: a cheap smooth field
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
    init = conc.Data.copy() if G <= 256 else None
    ref = None
    for late, tx in variants[G]:
        os.environ["GOPF_KSPACE_LATE"] = str(late)
        os.environ["GOPF_KSPACE_TX"] = str(tx)
        if init is not None:
            conc.Data[:] = init
        solver.Upload()
        solver.StepDevice(5)
        if init is not None:
            solver.Download()
            if ref is None:
                ref = conc.Data.copy()
            same = bool(np.array_equal(ref, conc.Data))
            solver.Upload()
        else:
            same = None
        torch.cuda.synchronize()
        solver.ProfileBegin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        solver.StepDevice(10)
        e1.record(stream)
        torch.cuda.synchronize()
        prof = solver.ProfileEnd()
        ms = e0.elapsed_time(e1) / 10
        ks = [k for k in prof if k["kernel"] == "fused_kspace"][0]
        kms = ks["total_ms"] / ks["launches"]
        print(f"{G}^3 late={late} tx={tx or 'auto'}: step {ms:.3f} ms = {n / ms / 1e6:.2f} G cell-updates/s; "
              f"fused_kspace {kms:.3f} ms = {64.0 * n / kms / 1e6:.0f} GB/s; bitwise same as first variant: {same}", flush=True)
        others = ", ".join(f"{k['kernel']} {k['total_ms'] / k['launches']:.3f}" for k in prof if k["kernel"] != "fused_kspace" and k["launches"])
        print(f"      {others}", flush=True)
    del solver, model, conc
    torch.cuda.empty_cache()
