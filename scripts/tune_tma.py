"""Per-kernel CUDA-event times of the fused Cahn-Hilliard step with the copy-engine-fed (TMA) long-line
kernels switched on / off and with different tensor-map L2 promotions.

    python scripts/tune_tma.py 1024 [512]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import pfutil as gpfutil  # noqa: E402
from gopf_b200 import synthetic  # noqa: E402

BYTES = {"fused_kspace": 64.0, "fused_real": 32.0, "pass_inverse_mid": 32.0, "pass_forward_mid": 32.0}
grids = [int(a) for a in sys.argv[1:]] or [1024]
VARIANTS = [("register kernels", {"GOPF_TMA": "0", "GOPF_BLOCKED": "0"}),
            ("register kernels, blocked s=7", {"GOPF_TMA": "0", "GOPF_BLOCKED": "1"}),
            ("tma all, blocked s=7", {"GOPF_TMA": "1", "GOPF_BLOCKED": "1"}),
            ("tma all, blocked s=6", {"GOPF_TMA": "1", "GOPF_BLOCKED": "1", "GOPF_BLOCK_LOG": "6"}),
            ("tma all, blocked s=8", {"GOPF_TMA": "1", "GOPF_BLOCKED": "1", "GOPF_BLOCK_LOG": "8"}),
            ("tma pass+real, register kspace, blocked s=7", {"GOPF_TMA": "1", "GOPF_BLOCKED": "1", "GOPF_TMA_KSPACE": "0"}),
            ("tma pass+real, register kspace, row-major", {"GOPF_TMA": "1", "GOPF_BLOCKED": "0", "GOPF_TMA_KSPACE": "0"}),
            ("tma all, row-major", {"GOPF_TMA": "1", "GOPF_BLOCKED": "0"}),
            ("tma all, blocked s=7, spectrum prefetch at compute start", {"GOPF_TMA": "1", "GOPF_BLOCKED": "1", "GOPF_TMA_SPF": "1"}),
            ("tma all, blocked s=7, no spectrum prefetch", {"GOPF_TMA": "1", "GOPF_BLOCKED": "1", "GOPF_TMA_SPF": "2"}),
            ("final, complex-carrying real-space kernel", {"GOPF_REAL_PAIRS": "0"}),
            ("final, paired real lines", {"GOPF_REAL_PAIRS": "1"})]
if os.environ.get("TUNE_VARIANTS"):
    VARIANTS = [v for v in VARIANTS if any(w in v[0] for w in os.environ["TUNE_VARIANTS"].split(";"))]
KEYS = ("GOPF_TMA", "GOPF_TMA_L2", "GOPF_TMA_PASS", "GOPF_TMA_REAL", "GOPF_TMA_KSPACE", "GOPF_BLOCKED", "GOPF_BLOCK_LOG", "GOPF_TMA_SPF", "GOPF_REAL_PAIRS")
for G in grids:
    n = G ** 3
    os.environ["GOPF_TMA_MIN_N"] = str(min(G, 1024))
    model = gpf.NewModel()
    conc = gpf.NewField("conc", n, None)
    conc.Data[::7] = 0.5  # timing only: a cheap field (the kernels have no data-dependent branches)
    model.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    model.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    model.AddField(conc)
    model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    solver = gpf.NewSolver(model, [G, G, G], synthetic.CAHN_HILLIARD_DT, device=0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    solver.SetStream(stream.cuda_stream)
    print("grid", G, "upload", file=sys.stderr, flush=True)
    solver.Upload()
    torch.cuda.synchronize()
    print("  uploaded", file=sys.stderr, flush=True)
    steps = 10 if G >= 1024 else 30
    for name, env in VARIANTS:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            print("variant", name, file=sys.stderr, flush=True)
            gpfutil.TmaLaunchCount(reset=True)
            solver.StepDevice(3)
            torch.cuda.synchronize()
            print("  warm", file=sys.stderr, flush=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            solver.StepDevice(steps)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            solver.ProfileBegin()
            solver.StepDevice(steps)
            torch.cuda.synchronize()
            prof = solver.ProfileEnd()
            rec = {"grid": G, "variant": name, "ms_per_step": round(ms, 4), "g_cell_updates_per_s": round(n / ms / 1e6, 2),
                   "tma_launches": gpfutil.TmaLaunchCount()}
            for k in prof:
                if k["launches"]:
                    kms = k["total_ms"] / k["launches"]
                    rec[k["kernel"]] = {"ms": round(kms, 4), "gbs": round(BYTES.get(k["kernel"], 32.0) * n / kms / 1e6)}
            print(json.dumps(rec), flush=True)
        except Exception as exc:
            print(json.dumps({"grid": G, "variant": name, "error": str(exc)[:300]}), flush=True)
    solver.close()
    del solver, model, conc
    torch.cuda.empty_cache()
