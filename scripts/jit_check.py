"""GPU check of the run-time specialisation (gopf_b200/csrc/jit.h): cfg 4 (strain_single_precipitate)
and cfg 5 (pfcPhases + noise + filter) stepped with the interpreter kernels and with the NVRTC
images on the same inputs.  Appends one JSON line per run to gpurun_out/jit_check.jsonl: fields'
max abs difference between the two, kernels specialised, per-kernel time and algorithmic GB/s.
No torch, no oracle.

    python scripts/jit_check.py [grid edge, default 256] [steps, default 5]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from gopf_b200 import elasticity as gel  # noqa: E402
from gopf_b200 import pf as gpf  # noqa: E402
from gopf_b200 import workloads  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "jit_check.jsonl")


def emit(rec):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)


def run(kind, G, jit, steps):
    dims = [G, G, G]
    if kind == "precipitate":
        m, conc, phase, s, _ = workloads.build_precipitate(gpf, gpf, gel, dims, expressions=True)
        fields = [conc, phase]
    else:
        m, f, s = workloads.build_pfc(gpf, gpf, dims, noise="device")
        fields = [f]
    s.SetJit(jit)
    s.Upload()
    s.StepDevice(2)
    s.Synchronize()
    s.ProfileBegin()
    t0 = time.perf_counter()
    s.StepDevice(steps)
    s.Synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    prof = s.ProfileEnd()
    s.Download()
    out = [f.Data.copy() for f in fields]
    kernels = {}
    for k in prof:
        if k["launches"]:
            avg = k["total_ms"] / k["launches"]
            kernels[k["kernel"]] = {"avg_ms": round(avg, 4), "per_step": k["launches"] / steps,
                                    "gbs": round(k["bytes_per_launch"] / (avg * 1e-3) / 1e9, 1)}
    info = {"workload": kind, "grid": G, "jit": bool(jit), "jit_inpass": bool(jit) and os.environ.get("GOPF_JIT_INPASS") == "1",
            "jit_kernels": s.JitKernels(), "jit_log": s.JitLog(),
            "ms_per_step_profiled": round(wall_ms, 3), "kernels": kernels}
    s.close()
    return out, info


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    for kind in ("precipitate", "pfc"):
        try:
            ref, info0 = run(kind, G, False, steps)
            emit(info0)
            got, info1 = run(kind, G, True, steps)
            info1["max_abs_diff_vs_interpreter"] = [float(np.max(np.abs(a - b))) for a, b in zip(ref, got)]
            info1["max_abs_field"] = [float(np.max(np.abs(a))) for a in ref]
            emit(info1)
        except Exception as e:  # keep going: the other workload may still tell something
            emit({"workload": kind, "error": repr(e)})


if __name__ == "__main__":
    main()
