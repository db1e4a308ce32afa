"""Benchmark-scale golden fingerprints ("parity vs oracle at 64^3/128^3/256^3 x100").

A 256^3 trajectory is 256 MiB per snapshot, too large to commit, and the oracle needs minutes
per 100 steps, too slow for every test run.  This script runs the pinned oracle once (here, on
CPU) and freezes a FINGERPRINT of its state after 10 and after 100 semi-implicit Euler steps of
the benchmark's own seeded Cahn-Hilliard field (pf/euler.go:16-47 through oracle/pf.py):

  * the field values at 16 384 fixed pseudo-random cells (SplitMix64 indices),
  * sum, L2 norm and max|.| of the real part, max|imag|.

tests/test_golden_gpu.py compares the CUDA path with these at full size: relative L2 over the
sampled cells <= 1e-10 (an unbiased estimate of the full-field figure) and the norms to 1e-11.

    python tests/golden/make_golden_large.py [edge ...]      # default: 128 256
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gopf_b200 import synthetic  # noqa: E402  (seeded input stream only)
from oracle import pf as opf  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NSAMPLES = 16384


def sample_indices(n_cells: int) -> np.ndarray:
    u = synthetic.splitmix64_uniform(12345, NSAMPLES)
    return np.minimum((u * n_cells).astype(np.int64), n_cells - 1)


def fingerprint(data: np.ndarray, idx: np.ndarray) -> dict:
    re = data.real
    return {"samples": data[idx].copy(), "sum": np.float64(re.sum(dtype=np.longdouble)),
            "l2": np.float64(np.sqrt(np.sum(re.astype(np.longdouble) ** 2))), "max_abs": np.float64(np.max(np.abs(re))),
            "max_imag": np.float64(np.max(np.abs(data.imag)))}


def main(edges):
    workers = os.cpu_count() or 1
    for edge in edges:
        dims = [edge] * 3
        n = edge ** 3
        m = opf.NewModel()
        f = opf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
        m.AddScalar(opf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
        m.AddScalar(opf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
        m.AddField(f)
        m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
        s = opf.NewSolver(m, dims, synthetic.CAHN_HILLIARD_DT, workers=workers)
        idx = sample_indices(n)
        out = {"indices": idx, "edge": np.int64(edge)}
        done = 0
        t0 = time.perf_counter()
        for k in (10, 100):
            s.Propagate(k - done)
            done = k
            for key, val in fingerprint(f.Data, idx).items():
                out[f"after_{k}_{key}"] = val
            print(f"{edge}^3: {k} steps, {time.perf_counter() - t0:.1f} s", flush=True)
        np.savez_compressed(os.path.join(OUT, f"ch_3d_{edge}_fingerprint.npz"), **out)


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [128, 256])
