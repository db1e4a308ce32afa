"""Sweep of the strided-pass tile width (GOPF_PASS_TX) and of the next-wave L2 prefetch
(GOPF_PREFETCH) on single axis passes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tune_pass import bench  # noqa: E402

cases = [([256, 256, 256], 0), ([256, 256, 256], 1), ([512, 512, 512], 0), ([512, 512, 512], 1), ([1024, 64, 1024], 0),
         ([64, 1024, 1024], 1), ([1024, 128, 1024], 0)]
for dims, axis in cases:
    row = []
    for pf in (0, 1):
        for tx in (4, 8, 16):
            os.environ["GOPF_PREFETCH"] = str(pf)
            os.environ["GOPF_PASS_TX"] = str(tx)
            try:
                gbs, ms = bench(dims, axis, 0, reps=5)
                row.append(f"pf{pf} tx{tx}: {gbs:5.0f}")
            except Exception:
                row.append(f"pf{pf} tx{tx}:  n/a")
    print(f"{'x'.join(map(str, dims)):>14s} axis {axis}  " + "  ".join(row), flush=True)
