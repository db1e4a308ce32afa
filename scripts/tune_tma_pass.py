"""Size sweep of single strided passes with 1024-cell lines: register kernel vs copy-engine kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tune_pass import bench  # noqa: E402

cases = [([32, 1024, 1024], 1), ([128, 1024, 1024], 1), ([512, 1024, 1024], 1), ([1024, 1024, 1024], 1),
         ([1024, 32, 1024], 0), ([1024, 128, 1024], 0), ([1024, 1024, 1024], 0)]
for dims, axis in cases:
    row = []
    for tma in ("0", "1"):
        os.environ["GOPF_TMA"] = tma
        try:
            gbs, ms = bench(dims, axis, 0, reps=5)
            row.append(f"tma={tma}: {gbs:5.0f} GB/s ({ms:.3f} ms)")
        except Exception as exc:
            row.append(f"tma={tma}: {str(exc)[:60]}")
    print(f"{'x'.join(map(str, dims)):>16s} axis {axis}  " + "  ".join(row), flush=True)
