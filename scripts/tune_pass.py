"""Micro-benchmark of single axis passes through gopf_fft_exec_axis_device.
  python scripts/tune_pass.py            # default sweep
Prints GB/s (32 B per cell per pass) for several lengths, axes and tile widths."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pfutil as gpfutil  # noqa: E402
from gopf_b200._lib import check, lib  # noqa: E402


def bench(dims, axis, tx, reps=10):
    n = 1
    for d in dims:
        n *= d
    plan = gpfutil.NewFFTW(dims)
    a = torch.randn(n, dtype=torch.complex128, device="cuda")
    b = torch.empty_like(a)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        def run():
            check(lib().gopf_fft_exec_axis_device(plan._h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                                                  -1, axis, tx, ctypes.c_void_p(stream.cuda_stream)))
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            run()
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return 32.0 * n / (ms * 1e-3) / 1e9, ms


if __name__ == "__main__":
    cases = [
        ([256, 256, 256], 1), ([256, 256, 256], 0),
        ([512, 64, 512], 1), ([512, 512, 512], 1), ([512, 64, 512], 0), ([64, 512, 512], 2),
        ([1024, 64, 1024], 0), ([64, 1024, 1024], 1), ([64, 1024, 1024], 2),
        ([128, 128, 128], 1), ([2048, 32, 2048], 0),
    ]
    for dims, axis in cases:
        row = []
        for tx in ((0,) if axis == 2 else (2, 4, 8, 16)):
            try:
                gbs, ms = bench(dims, axis, tx)
                row.append(f"tx={tx}: {gbs:7.0f} GB/s ({ms:.3f} ms)")
            except Exception as e:  # unsupported tile width for this length
                row.append(f"tx={tx}: n/a")
        print(f"{'x'.join(map(str, dims)):>14s} axis {axis} N={dims[axis]:5d}  " + "  ".join(row), flush=True)
