"""Strided line passes on the blocked k-space layout [n0/2^s][n1][2^s][n2] against the row-major one.

  python scripts/tune_blocked.py check        # correctness vs torch.fft on a small grid (both kernel families)
  python scripts/tune_blocked.py [n]          # GB/s (32 B per cell per pass) of the axis-0 and axis-1 passes, n^3

Axis-0 lines of a row-major n^3 array have a row stride of n^2 cells (16 MB at n = 1024: every row of a tile in
its own 2-MB page); in the blocked layout the 2^s rows of a block are n2 cells apart and only the n0/2^s blocks
are far from each other.  The middle-axis passes then read row-major and write blocked (or the reverse)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gopf_b200 import pfutil as gpfutil  # noqa: E402
from gopf_b200._lib import check, lib  # noqa: E402

NONE = 31


def std_axis0(n0, n1, n2):  # slabs = 1, cols = n1*n2
    return 1, n1 * n2, [n0 * n1 * n2, n1 * n2, 0, NONE, 0, NONE]


def std_axis1(n0, n1, n2):  # slabs = n0, cols = n2
    return n0, n2, [n1 * n2, n2, 0, NONE, 0, NONE]


def blk_axis0(n0, n1, n2, s):  # lines along axis 0 of the blocked array: slabs = n1 (axis-1 index), cols = n2
    return n1, n2, [(1 << s) * n2, n2, n1 * (1 << s) * n2, s, 0, NONE]


def blk_axis1(n0, n1, n2, s):  # lines along axis 1 of the blocked array: slabs = n0 (split), cols = n2
    return n0, n2, [n2, (1 << s) * n2, 0, NONE, n1 * (1 << s) * n2, s]


def to_blocked(x, s):
    n0, n1, n2 = x.shape
    return x.view(n0 >> s, 1 << s, n1, n2).permute(0, 2, 1, 3).contiguous()


def from_blocked(y, dims, s):
    n0, n1, n2 = dims
    return y.view(n0 >> s, n1, 1 << s, n2).permute(0, 2, 1, 3).contiguous().view(n0, n1, n2)


def run_pass(plan, a, b, sign, axis, slabs, cols, imap, omap, stream=None, tx=0):
    im = (ctypes.c_int64 * 6)(*imap)
    om = (ctypes.c_int64 * 6)(*omap)
    check(lib().gopf_fft_exec_rows_device(plan._h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), sign, axis,
                                          ctypes.c_int64(slabs), ctypes.c_int64(cols), im, om, tx,
                                          ctypes.c_void_p(stream.cuda_stream if stream else 0)))


def timeit(fn, reps=5):
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for _ in range(2):
            fn(stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn(stream)
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def check_small():
    dims = [512, 16, 64]
    s = 3
    plan = gpfutil.NewFFTW(dims)
    x = torch.randn(dims, dtype=torch.complex128, device="cuda")
    worst = 0.0
    for tma in ("0", "1"):
        os.environ["GOPF_TMA"] = tma
        os.environ["GOPF_TMA_MIN_N"] = "512"
        before = gpfutil.TmaLaunchCount()
        # axis 0 on the blocked array, in place
        xb = to_blocked(x, s)
        slabs, cols, m = blk_axis0(*dims, s)
        run_pass(plan, xb, xb, -1, 0, slabs, cols, m, m)
        torch.cuda.synchronize()
        ref = torch.fft.fft(x, dim=0)
        err0 = (from_blocked(xb, dims, s) - ref).abs().max().item() / ref.abs().max().item()
        # axis 1: row-major in, blocked out, and back
        plan1 = gpfutil.NewFFTW([16, 512, 64])
        d1 = [16, 512, 64]
        x1 = torch.randn(d1, dtype=torch.complex128, device="cuda")
        y1 = torch.empty_like(x1)
        sl, co, mi = std_axis1(*d1)
        _, _, mo = blk_axis1(*d1, 2)
        run_pass(plan1, x1, y1, -1, 1, sl, co, mi, mo)
        torch.cuda.synchronize()
        ref1 = torch.fft.fft(x1, dim=1)
        err1 = (from_blocked(y1.view(-1), d1, 2) - ref1).abs().max().item() / ref1.abs().max().item()
        z1 = torch.empty_like(x1)
        run_pass(plan1, y1, z1, 1, 1, sl, co, mo, mi)
        torch.cuda.synchronize()
        err2 = (z1 / 512 - x1).abs().max().item()
        used = gpfutil.TmaLaunchCount() - before
        print(f"GOPF_TMA={tma}: axis-0 blocked {err0:.2e}  axis-1 row-major->blocked {err1:.2e}  round trip {err2:.2e}  "
              f"copy-engine launches {used}", flush=True)
        worst = max(worst, err0, err1, err2)
    print("ok" if worst < 1e-12 else "MISMATCH", flush=True)
    return worst < 1e-12


def sweep(n):
    dims = [n, n, n]
    cells = n ** 3
    plan = gpfutil.NewFFTW(dims)
    a = torch.zeros(cells, dtype=torch.complex128, device="cuda")
    b = torch.zeros(cells, dtype=torch.complex128, device="cuda")
    a[1] = 1.0

    def report(name, slabs, cols, imap, omap, axis, inplace):
        row = []
        for tma in ("0", "1"):
            os.environ["GOPF_TMA"] = tma
            try:
                ms = timeit(lambda st: run_pass(plan, a, a if inplace else b, -1, axis, slabs, cols, imap, omap, st))
                row.append(f"tma={tma}: {32.0 * cells / ms * 1e-6:5.0f} GB/s ({ms:7.3f} ms)")
            except Exception as exc:  # noqa: BLE001
                row.append(f"tma={tma}: {str(exc)[:50]}")
        print(f"{name:<46s} {'in place ' if inplace else 'out of pl'}  " + "  ".join(row), flush=True)

    sl, co, m = std_axis1(*dims)
    report("axis 1 row-major", sl, co, m, m, 1, True)
    report("axis 1 row-major", sl, co, m, m, 1, False)
    sl0, co0, m0 = std_axis0(*dims)
    report("axis 0 row-major", sl0, co0, m0, m0, 0, True)
    for s in (3, 4, 5, 6, 7):
        slb, cob, mb = blk_axis0(*dims, s)
        report(f"axis 0 blocked s={s}", slb, cob, mb, mb, 0, True)
        _, _, mo = blk_axis1(*dims, s)
        report(f"axis 1 row-major -> blocked s={s}", sl, co, m, mo, 1, False)
        report(f"axis 1 blocked s={s} -> row-major", sl, co, mo, m, 1, False)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "check":
        sys.exit(0 if check_small() else 1)
    sweep(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
