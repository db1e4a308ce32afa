"""Vendor-library comparison point: cuFFT Z2Z (through torch.fft) on the bench grids, as GB/s of the 96 B/cell model."""
import sys
import torch

for n in [int(a) for a in sys.argv[1:]] or [256, 512, 1024]:
    x = torch.zeros((n, n, n), dtype=torch.complex128, device="cuda")
    x[0, 0, 1] = 1.0
    for _ in range(2):
        y = torch.fft.fftn(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        y = torch.fft.fftn(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"cufft z2z {n}^3 out-of-place: {ms:.3f} ms  {96.0 * n**3 / ms * 1e-6:.0f} GB/s of the 96 B/cell model", flush=True)
    del x, y
