"""Probe of the peer-to-peer bandwidth between GPU 0 and GPU 1 of the box (single process):
DMA copies (cudaMemcpyPeerAsync through torch) one way and both ways at once.  Context for the
NVLink roofline of the peer-store exchange."""
import torch

assert torch.cuda.device_count() >= 2
n = 1 << 30  # 4 GiB of float32
a0 = torch.empty(n, dtype=torch.float32, device="cuda:0")
b1 = torch.empty(n, dtype=torch.float32, device="cuda:1")
a1 = torch.empty(n, dtype=torch.float32, device="cuda:1")
b0 = torch.empty(n, dtype=torch.float32, device="cuda:0")
print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1))
s0 = torch.cuda.Stream(device="cuda:0")
s1 = torch.cuda.Stream(device="cuda:1")


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.device(0):
            e0.record(s0)
        fn()
        with torch.cuda.device(0):
            e1.record(s0)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        best = min(best, e0.elapsed_time(e1))
    return best


def one_way():
    with torch.cuda.device(0), torch.cuda.stream(s0):
        b1.copy_(a0, non_blocking=True)


def both_ways():
    with torch.cuda.device(1), torch.cuda.stream(s1):
        b0.copy_(a1, non_blocking=True)
    with torch.cuda.device(0), torch.cuda.stream(s0):
        b1.copy_(a0, non_blocking=True)


gb = 4 * n / 1e9
t = timed(one_way)
print(f"one way   0->1: {gb / (t * 1e-3):7.1f} GB/s ({t:.2f} ms)")
t = timed(both_ways)
print(f"both ways 0->1 (1->0 concurrently): {gb / (t * 1e-3):7.1f} GB/s per direction, lower bound ({t:.2f} ms)")
