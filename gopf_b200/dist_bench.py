"""bench.py arm for N > 1: slab-sharded 3-D Cahn-Hilliard (BASELINE.json configs[2]),
one rank per GPU under torchrun.  The transpose exchange is fused into the producing passes
as peer stores over NVLink (--exchange peer, default) or runs as an NCCL all-to-all between
the local phases (--exchange nccl, the baseline).
Strong scaling: the grid is fixed (default 1024^3) as N grows."""
from __future__ import annotations

import json
import os
import time

import numpy as np

METRIC = "cell-updates/s"


class TimedPhases:
    """Wraps the phase object and the exchange with CUDA events (profiling region only)."""

    def __init__(self, phases, a2a, torch, stream):
        self.p, self.a2a, self.torch, self.stream = phases, a2a, torch, stream
        self.ev = {}

    def _timed(self, name, fn, *args):
        a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        fn(*args)
        b.record(self.stream)
        self.ev.setdefault(name, []).append((a, b))

    def __getattr__(self, name):
        fn = getattr(self.p, name)
        if name in ("advance", "get_time", "kernel_launches", "set_grid_cap", "peer_unmap", "sm_count"):
            return fn
        return lambda *args: self._timed(name, fn, *args)

    def all_to_all(self, dst, src):
        self._timed("all_to_all", self.a2a, dst, src)

    def barrier(self):
        self._timed("barrier", self.a2a)

    def summary(self):
        self.torch.cuda.synchronize()
        return {k: {"launches": len(v), "avg_ms": float(np.mean([a.elapsed_time(b) for a, b in v]))}
                for k, v in self.ev.items()}


def run(args):
    import torch
    import torch.distributed as tdist

    from . import dist as gdist
    from . import pf as gpf
    from . import pfutil as gpfutil
    from . import synthetic
    import bench as B  # ClockSampler, measured_hbm_peak

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.grid
    cells = n ** 3 // world
    exchange = getattr(args, "exchange", "peer")
    nchunks = getattr(args, "chunks", 8)
    # SMs given to the NVLink-bound peer-storing pass while it overlaps the next chunk, measured on the 1024^3 step with
    # this round's kernels (profiles/r2_tuning_notes.md):  P = 2: 48: 20.73, 64: 20.15, 80: 21.62, 96: 25.09 ms (the
    # pass also moves its own half through HBM);  P = 8: 48: 6.59, 64: 6.42, 76: 6.36 ms (round 1, slower compute
    # side: 48 best);  P = 4 not swept: 64.
    comm_ctas = getattr(args, "comm_ctas", 0) or {2: 64, 4: 64, 8: 76}.get(world, 48)
    # the same sharded CUDA path at 256^3 against the oracle, before anything is timed
    parity = None if getattr(args, "no_parity", False) else B.parity_sharded(world, rank, local, exchange, nchunks, comm_ctas)
    model = gpf.NewModel()
    conc = gpf.NewField("conc", cells, None, pinned=True)
    synthetic.cahn_hilliard_initial(cells, 0, offset=rank * cells, out=conc.Data)
    model.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    model.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    model.AddField(conc)
    model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    solver = gdist.ShardedSolver(model, n, synthetic.CAHN_HILLIARD_DT, device=local, exchange=exchange, nchunks=nchunks,
                                 comm_ctas=comm_ctas)
    exchange = solver.exchange  # what actually runs (peer mapping can fall back to nccl on every rank)
    stream = solver.stream

    solver.Upload()
    solver.StepDevice(args.warmup)
    solver.Synchronize()
    solver.phases.kernel_launches(reset=True)
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    tdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    solver.StepDevice(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    tdist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)  # the job's time is the slowest rank's
    ms = float(ms.item())
    launches = solver.phases.kernel_launches(reset=True)
    clocks = sampler.stop() if rank == 0 else None

    # per-phase CUDA events over an identical region
    with torch.cuda.stream(stream):
        if exchange == "peer":
            timed = TimedPhases(solver.phases, solver.barrier, torch, stream)
            solver.a_valid = gdist.run_steps_peer(timed, timed.barrier, solver.S, solver.A, args.steps, solver.a_valid,
                                                  solver.slab, nchunks, comm_ctas)
        elif exchange == "dma":
            timed = TimedPhases(solver.phases, solver.barrier, torch, stream)
            solver.a_valid = gdist.run_steps_dma(timed, timed.barrier, solver.S, solver.A, solver.B, args.steps,
                                                 solver.a_valid, solver.slab, nchunks)
        else:
            timed = TimedPhases(solver.phases, solver.all_to_all, torch, stream)
            solver.a_valid = gdist.run_steps(timed, timed.all_to_all, solver.S, solver.A, solver.B, args.steps,
                                             solver.a_valid)
    phases = timed.summary()

    # end to end: Solver.Propagate on the host slabs (H2D, step, D2H inside the timed region)
    e2e_steps = max(2, min(args.steps, 3))
    solver.Propagate(1)
    tdist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.Propagate(1)
    tdist.barrier()
    e2e_dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    tdist.all_reduce(e2e_dt, op=tdist.ReduceOp.MAX)
    e2e_dt = float(e2e_dt.item())

    if rank == 0:
        total = n ** 3
        value = total * args.steps / (ms * 1e-3)
        peak, peak_src = B.measured_hbm_peak()
        comp = {k: v for k, v in phases.items() if k not in ("all_to_all", "barrier", "forward_mid_peer_planes")
                and not k.startswith("exchange")}
        top = max(comp.items(), key=lambda kv: kv[1]["avg_ms"] * kv[1]["launches"])
        # algorithmic bytes of one launch: the phase's bytes per step / its launches per step (chunked runs)
        top_bytes = (64.0 if top[0].startswith("kspace_step") else 32.0) * cells / (top[1]["launches"] / args.steps)
        a2a_bytes_out = 16.0 * cells * (world - 1) / world  # per exchange, per GPU, each way
        if exchange == "peer":
            # the exchange rides inside the two peer-storing kernels; their duration bounds the
            # NVLink rate from below
            nv = {k: a2a_bytes_out / (phases[k]["avg_ms"] * 1e-3) / 1e9 for k in ("forward_mid_peer", "kspace_step_peer")
                  if k in phases}
            nvlink = {"bytes_out_per_exchange_per_gpu": a2a_bytes_out, "exchanges_per_step": 2,
                      "achieved_gbs_per_direction_lower_bound": nv, "peak_gbs_per_direction": 900.0,
                      "measured_peer_store_gbs_128B_segments": 717.0,
                      "barrier_avg_ms": phases.get("barrier", {}).get("avg_ms")}
            exch_desc = "fused into the producing passes: peer stores over NVLink (CUDA IPC), 2 stream barriers per step"
            if nchunks > 1:
                # forward_mid_peer_planes runs on the second stream: its events on the compute stream
                # time the launch only; what the step sees of it is the wait in exchange_join
                nvlink["forward_exchange_exposed_wait_ms"] = phases.get("exchange_join", {}).get("avg_ms")
                exch_desc += (f"; real-space side pipelined in {nchunks} plane chunks, the peer-storing pass confined to "
                              f"{comm_ctas} SMs on a second stream")
        elif exchange == "dma":
            join = phases.get("exchange_join", {"avg_ms": float("nan"), "launches": 0})
            nvlink = {"bytes_out_per_exchange_per_gpu": a2a_bytes_out, "exchanges_per_step": 2,
                      "exposed_wait_ms_per_exchange": join["avg_ms"], "peak_gbs_per_direction": 900.0,
                      "measured_dma_gbs_per_direction": 779.0,
                      "barrier_avg_ms": phases.get("barrier", {}).get("avg_ms")}
            exch_desc = (f"copy engines over NVLink into peer-mapped buffers (CUDA IPC), pipelined under the kernels in "
                         f"{nchunks} chunks, 2 stream barriers per step")
        else:
            a2a = phases.get("all_to_all", {"avg_ms": float("nan"), "launches": 0})
            nvlink = {"bytes_out_per_exchange_per_gpu": a2a_bytes_out, "exchanges_per_step": 2,
                      "achieved_gbs_per_direction": a2a_bytes_out / (a2a["avg_ms"] * 1e-3) / 1e9 if a2a["launches"] else None,
                      "peak_gbs_per_direction": 900.0}
            exch_desc = "2 NCCL all-to-all per step"
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": B.ch_config(n, world),
            "detail": {"exchange": exch_desc, "chunks": nchunks, "comm_ctas": comm_ctas, "lib_sha16": B.lib_sha16(),
                       "tma_launches": gpfutil.TmaLaunchCount()},
            "parity": parity,
            "clocks": clocks,
            "e2e": {"value": total * e2e_steps / e2e_dt, "unit": METRIC, "h2d_bytes_per_step": 16 * total,
                    "d2h_bytes_per_step": 16 * total, "steps": e2e_steps,
                    "call": "ShardedSolver.Propagate(1) on pinned host slabs, every rank"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": top[0], "achieved": top_bytes / (top[1]["avg_ms"] * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": top_bytes / (top[1]["avg_ms"] * 1e-3) / 1e9 / peak,
                         "traffic": None, "peak_source": peak_src,
                         "step_model": {"bytes_per_cell_update": 192.0, "achieved_per_gpu": value * 192.0 / 1e9 / world,
                                        "frac": value * 192.0 / 1e9 / world / peak},
                         "phases": phases,
                         "nvlink": nvlink},
        }
        print(json.dumps(line), flush=True)
    solver.close()
    tdist.barrier()
    tdist.destroy_process_group()
