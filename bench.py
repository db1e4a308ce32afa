#!/usr/bin/env python
"""Benchmark of the spectral time-stepping hot path (BASELINE.json metric:
cell-updates/s of the fp64 semi-implicit Cahn-Hilliard step; % of HBM roofline).

    python bench.py --gpus 1 --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU path (oracle port)
    torchrun ... bench.py --gpus N ...                     # slab-sharded 3-D grid, one rank per GPU

Prints ONE JSON line (rank 0).  A "step" is one semi-implicit Euler step of the whole grid
(pf/euler.go:16-47).  The workload is the SAME for every N: 3-D Cahn-Hilliard 1024^3
(BASELINE.json configs[2], the grid the metric's 1/2/4/8-GPU figures and the north-star target are
quoted on; it fits one B200: 3 arrays of 16 GiB), so the N = 1 line is the base of a true
strong-scaling curve.  configs[1] (256^3 on one GPU) rides in the N = 1 line as the `cfg2`
sub-record with its own roofline and end-to-end figure, and cfg 2 + SquaredGradient, cfg 4 and
cfg 5 as the `workloads` records.  `value` is measured with the spectrum resident in HBM;
`e2e` goes through the reference-facing call (Solver.Propagate on host Field.Data: H2D, step,
D2H inside the timed region).  Every line carries `parity`: the same CUDA path at 256^3 against
the oracle, run just before the timed region.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cell-updates/s"
BYTES_PER_CELL_UPDATE_3D = 192.0  # SURVEY 8d contract: T_min = 2 transforms x 32*rank bytes
FALLBACK_HBM_GBS = 6650.0         # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
DEFAULT_GRID = 1024               # every N (strong scaling); --grid 256 runs configs[1] as the headline
PARITY_GRID, PARITY_STEPS = 256, 3


def ch_config(G, n_gpus=1):
    """`config` of the metric's workload -- identical in the CUDA arm and the reference arm."""
    from gopf_b200 import synthetic
    return {"workload": f"cahn-hilliard-3d-{G}^3-semi-implicit-euler", "grid": [G, G, G], "dt": synthetic.CAHN_HILLIARD_DT,
            "equation": synthetic.CAHN_HILLIARD_EQUATION, "stepper": "euler",
            "parallelism": "single-gpu" if n_gpus <= 1 else f"slab{n_gpus}",
            "cache": f"per-GPU arrays of {16 * G ** 3 / max(1, n_gpus) / 2**20:.0f} MiB each exceed the 126 MB L2 (no flush needed)"}


def lib_sha16():
    try:
        h = hashlib.sha256()
        with open(os.path.join(ROOT, "gopf_b200", "lib", "libgopfcuda.so"), "rb") as f:
            for blk in iter(lambda: f.read(1 << 20), b""):
                h.update(blk)
        return h.hexdigest()[:16]
    except Exception:
        return None


def ncu_traffic(kernel, G):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at G^3 from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written by scripts/summarize_ncu.py --traffic)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tab = json.load(f)
    except Exception:
        return None, None
    for e in tab.get("entries", []):
        if e.get("kernel") == kernel and int(e.get("grid", 0)) == G:
            return float(e["dram_bytes"]), e.get("source")
    return None, None


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for k, nm in enumerate(names):
                    if "Active" == r[5 + k].strip():
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_min_mhz=float(min(sm)), sm_max_mhz=float(max(mx)),
                       power_w_max=float(max(pw)) if pw else None, reasons=sorted(reasons), samples=len(sm))
        return out


# ---- the metric's model -------------------------------------------------------------------------
def build_ch(mod, G, *, pinned=False, cells=None, offset=0, workers=None):
    """3-D Cahn-Hilliard (examples/cahnHilliard/main.go:12-44 on the synthetic seeded field) on module
    `mod` (gopf_b200.pf or, for the CPU legs, the oracle).  Returns (model, field, solver)."""
    from gopf_b200 import synthetic
    n = G ** 3 if cells is None else cells
    m = mod.NewModel()
    if pinned:
        f = mod.NewField("conc", n, None, pinned=True)
        synthetic.cahn_hilliard_initial(n, 0, offset=offset, out=f.Data)
    else:
        f = mod.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0, offset=offset))
    m.AddScalar(mod.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(mod.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    if cells is not None:
        return m, f, None
    if workers is not None:
        return m, f, mod.NewSolver(m, [G, G, G], synthetic.CAHN_HILLIARD_DT, workers=workers)
    return m, f, mod.NewSolver(m, [G, G, G], synthetic.CAHN_HILLIARD_DT)


def oracle_reference_field(G, steps, workers=None):
    """The oracle's field after `steps` steps at G^3 (the checker; used by `parity` only)."""
    from oracle import pf as opf
    _, f, s = build_ch(opf, G, workers=workers or (os.cpu_count() or 1))
    s.Propagate(steps)
    return f.Data


def parity_single_gpu(dev):
    """256^3, 3 steps, this repo's fused CUDA path against the oracle (pf/euler.go:16-47)."""
    from gopf_b200 import pf as gpf
    t0 = time.perf_counter()
    _, f, s = build_ch(gpf, PARITY_GRID)
    s.Upload()
    s.StepDevice(PARITY_STEPS)
    s.Download()
    got = f.Data.copy()
    s.close()
    ref = oracle_reference_field(PARITY_GRID, PARITY_STEPS)
    err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    return {"rel_l2": err, "grid": PARITY_GRID, "steps": PARITY_STEPS, "tolerance": 1e-10, "ok": bool(err <= 1e-10),
            "against": "oracle restatement of pf/euler.go (scipy.fft), same seeded field", "n_gpus": 1,
            "seconds": round(time.perf_counter() - t0, 1)}


def parity_sharded(world, rank, local, exchange, nchunks, comm_ctas):
    """256^3, 3 steps, the slab-sharded CUDA path on `world` GPUs against the oracle (rank 0 compares)."""
    import torch
    import torch.distributed as tdist
    from gopf_b200 import dist as gdist
    from gopf_b200 import pf as gpf
    from gopf_b200 import synthetic
    t0 = time.perf_counter()
    n = PARITY_GRID
    cells = n ** 3 // world
    model, f, _ = build_ch(gpf, n, cells=cells, offset=rank * cells)
    s = gdist.ShardedSolver(model, n, synthetic.CAHN_HILLIARD_DT, device=local, exchange=exchange,
                            nchunks=min(nchunks, max(1, n // world // 4)), comm_ctas=comm_ctas)
    s.Upload()
    s.StepDevice(PARITY_STEPS)
    s.Download()
    mine = torch.view_as_real(torch.from_numpy(f.Data).to(torch.device("cuda", local))).reshape(-1)
    full = torch.empty(world * mine.numel(), dtype=mine.dtype, device=mine.device)
    tdist.all_gather_into_tensor(full, mine)
    used = s.exchange
    s.close()
    rec = None
    if rank == 0:
        got = torch.view_as_complex(full.reshape(-1, 2)).cpu().numpy()
        ref = oracle_reference_field(n, PARITY_STEPS)
        err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        rec = {"rel_l2": err, "grid": n, "steps": PARITY_STEPS, "tolerance": 1e-10, "ok": bool(err <= 1e-10),
               "against": "oracle restatement of pf/euler.go (scipy.fft), same seeded field", "n_gpus": world,
               "exchange": used, "seconds": round(time.perf_counter() - t0, 1)}
    del full, mine
    torch.cuda.empty_cache()
    tdist.barrier()
    return rec


def cpu_baseline(workers, budget_s):
    """Times the oracle on a bounded sample of the metric's workload: whole Euler steps of a 256^3
    grid (cells/s is size-normalised; a 1024^3 step is ~3 min on one core)."""
    from oracle import pf as opf
    G = 256
    _, _, solver = build_ch(opf, G, workers=workers)
    t0 = time.perf_counter()
    solver.Propagate(1)  # first step doubles as warm-up of pocketfft plans
    first = time.perf_counter() - t0
    steps = int(max(1, min(20, budget_s // max(first, 1e-3))))
    t0 = time.perf_counter()
    solver.Propagate(steps)
    dt = time.perf_counter() - t0
    return {"value": G ** 3 * steps / dt, "unit": METRIC, "cores": workers, "kind": "port",
            "sample": f"{steps} Euler steps of a {G}^3 grid after 1 warm-up step (oracle port of pf/euler.go, scipy.fft "
                      f"workers={workers}); the Go reference is single-goroutine with unthreaded FFTW",
            "s_per_step": dt / steps}


def run_reference(args):
    """--impl reference: the reference's own CPU path.  Go + FFTW cannot be built in this image
    (no go toolchain, no libfftw3), so this is the oracle port with every host thread.  `config` is
    the CUDA arm's; each timed step is one Euler step of a bounded SAMPLE of that workload (a 256^3
    grid unless --grid is smaller), stated in `sample` and `cpu_baseline.sample`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pf as opf
    workers = os.cpu_count() or 1
    sample = min(args.grid, 256)
    _, _, solver = build_ch(opf, sample, workers=workers)
    solver.Propagate(args.warmup)
    t0 = time.perf_counter()
    solver.Propagate(args.steps)
    dt = time.perf_counter() - t0
    value = sample ** 3 * args.steps / dt
    desc = (f"each of the {args.steps} timed steps (after {args.warmup} warm-up steps) = one Euler step of a {sample}^3 grid "
            f"(oracle port of pf/euler.go, scipy.fft workers={workers}); cells/s is size-normalised"
            + ("" if sample == args.grid else f"; the full {args.grid}^3 grid is ~{(args.grid / sample) ** 3:.0f}x the cells per step"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": ch_config(args.grid, args.gpus),
        "sample": {"grid": [sample] * 3, "is_full_grid": sample == args.grid, "what": desc},
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": workers, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- one-GPU measurement of the metric's workload ----------------------------------------------------
def kernel_table(prof, steps):
    kernels = []
    for k in prof:
        if k["launches"] == 0:
            continue
        avg_ms = k["total_ms"] / k["launches"]
        kernels.append({"kernel": k["kernel"], "launches_per_step": k["launches"] / steps, "avg_ms": avg_ms,
                        "algorithmic_bytes": k["bytes_per_launch"],
                        "gbs": k["bytes_per_launch"] / (avg_ms * 1e-3) / 1e9})
    kernels.sort(key=lambda k: -k["avg_ms"] * k["launches_per_step"])
    return kernels


def measure_ch(G, dev, warmup, steps, blocks=5, e2e=True):
    """Fused Cahn-Hilliard step at G^3 on one GPU: `blocks` timed blocks of exactly `steps` steps each
    (CUDA events on the solver's stream; the median block is the value), per-kernel events over one
    more identical block, then the end-to-end call on pinned host memory."""
    import torch
    from gopf_b200 import pf as gpf

    n = G ** 3
    model, conc, solver = build_ch(gpf, G, pinned=True)
    stream = torch.cuda.Stream(device=dev)  # a real stream: the legacy default stream (handle 0) cannot be handed over
    torch.cuda.set_stream(stream)
    solver.SetStream(stream.cuda_stream)
    assert solver.IsFused, "Cahn-Hilliard must take the fused single-field path"
    solver.Upload()
    solver.StepDevice(warmup)
    torch.cuda.synchronize()
    solver.KernelLaunches(reset=True)
    from gopf_b200 import pfutil as gpfutil
    gpfutil.TmaLaunchCount(reset=True)
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.25)
    block_ms = []
    for _ in range(blocks):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        solver.StepDevice(steps)
        e1.record(stream)
        torch.cuda.synchronize()
        block_ms.append(e0.elapsed_time(e1))
    launches = solver.KernelLaunches(reset=True) // blocks
    tma_launches = gpfutil.TmaLaunchCount(reset=True) // blocks
    solver.ProfileBegin()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    solver.StepDevice(steps)
    p1.record(stream)
    torch.cuda.synchronize()
    prof = solver.ProfileEnd()
    prof_ms = p0.elapsed_time(p1)
    clocks = sampler.stop()

    ms = float(np.median(block_ms))
    value = n * steps / (ms * 1e-3)
    peak, peak_src = measured_hbm_peak()
    kernels = kernel_table(prof, steps)
    top = kernels[0]
    traffic, traffic_src = ncu_traffic(top["kernel"], G)
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
                "frac": top["gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": top["algorithmic_bytes"], "avg_launch_ms": top["avg_ms"],
                "step_model": {"bytes_per_cell_update": BYTES_PER_CELL_UPDATE_3D,
                               "achieved": value * BYTES_PER_CELL_UPDATE_3D / 1e9,
                               "frac": value * BYTES_PER_CELL_UPDATE_3D / 1e9 / peak,
                               "bytes_per_cell_update_moved_by_this_path":
                                   sum(k["algorithmic_bytes"] * k["launches_per_step"] for k in kernels) / n},
                "kernels": kernels, "profiled_ms_per_step": prof_ms / steps}
    rec = {"value": value, "ms_per_step": ms / steps, "timed_blocks": {"blocks": blocks, "steps_per_block": steps,
                                                                        "block_ms": [round(b, 4) for b in block_ms],
                                                                        "spread": (max(block_ms) - min(block_ms)) / ms,
                                                                        "value_from": "median block"},
           "clocks": clocks, "gpu_launches": launches, "tma_launches": tma_launches, "roofline": roofline}
    if e2e:
        # end to end through Solver.Propagate on host buffers: H2D 16 B/cell + forward FFT + step + inverse FFT + D2H
        e2e_steps = 3 if G >= 1024 else max(3, min(steps, 20))
        solver.Propagate(1)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            solver.Propagate(1)
        e2e_dt = time.perf_counter() - t0
        rec["e2e"] = {"value": n * e2e_steps / e2e_dt, "unit": METRIC, "h2d_bytes_per_step": 16 * n,
                      "d2h_bytes_per_step": 16 * n, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / e2e_steps,
                      "call": "gopf_solver_propagate(s, 1) on pinned host Field.Data"}
        # the reference example's epoch of 10 steps between host callbacks (examples/cahnHilliard/main.go:44)
        t0 = time.perf_counter()
        solver.Propagate(10)
        rec["e2e"]["epoch10_value"] = n * 10 / (time.perf_counter() - t0)
    solver.close()
    del solver, model, conc
    torch.cuda.empty_cache()
    return rec


# SURVEY.md 8d contract bytes per cell-update (T_min x 96 B in 3-D) and the reference's own
# transform count for context
WORKLOADS = {
    "precipitate": {"contract_bytes": 576.0, "ref_bytes": 2016.0, "default_grid": 512,
                    "name": "strain-single-precipitate-3d-{G}^3-khachaturyan-elasticity-volume-constraint"},
    "pfc": {"contract_bytes": 192.0, "ref_bytes": 480.0, "default_grid": 512,
            "name": "pfc-3d-{G}^3-pair-correlation-white-noise-vandeven"},
    # SURVEY.md 8d, cfg 2 variant: + SquaredGradient as an explicit term: T_min = 2 + 4 = 6 transforms in 3-D
    # (reference 3 + 6 = 9)
    "ch_sqgrad": {"contract_bytes": 576.0, "ref_bytes": 864.0, "default_grid": 256,
                  "name": "cahn-hilliard-3d-{G}^3-semi-implicit-euler-plus-squared-gradient"},
}


def build_workload(kind, pf, terms, elasticity, dims, pinned, device_noise=True):
    from gopf_b200 import workloads
    if kind == "ch_sqgrad":
        from gopf_b200 import synthetic
        n = int(np.prod(dims))
        m = pf.NewModel()
        if pinned:
            f = pf.NewField("conc", n, None, pinned=True)
            f.Data[:] = 0.1 * synthetic.cahn_hilliard_initial(n, 5)
        else:
            f = pf.NewField("conc", n, 0.1 * synthetic.cahn_hilliard_initial(n, 5))
        m.AddScalar(pf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
        m.AddScalar(pf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
        m.AddField(f)
        m.RegisterExplicitTerm("GRAD_SQ", terms.NewSquareGradient("conc", dims), None)
        m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION + " + GRAD_SQ")
        return m, pf.NewSolver(m, dims, 0.05)
    if kind == "precipitate":
        m, conc, phase, solver, _ = workloads.build_precipitate(pf, terms, elasticity, dims, expressions=pf.__name__.startswith("gopf_b200"),
                                                               pinned=pinned)
        return m, solver
    if device_noise:
        # with the k-space generator compiled in, the noise costs no transform and the model stays on the fused kernels
        m, f, solver = workloads.build_pfc(pf, terms, dims, noise="device", pinned=pinned, kspace_noise=pf.HasKSpaceNoise())
    else:
        from oracle import terms as oterms
        m, f, solver = workloads.build_pfc(pf, terms, dims, noise=oterms.WhiteNoise(workloads.PFC_NOISE_STRENGTH).Generate)
    return m, solver


def measure_workload(kind, G, dev, warmup, steps, stepper="euler", jit=True, e2e_steps=0):
    """cfg 4 / cfg 5 / cfg 2 + SquaredGradient on one GPU.  Returns the record (value, per-kernel
    events, fraction of the workload's own contract roofline)."""
    import torch
    from gopf_b200 import elasticity as gel
    from gopf_b200 import pf as gpf

    W = WORKLOADS[kind]
    dims = [G, G, G]
    n = G ** 3
    model, solver = build_workload(kind, gpf, gpf, gel, dims, pinned=True)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    solver.SetStream(stream.cuda_stream)
    if stepper == "rk4":
        solver.SetStepper("rk4")
    elif stepper == "implicit_euler":  # configs[1] names this stepper; one step = one Newton-Krylov solve
        solver.Stepper = gpf.ImplicitEuler(solver.Dt)
    solver.SetJit(jit)
    solver.Upload()
    solver.StepDevice(warmup)
    torch.cuda.synchronize()
    solver.KernelLaunches(reset=True)
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    solver.StepDevice(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = solver.KernelLaunches(reset=True)
    solver.ProfileBegin()
    solver.StepDevice(steps)
    torch.cuda.synchronize()
    prof = solver.ProfileEnd()
    clocks = sampler.stop()
    value = n * steps / (ms * 1e-3)
    peak, peak_src = measured_hbm_peak()
    kernels = kernel_table(prof, steps)
    top = kernels[0]
    moved = sum(k["algorithmic_bytes"] * k["launches_per_step"] for k in kernels) / n
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
                "frac": top["gbs"] / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": top["algorithmic_bytes"], "avg_launch_ms": top["avg_ms"],
                "step_model": {"bytes_per_cell_update": W["contract_bytes"], "achieved": value * W["contract_bytes"] / 1e9,
                               "frac": value * W["contract_bytes"] / 1e9 / peak,
                               "bytes_per_cell_update_moved_by_this_path": moved,
                               "bytes_per_cell_update_reference_structure": W["ref_bytes"]},
                "kernels": kernels}
    rec = {"workload": W["name"].format(G=G), "grid": dims, "stepper": stepper, "value": value, "unit": METRIC,
           "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
           "path": "fused single-field kernels" if solver.IsFused else "general multi-field path",
           "specialised_kernels": solver.JitKernels(), "clocks": clocks, "gpu_launches": launches, "roofline": roofline}
    if e2e_steps > 0:
        solver.Propagate(1)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            solver.Propagate(1)
        e2e_dt = time.perf_counter() - t0
        nf = len(model.Fields)
        rec["e2e"] = {"value": n * e2e_steps / e2e_dt, "unit": METRIC, "h2d_bytes_per_step": 16 * n * nf,
                      "d2h_bytes_per_step": 16 * n * nf, "steps": e2e_steps,
                      "call": "gopf_solver_propagate(s, 1) on pinned host Field.Data"}
    solver.close()
    del solver, model
    torch.cuda.empty_cache()
    return rec


def run_general_workload(args):
    """--workload precipitate | pfc | ch_sqgrad as the line's own workload (one GPU)."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference)")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    G = args.grid
    rec = measure_workload(args.workload, G, dev, args.warmup, args.steps, stepper=args.stepper, jit=not args.no_jit, e2e_steps=3)
    n = G ** 3
    line = {"metric": METRIC, "value": rec["value"], "unit": METRIC, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": rec["workload"], "grid": rec["grid"], "stepper": args.stepper,
                       "cache": f"arrays of {16 * n / 2**20:.0f} MiB each exceed the 126 MB L2 (no flush needed)",
                       "path": rec["path"], "specialised_kernels": rec["specialised_kernels"]},
            "clocks": rec["clocks"], "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"]}
    if not args.no_cpu_baseline:
        # oracle on a bounded sample: the same model at 64^3 (cells/s is size-normalised), single thread
        from oracle import elasticity as oel
        from oracle import pf as opf
        from oracle import terms as oterms
        sg = 64
        om, osolver = build_workload(args.workload, opf, oterms, oel, [sg] * 3, pinned=False, device_noise=False)
        osolver.Propagate(1)
        t0 = time.perf_counter()
        osolver.Propagate(3)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sg ** 3 * 3 / dt, "unit": METRIC, "cores": 1, "kind": "port",
                                "sample": f"3 steps of the same model at {sg}^3 after 1 warm-up step, oracle port, scipy.fft workers=1"}
    print(json.dumps(line), flush=True)


def run_single_gpu(args):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference)")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    G = args.grid
    parity = parity_single_gpu(dev) if not args.no_parity else None
    main_rec = measure_ch(G, dev, args.warmup, args.steps, blocks=args.blocks)
    line = {
        "metric": METRIC, "value": main_rec["value"], "unit": METRIC, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main_rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": ch_config(G),
        "detail": {"path": "fused single-field kernels", "lib_sha16": lib_sha16(),
                   "copy_engine_kernel_launches_per_block": main_rec["tma_launches"],
                   "scaling_note": "the same 1024^3 grid runs at every N (strong scaling); cfg2 = BASELINE.json configs[1]"},
        "timed_blocks": main_rec["timed_blocks"], "clocks": main_rec["clocks"], "e2e": main_rec["e2e"],
        "gpu_launches": main_rec["gpu_launches"], "roofline": main_rec["roofline"],
    }
    if parity is not None:
        line["parity"] = parity
    if G != 256 and not args.no_cfg2:
        # BASELINE.json configs[1]: 256^3 on one B200, with its own roofline and end-to-end figure
        try:
            c2 = measure_ch(256, dev, max(args.warmup, 5), max(args.steps, 50), blocks=args.blocks)
            c2.update({"config": ch_config(256), "unit": METRIC, "steps": max(args.steps, 50), "warmup": max(args.warmup, 5)})
            line["cfg2"] = c2
        except Exception as exc:
            line["cfg2"] = {"error": str(exc)[:300]}
    if not args.no_workloads:
        # cfg 2 + SquaredGradient, cfg 4, cfg 5 at their stated sizes: a few steps each, per-kernel events,
        # fraction of each workload's own contract roofline
        recs = []
        for kind in ("ch_sqgrad", "precipitate", "pfc"):
            try:
                recs.append(measure_workload(kind, WORKLOADS[kind]["default_grid"], dev, 3, 10))
            except Exception as exc:
                recs.append({"workload": kind, "error": str(exc)[:300]})
        # BASELINE.json configs[1] as worded: ImplicitEuler (Newton-Krylov on the device) + SquareGradientTerm at
        # 256^3.  One step = one non-linear solve = many right-hand-side evaluations, so the contract fraction of
        # the Euler step does not apply: the record is the time per step and the launches behind it.
        try:
            r = measure_workload("ch_sqgrad", WORKLOADS["ch_sqgrad"]["default_grid"], dev, 3, 2, stepper="implicit_euler")
            r["workload"] += " (stepper: ImplicitEuler, DefaultNonLinSolver settings)"
            r["roofline"]["step_model"]["note"] = "fraction of the Euler step's contract; not meaningful for a Newton-Krylov solve"
            recs.append(r)
        except Exception as exc:
            recs.append({"workload": "ch_sqgrad/implicit_euler", "error": str(exc)[:300]})
        line["workloads"] = recs
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(1, args.cpu_budget)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gopf_b200", choices=["gopf_b200", "reference"])
    ap.add_argument("--grid", type=int, default=0, help=f"cubic grid edge (default {DEFAULT_GRID} at every N)")
    ap.add_argument("--blocks", type=int, default=5, help="N = 1: timed blocks of --steps steps each; the median block is the value")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg2", action="store_true", help="N = 1: skip the 256^3 (configs[1]) sub-record")
    ap.add_argument("--no-workloads", action="store_true", help="N = 1: skip the cfg 2+SquaredGradient / cfg 4 / cfg 5 records")
    ap.add_argument("--no-parity", action="store_true", help="skip the 256^3 check against the oracle before the timed region")
    ap.add_argument("--stepper", default="euler", choices=["euler", "rk4", "implicit_euler"],
                    help="general workloads: time stepper (the default 3-D Cahn-Hilliard arm is semi-implicit Euler)")
    ap.add_argument("--no-jit", action="store_true",
                    help="general workloads: interpreter kernels instead of the NVRTC-specialised ones")
    ap.add_argument("--workload", default="ch", choices=["ch", "precipitate", "pfc", "ch_sqgrad"],
                    help="ch: Cahn-Hilliard (BASELINE.json configs 1-3, the metric's workload); precipitate: cfg 4; pfc: cfg 5")
    ap.add_argument("--exchange", default="peer", choices=["peer", "dma", "nccl"],
                    help="sharded runs: peer stores fused into the passes, copy-engine copies pipelined under the "
                         "kernels, or NCCL all-to-all")
    ap.add_argument("--chunks", type=int, default=8, help="sharded runs: plane / column chunks the exchange is pipelined in")
    ap.add_argument("--comm-ctas", type=int, default=0,
                    help="--exchange peer: SMs given to the NVLink-bound peer-storing pass while it overlaps the next chunk")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "ch" and args.stepper != "euler":
        raise SystemExit("bench.py: the 3-D Cahn-Hilliard arm is the semi-implicit Euler step of the metric; "
                         "--stepper applies to --workload precipitate | pfc | ch_sqgrad")
    if args.workload != "ch":
        if max(args.gpus, world) > 1:
            raise SystemExit("bench.py: --workload precipitate | pfc | ch_sqgrad are one-GPU lines")
        if args.grid == 0:
            args.grid = WORKLOADS[args.workload]["default_grid"]
        if args.impl == "reference":
            raise SystemExit("bench.py: --impl reference times the metric's workload (ch)")
        run_general_workload(args)
        return
    if args.grid == 0:
        args.grid = DEFAULT_GRID
    if args.impl == "reference":
        run_reference(args)
        return
    if max(args.gpus, world) > 1:
        from gopf_b200 import dist_bench
        dist_bench.run(args)
        return
    run_single_gpu(args)


if __name__ == "__main__":
    main()
