#!/usr/bin/env python
"""Benchmark of the spectral time-stepping hot path (BASELINE.json metric:
cell-updates/s of the fp64 semi-implicit Cahn-Hilliard step; % of HBM roofline).

    python bench.py --gpus 1 --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU path (oracle port)
    torchrun ... bench.py --gpus N ...                     # slab-sharded 3-D grid, one rank per GPU

Prints ONE JSON line (rank 0).  A "step" is one semi-implicit Euler step of the
whole grid (pf/euler.go:16-47).  `value` is measured with the spectrum resident in
HBM; `e2e` goes through the reference-facing call (Solver.Propagate on host
Field.Data: H2D, step, D2H inside the timed region).
"""
from __future__ import annotations

import argparse
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


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
# capture (profiles/r1c_ncu_full_step.md), keyed by (kernel class, cubic grid edge)
NCU_TRAFFIC = {("fused_kspace", 256): 536.918016e6 + 479.948032e6,
               ("fused_real", 256): 268.504064e6 + 210.432768e6,
               ("pass_inverse_mid", 256): 268.747008e6 + 209.792256e6,
               ("pass_forward_mid", 256): 268.631296e6 + 209.894144e6}


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if "Active" == r[5 + k].strip():
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def oracle_ch_solver(dims, workers):
    """CPU restatement of the same workload (oracle/ -- the checker, used here only as the
    timed CPU baseline)."""
    from oracle import pf as opf
    from oracle import pfutil as opfutil
    from gopf_b200 import synthetic
    n = opfutil.prod_int(dims)
    m = opf.NewModel()
    f = opf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
    m.AddScalar(opf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(opf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    return opf.NewSolver(m, dims, synthetic.CAHN_HILLIARD_DT, workers=workers), n


def cpu_baseline(dims, workers, budget_s):
    """Times the oracle on a bounded sample: as many whole steps of the full grid as fit the
    budget (at least 1)."""
    solver, n = oracle_ch_solver(dims, workers)
    t0 = time.perf_counter()
    solver.Propagate(1)  # first step doubles as warm-up of pocketfft plans
    first = time.perf_counter() - t0
    steps = int(max(1, min(20, budget_s // max(first, 1e-3))))
    t0 = time.perf_counter()
    solver.Propagate(steps)
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": METRIC, "cores": workers, "kind": "port",
            "sample": f"{steps} steps of {'x'.join(map(str, dims))} after 1 warm-up step, scipy.fft workers={workers}",
            "s_per_step": dt / steps}


def run_reference(args):
    """--impl reference: the reference's own CPU path.  Go + FFTW cannot be built in this image
    (no go toolchain, no libfftw3), so this is the oracle port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    # bounded sample: probe at 64^3, then take the largest cubic sample <= the grid whose
    # (steps + warmup) steps fit ~150 s (cells/s is size-normalised)
    probe, _ = oracle_ch_solver([64] * 3, workers)
    probe.Propagate(1)
    t0 = time.perf_counter()
    probe.Propagate(2)
    per_cell = (time.perf_counter() - t0) / 2.0 / 64 ** 3
    total = args.steps + args.warmup
    sample = args.grid
    while per_cell * sample ** 3 * 1.5 * total > 150.0 and sample > 32:
        sample //= 2
    solver, n = oracle_ch_solver([sample] * 3, workers)
    solver.Propagate(args.warmup)
    t0 = time.perf_counter()
    solver.Propagate(args.steps)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cahn-hilliard-3d-{args.grid}^3-semi-implicit-euler", "grid": [args.grid] * 3,
                   "dt": 0.1, "equation": "dconc/dt = LAP conc^3 + m1*LAP conc + m1*gamma*LAP^2 conc"},
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": workers, "kind": "port",
                         "sample": f"each step = one Euler step of {sample}^3 (oracle port of pf/euler.go, "
                                   f"scipy.fft workers={workers}); cells/s is size-normalised"},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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


def run_general_workload(args):
    """cfg 4 / cfg 5 on one GPU: the general (multi-field, catalog-term) path."""
    import torch
    from gopf_b200 import elasticity as gel
    from gopf_b200 import pf as gpf

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference)")
    W = WORKLOADS[args.workload]
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    G = args.grid
    dims = [G, G, G]
    n = G ** 3
    model, solver = build_workload(args.workload, gpf, gpf, gel, dims, pinned=True)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    solver.SetStream(stream.cuda_stream)
    if args.stepper == "rk4":
        solver.SetStepper("rk4")
    elif args.stepper == "implicit_euler":  # configs[1] names this stepper; one step = one Newton-Krylov solve
        solver.Stepper = gpf.ImplicitEuler(solver.Dt)
    # registered functions and the k-space update as NVRTC images (profiles/r1d_jit_notes.md)
    solver.SetJit(not args.no_jit)
    solver.Upload()
    solver.StepDevice(args.warmup)
    torch.cuda.synchronize()
    solver.KernelLaunches(reset=True)
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    solver.StepDevice(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = solver.KernelLaunches(reset=True)
    solver.ProfileBegin()
    solver.StepDevice(args.steps)
    torch.cuda.synchronize()
    prof = solver.ProfileEnd()
    clocks = sampler.stop()
    value = n * args.steps / (ms * 1e-3)
    peak, peak_src = measured_hbm_peak()
    kernels = []
    for k in prof:
        if k["launches"] == 0:
            continue
        avg_ms = k["total_ms"] / k["launches"]
        kernels.append({"kernel": k["kernel"], "launches_per_step": k["launches"] / args.steps, "avg_ms": avg_ms,
                        "algorithmic_bytes": k["bytes_per_launch"], "gbs": k["bytes_per_launch"] / (avg_ms * 1e-3) / 1e9})
    kernels.sort(key=lambda k: -k["avg_ms"] * k["launches_per_step"])
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
    e2e_steps = 3
    solver.Propagate(1)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.Propagate(1)
    e2e_dt = time.perf_counter() - t0
    nf = len(model.Fields)
    e2e = {"value": n * e2e_steps / e2e_dt, "unit": METRIC, "h2d_bytes_per_step": 16 * n * nf, "d2h_bytes_per_step": 16 * n * nf,
           "steps": e2e_steps, "call": "gopf_solver_propagate(s, 1) on pinned host Field.Data"}
    line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": W["name"].format(G=G), "grid": dims, "stepper": args.stepper,
                       "cache": f"arrays of {16 * n / 2**20:.0f} MiB each exceed the 126 MB L2 (no flush needed)",
                       "path": "fused single-field kernels" if solver.IsFused else "general multi-field path",
                       "specialised_kernels": solver.JitKernels()},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline}
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
    from gopf_b200 import pf as gpf
    from gopf_b200 import synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference)")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    G = args.grid
    dims = [G, G, G]
    n = G ** 3

    model = gpf.NewModel()
    conc = gpf.NewField("conc", n, None, pinned=True)
    synthetic.cahn_hilliard_initial(n, 0, out=conc.Data)
    model.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    model.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    model.AddField(conc)
    model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    solver = gpf.NewSolver(model, dims, synthetic.CAHN_HILLIARD_DT, device=dev)
    stream = torch.cuda.Stream(device=dev)  # a real stream: the legacy default stream (handle 0) cannot be handed over
    torch.cuda.set_stream(stream)
    solver.SetStream(stream.cuda_stream)
    assert solver.IsFused, "Cahn-Hilliard must take the fused single-field path"

    # ---- device-resident throughput ("value") ------------------------------------------
    solver.Upload()
    solver.StepDevice(args.warmup)
    torch.cuda.synchronize()
    solver.KernelLaunches(reset=True)
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    solver.StepDevice(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = solver.KernelLaunches(reset=True)
    # ---- per-kernel CUDA events over an identical region (roofline) ------------------------
    solver.ProfileBegin()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    solver.StepDevice(args.steps)
    p1.record(stream)
    torch.cuda.synchronize()
    prof = solver.ProfileEnd()
    prof_ms = p0.elapsed_time(p1)
    clocks = sampler.stop()

    value = n * args.steps / (ms * 1e-3)
    peak, peak_src = measured_hbm_peak()
    kernels = []
    for k in prof:
        if k["launches"] == 0:
            continue
        avg_ms = k["total_ms"] / k["launches"]
        kernels.append({"kernel": k["kernel"], "launches_per_step": k["launches"] / args.steps, "avg_ms": avg_ms,
                        "algorithmic_bytes": k["bytes_per_launch"],
                        "gbs": k["bytes_per_launch"] / (avg_ms * 1e-3) / 1e9})
    kernels.sort(key=lambda k: -k["avg_ms"] * k["launches_per_step"])
    top = kernels[0]
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
                "frac": top["gbs"] / peak, "traffic": NCU_TRAFFIC.get((top["kernel"], G)),
                "traffic_source": "profiles/r1c_ncu_full_step.md" if (top["kernel"], G) in NCU_TRAFFIC else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": top["algorithmic_bytes"], "avg_launch_ms": top["avg_ms"],
                "step_model": {"bytes_per_cell_update": BYTES_PER_CELL_UPDATE_3D,
                               "achieved": value * BYTES_PER_CELL_UPDATE_3D / 1e9,
                               "frac": value * BYTES_PER_CELL_UPDATE_3D / 1e9 / peak},
                "kernels": kernels, "profiled_ms_per_step": prof_ms / args.steps}

    # ---- end to end through Solver.Propagate on host buffers --------------------------------
    e2e_steps = max(3, min(args.steps, 20))
    solver.Propagate(1)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.Propagate(1)  # H2D 16 B/cell + forward FFT + step + inverse FFT + D2H 16 B/cell
    e2e_dt = time.perf_counter() - t0
    e2e = {"value": n * e2e_steps / e2e_dt, "unit": METRIC, "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 16 * n,
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / e2e_steps,
           "call": "gopf_solver_propagate(s, 1) on pinned host Field.Data"}
    # the reference example's epoch of 10 steps between host callbacks (examples/cahnHilliard/main.go:44)
    t0 = time.perf_counter()
    solver.Propagate(10)
    e2e["epoch10_value"] = n * 10 / (time.perf_counter() - t0)

    base = cpu_baseline(dims, 1, args.cpu_budget) if not args.no_cpu_baseline else None
    # The sharded arm (N > 1) runs 1024^3 with the grid fixed (strong scaling); its 1-GPU point is
    # measured here so the 1 -> N efficiency of that workload can be read off the N = 1 line too.
    scaling_base = None
    if args.scaling_base and G != 1024:
        try:
            del solver, model, conc
            torch.cuda.empty_cache()
            scaling_base = single_gpu_throughput(1024, dev, max(3, args.warmup), min(args.steps, 10))
        except Exception as exc:  # e.g. not enough host memory for the 16 GiB pinned field
            scaling_base = {"error": str(exc)[:200]}
    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cahn-hilliard-3d-{G}^3-semi-implicit-euler", "grid": dims, "dt": synthetic.CAHN_HILLIARD_DT,
                   "equation": synthetic.CAHN_HILLIARD_EQUATION, "stepper": "euler",
                   "cache": f"arrays of {16 * n / 2**20:.0f} MiB each exceed the 126 MB L2 (no flush needed)",
                   "path": "fused single-field kernels"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
    }
    if base is not None:
        line["cpu_baseline"] = base
    if scaling_base is not None:
        line["strong_scaling_base"] = scaling_base
    print(json.dumps(line), flush=True)


def single_gpu_throughput(G, dev, warmup, steps):
    """Device-resident throughput of the fused Cahn-Hilliard step at G^3 on one GPU (same seeded
    field as the sharded arm)."""
    import torch
    from gopf_b200 import pf as gpf
    from gopf_b200 import synthetic
    n = G ** 3
    model = gpf.NewModel()
    conc = gpf.NewField("conc", n, None, pinned=True)
    synthetic.cahn_hilliard_initial(n, 0, out=conc.Data)
    model.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    model.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    model.AddField(conc)
    model.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    solver = gpf.NewSolver(model, [G, G, G], synthetic.CAHN_HILLIARD_DT, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    solver.SetStream(stream.cuda_stream)
    solver.Upload()
    solver.StepDevice(warmup)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    solver.StepDevice(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peak, _ = measured_hbm_peak()
    value = n / (ms * 1e-3)
    return {"workload": f"cahn-hilliard-3d-{G}^3-semi-implicit-euler", "n_gpus": 1, "value": value, "unit": METRIC,
            "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "step_model_frac": value * BYTES_PER_CELL_UPDATE_3D / 1e9 / peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gopf_b200", choices=["gopf_b200", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="cubic grid edge (default 256 on 1 GPU, 1024 sharded)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scaling-base", dest="scaling_base", action="store_false",
                    help="N = 1: skip the extra 1024^3 single-GPU measurement (strong-scaling base of the sharded arm)")
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
    ap.add_argument("--comm-ctas", type=int, default=48,
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
            raise SystemExit("bench.py: the sharded path covers the Cahn-Hilliard workload; cfg 4 / cfg 5 run on one GPU")
        if args.grid == 0:
            args.grid = WORKLOADS[args.workload]["default_grid"]
        if args.impl == "reference":
            raise SystemExit("bench.py: --impl reference times the metric's workload (ch)")
        run_general_workload(args)
        return
    if args.grid == 0:
        args.grid = 256 if max(args.gpus, world) == 1 else 1024
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
