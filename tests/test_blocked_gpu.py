"""Blocked k-space layout of the fused Cahn-Hilliard path ([n0/2^s][n1][2^s][n2], csrc/solver.h).

The layout only changes WHERE cells sit between the middle-axis passes and the k-space kernel: every line runs
through the same Stockham engine on the same values, so a run with the layout on must be BITWISE the run with
it off, for the register-resident kernels and for the copy-engine ones, and -- like every fused run -- within
1e-10 of the oracle (pf/euler.go:16-47).  Host synchronisations in the middle of a run convert back and forth."""
import os

import numpy as np
import pytest

from gopf_b200 import pf as gpf
from gopf_b200 import pfutil as gpfutil
from gopf_b200 import synthetic
from oracle import pf as opf

pytestmark = pytest.mark.gpu

KEYS = ("GOPF_BLOCKED", "GOPF_BLOCK_LOG", "GOPF_TMA", "GOPF_TMA_MIN_N", "GOPF_TMA_KSPACE", "GOPF_TMA_PASS", "GOPF_TMA_REAL",
        "GOPF_REAL_PAIRS")


@pytest.fixture()
def env():
    saved = {k: os.environ.get(k) for k in KEYS}
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _ch(dims, steps, sync_every=0, expect_blocked=None):
    n = int(np.prod(dims))
    m = gpf.NewModel()
    f = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
    m.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    s = gpf.NewSolver(m, dims, synthetic.CAHN_HILLIARD_DT)
    assert s.IsFused
    s.Upload()
    done = 0
    while done < steps:
        k = min(sync_every or steps, steps - done)
        s.StepDevice(k)
        done += k
        if expect_blocked is not None:
            log, active = s.BlockedLayout()
            assert (log > 0) == expect_blocked and active == expect_blocked
        if sync_every and done < steps:
            s.Download()  # converts the spectrum back; the next step re-enters the blocked layout
            assert s.BlockedLayout()[1] is False
    s.Download()
    out = f.Data.copy()
    s.close()
    return out


def _oracle(dims, steps):
    n = int(np.prod(dims))
    om = opf.NewModel()
    of = opf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
    om.AddScalar(opf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    om.AddScalar(opf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    om.AddField(of)
    om.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    opf.NewSolver(om, dims, synthetic.CAHN_HILLIARD_DT, workers=os.cpu_count() or 1).Propagate(steps)
    return of.Data


# (the fused 3-D path takes cubic grids only: the reference's Freq is axis-consistent only there, fft_plan.cu)
@pytest.mark.parametrize("dims,log", [([32, 32, 32], 2), ([64, 64, 64], 4), ([16, 16, 16], 3), ([128, 128, 128], 7)],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else f"s{v}")
def test_register_kernels_blocked_bitwise_and_vs_oracle(env, dims, log):
    env["GOPF_BLOCKED"] = "0"
    ref = _ch(dims, 10, expect_blocked=False)
    env["GOPF_BLOCKED"] = "1"
    env["GOPF_BLOCK_LOG"] = str(log)
    got = _ch(dims, 10, sync_every=4, expect_blocked=True)
    assert np.array_equal(got, ref)
    want = _oracle(dims, 10)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-10


@pytest.mark.parametrize("kspace", ["1", "0"], ids=["kspace-tma", "kspace-reg"])
def test_copy_engine_kernels_blocked_bitwise_512(env, kspace):
    dims, steps = [512, 512, 512], 3
    env["GOPF_TMA_MIN_N"] = "512"
    env["GOPF_REAL_PAIRS"] = "0"  # the paired real-space kernel equals the others to rounding only (test_tma_gpu.py)
    env["GOPF_TMA"] = "0"
    env["GOPF_BLOCKED"] = "0"
    ref = _ch(dims, steps)
    env["GOPF_TMA"] = "1"
    env["GOPF_TMA_KSPACE"] = kspace
    env["GOPF_BLOCKED"] = "1"
    gpfutil.TmaLaunchCount(reset=True)
    got = _ch(dims, steps, sync_every=2, expect_blocked=True)
    assert gpfutil.TmaLaunchCount() > 0
    assert np.array_equal(got, ref)


def test_benchmark_grid_1024_cubed_blocked_by_default_bitwise(env):
    """The 1024^3 bench grid takes the blocked layout by default; two steps against the row-major run."""
    dims, steps = [1024, 1024, 1024], 2
    for k in KEYS:
        env.pop(k, None)
    got = _ch(dims, steps, expect_blocked=True)
    fp_got = (float(got.real.sum()), got[::4099].copy())
    del got
    env["GOPF_BLOCKED"] = "0"
    ref = _ch(dims, steps, expect_blocked=False)
    assert fp_got[0] == float(ref.real.sum())
    assert np.array_equal(fp_got[1], ref[::4099])
