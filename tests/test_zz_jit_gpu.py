"""Run-time specialisation (gopf_b200/csrc/jit.h) on the device: registered functions and the k-space
update compiled by NVRTC must reproduce the committed golden vectors (tests/golden/, generated from
the pinned oracle) at the same tolerance as the interpreter kernels, and must actually be the kernels
that ran (a silent fall-back to the interpreter fails the test).  No oracle code runs here.

The specialisation is opt-in in this round (GOPF_JIT=1 / Solver.SetJit), so this file is named to run
after the tests of the default kernels."""
import os

import numpy as np
import pytest

from gopf_b200 import elasticity as gel
from gopf_b200 import pf as gpf
from gopf_b200 import workloads

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("dims", [[32, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_precipitate_specialised(dims):
    g = load(f"precipitate_{'x'.join(map(str, dims))}.npz")
    m, conc, phase, s, vol = workloads.build_precipitate(gpf, gpf, gel, dims, expressions=True)
    s.SetJit(True)
    s.Solve(2, 5)
    assert s.JitKernels() == 4, s.JitLog()  # three registered functions + the k-space update
    assert rel_l2(conc.Data, g["conc"]) <= TOL and rel_l2(phase.Data, g["phase"]) <= TOL
    assert abs(s.LPMultiplier(0) - float(g["multiplier"][0])) <= 1e-9


def test_pfc_specialised():
    g = load("pfc_32x32_vandeven5.npz")
    m, f, s = workloads.build_pfc(gpf, gpf, [32, 32], noise=None, filt_order=5)
    s.SetJit(True)
    s.ForceGeneric(True)  # one field + one derived field would otherwise take the fused kernels, which have no use for it
    s.Solve(2, 5)
    assert s.JitKernels() == 2, s.JitLog()  # ideal-mixture polynomial + the k-space update (tabulated implicit side)
    assert rel_l2(f.Data, g["density"]) <= TOL


@pytest.mark.parametrize("dims", [[16, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_point_sources_specialised(dims):
    g = load(f"sources_{'x'.join(map(str, dims))}.npz")
    m, f, s = workloads.build_sourced_diffusion(gpf, gpf, dims)
    s.SetJit(True)
    s.Solve(2, 5)
    assert s.JitKernels() == 1, s.JitLog()  # the k-space update (the nonlinearity is a monomial, evaluated in-pass)
    assert rel_l2(f.Data, g["conc"]) <= TOL


def test_specialised_and_interpreted_kernels_agree_step_by_step():
    dims = [32, 32, 32]
    out = []
    for jit in (False, True):
        m, conc, phase, s, vol = workloads.build_precipitate(gpf, gpf, gel, dims, expressions=True)
        s.SetJit(jit)
        s.Solve(3, 4)
        assert s.JitKernels() == (4 if jit else 0), s.JitLog()
        out.append((conc.Data.copy(), phase.Data.copy(), s.LPMultiplier(0)))
    # same operations up to the compiler's choice of fused multiply-adds
    assert rel_l2(out[1][0], out[0][0]) <= 1e-13 and rel_l2(out[1][1], out[0][1]) <= 1e-13
    assert abs(out[1][2] - out[0][2]) <= 1e-9
