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


@pytest.mark.parametrize("dims", [[32, 32], [16, 16, 16], [64, 64, 64]], ids=lambda d: "x".join(map(str, d)))
def test_precipitate_functions_compiled_into_the_forward_pass(dims):
    out = []
    for inpass in (False, True):
        m, conc, phase, s, vol = workloads.build_precipitate(gpf, gpf, gel, dims, expressions=True)
        s.SetJit(True)
        s.SetJitInPass(inpass)
        s.Solve(2, 5)
        assert s.JitKernels() == 4, s.JitLog()  # three functions (pointwise or in-pass) + the k-space update
        out.append((conc.Data.copy(), phase.Data.copy(), s.LPMultiplier(0)))
    assert rel_l2(out[1][0], out[0][0]) <= 1e-13 and rel_l2(out[1][1], out[0][1]) <= 1e-13
    assert abs(out[1][2] - out[0][2]) <= 1e-9
    if dims != [64, 64, 64]:
        g = load(f"precipitate_{'x'.join(map(str, dims))}.npz")
        assert rel_l2(out[1][0], g["conc"]) <= TOL and rel_l2(out[1][1], g["phase"]) <= TOL


# ---- white noise drawn in k-space (Model.SetKSpaceNoise; device code behind -DGOPF_KNOISE) ---------------------
@pytest.mark.skipif(not gpf.HasKSpaceNoise(), reason="libgopfcuda's device code was built without -DGOPF_KNOISE")
@pytest.mark.parametrize("dims", [[64, 64], [32, 32, 32]], ids=lambda d: "x".join(map(str, d)))
def test_kspace_noise_field_and_fused_path(dims):
    """dconc/dt = NOISE from zero with dt = 1 gives the noise field itself: real, N(0, 2 Strength), white
    (the host-compiled generator passes the same checks in tests/test_host_emulation_cpu.py); the pfc model
    with such a noise term keeps one derived field and therefore takes the fused kernels."""
    n = int(np.prod(dims))
    strength = 0.125
    m = gpf.NewModel()
    f = gpf.NewField("conc", n, np.zeros(n, dtype=np.complex128))
    m.AddField(f)
    m.RegisterFunction("NOISE", gpf.WhiteNoise(strength, seed=5).Generate)
    m.AddEquation("dconc/dt = NOISE")
    m.SetKSpaceNoise(True)
    s = gpf.NewSolver(m, dims, 1.0)
    s.Propagate(1)
    x = f.Data.copy()
    var = 2.0 * strength
    assert np.max(np.abs(x.imag)) <= 1e-13 * np.max(np.abs(x.real))
    assert abs(np.mean(x.real)) < 5.0 * np.sqrt(var / n)
    assert abs(np.var(x.real) / var - 1.0) < 5.0 * np.sqrt(2.0 / n)
    grid = x.real.reshape(dims)
    for axis in range(len(dims)):
        assert abs(np.mean(grid * np.roll(grid, 1, axis=axis)) / var) < 5.0 / np.sqrt(n)
    mp, fp, sp = workloads.build_pfc(gpf, gpf, dims, noise=None)
    assert sp.IsFused
    m2 = gpf.NewModel()
    # build_pfc with device noise, k-space variant: still one derived field in use
    import math
    f2 = gpf.NewField("density", n, workloads.pfc_initial(n))
    m2.AddField(f2)
    a = workloads.PFC_LATTICE
    peaks = [gpf.Peak(1.0, 2.0 * math.pi / a, workloads.PFC_PEAK_WIDTH, 4),
             gpf.Peak(1.0 / math.sqrt(2.0), 2.0 * math.pi / (a / math.sqrt(2.0)), workloads.PFC_PEAK_WIDTH, 4)]
    term = gpf.PairCorrlationTerm(gpf.ReciprocalSpacePairCorrelation(workloads.PFC_EFF_TEMP, peaks), "density", 1.0, True)
    ideal = gpf.IdealMixtureTerm(gpf.IdealMix(1.0, 1.0), "density", 1.0, True)
    m2.RegisterImplicitTerm("EXCESS", term, None)
    m2.RegisterMixedTerm("IDEAL", ideal, [ideal.DerivedField(n, m2.Bricks)])
    m2.RegisterFunction("NOISE", gpf.WhiteNoise(workloads.PFC_NOISE_STRENGTH, seed=7).Generate)
    m2.AddEquation("ddensity/dt = IDEAL + EXCESS + NOISE")
    m2.SetKSpaceNoise(True)
    s2 = gpf.NewSolver(m2, dims, workloads.PFC_DT)
    assert s2.IsFused
    s2.Solve(2, 5)
    assert np.all(np.isfinite(f2.Data)) and np.max(np.abs(f2.Data.imag)) <= 1e-12
    # against the noise-free run the difference is of the noise's order, not zero and not large
    sp.Solve(2, 5)
    d = np.linalg.norm(f2.Data - fp.Data) / np.sqrt(n)
    assert 0.0 < d < 50.0 * math.sqrt(2.0 * workloads.PFC_NOISE_STRENGTH) * workloads.PFC_DT * 10
