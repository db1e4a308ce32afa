"""GPU parity of the catalog terms evaluated outside the k-space program:
ChargeTransport (pf/chargeTransport.go) and point Sources (pf/sourceTerm.go, pf/model.go:291-294),
through the C ABI against the oracle, plus the reference's own known-answer tests restated on the
device path.  Tolerance: relative L2 <= 1e-10 (BASELINE.json north_star)."""
import math

import numpy as np
import pytest

from gopf_b200 import pf as gpf
from oracle import pf as opf
from oracle import pfutil as opfutil
from oracle import terms as oterms

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _charge_models(dims, density, sigma, ext):
    """One density field with d rho/dt = MINUS_DIV_CURRENT on the device mirror and on the oracle."""
    n = opfutil.prod_int(dims)
    out = []
    for mod in (gpf, opf):
        m = mod.NewModel()
        f = mod.NewField("density", n, density.astype(np.complex128).copy())
        m.AddField(f)
        if mod is gpf:
            term = gpf.ChargeTransport(sigma, ext, "density")
        else:
            term = oterms.ChargeTransport(lambda i: sigma[i], ext, "density", opfutil.NewFFTW(dims))
        m.RegisterExplicitTerm("MINUS_DIV_CURRENT", term, None)
        m.AddEquation("ddensity/dt = MINUS_DIV_CURRENT")
        out.append((m, f, term))
    return out


@pytest.mark.parametrize("case", [0, 1])
def test_charge_transport_reference_known_answers(case):
    # pf/chargeTransport_test.go:10-124 on the device: isotropic conductivity, E_ext = (1, 0);
    # case 0 zero density, case 1 rho = sin(2 pi x) with -div j = -rho
    N = 32
    dims = [N, N]
    n = N * N
    i = np.arange(n)
    x = (i // N) / float(N)  # pfutil.Pos(...)[0] is the row
    density = np.zeros(n) if case == 0 else np.sin(2.0 * math.pi * x)
    sigma = np.broadcast_to(np.array([1.0, 1.0, 0.0]), (n, 3))
    (gm, gf, gterm), _ = _charge_models(dims, density, sigma, [1.0, 0.0])
    dt = 0.5
    s = gpf.NewSolver(gm, dims, dt)
    s.Upload()
    current = gterm.Current()
    expect_x = np.ones(n) if case == 0 else 1.0 - float(N) * np.cos(2.0 * math.pi * x) / (2.0 * math.pi)
    assert np.max(np.abs(current[0] - expect_x)) < 1e-10
    assert np.max(np.abs(current[1])) < 1e-10
    s.StepDevice(1)  # rho_1 = rho_0 + dt * (-div j)  (no implicit part)
    s.Download()
    div = (gf.Data - density) / dt
    expect = np.zeros(n) if case == 0 else -np.sin(2.0 * math.pi * x)
    assert np.max(np.abs(div.real - expect)) < 1e-10
    assert np.max(np.abs(div.imag)) < 1e-10


@pytest.mark.parametrize("dims", [[16, 32], [32, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("stepper", ["euler", "rk4"])
def test_charge_transport_vs_oracle(dims, stepper):
    # anisotropic, spatially varying conductivity (examples/electricConductivity: rotated grains)
    n = opfutil.prod_int(dims)
    rank = len(dims)
    rng = np.random.default_rng(7)
    density = 0.1 * rng.standard_normal(n)
    nv = 3 if rank == 2 else 6
    sigma = np.zeros((n, nv))
    sigma[:, :rank] = 1.0 + 0.5 * rng.random((n, rank))      # diagonal: positive
    sigma[:, rank:] = 0.2 * (rng.random((n, nv - rank)) - 0.5)  # off-diagonal
    ext = [1.0, -0.5, 0.25][:rank]
    (gm, gf, gterm), (om, of, oterm) = _charge_models(dims, density, sigma, ext)
    dt = 0.01
    gs = gpf.NewSolver(gm, dims, dt)
    osolver = opf.NewSolver(om, dims, dt)
    if stepper == "rk4":
        gs.SetStepper("rk4")
        osolver.SetStepper("rk4")
    gs.Solve(2, 3)
    osolver.Solve(2, 3)
    assert not gs.IsFused and gs.KernelLaunches() > 0
    assert rel_l2(gf.Data, of.Data) <= TOL
    # ChargeTransport.Current on the final state (realspace = true in the oracle: it transforms first)
    cur_g = gterm.Current()
    cur_o = oterm.Current(of, n, True)
    for d in range(rank):
        assert rel_l2(cur_g[d], cur_o[d]) <= TOL


def _source_models(dims, init, sources):
    n = opfutil.prod_int(dims)
    out = []
    for mod in (gpf, opf):
        m = mod.NewModel()
        f = mod.NewField("conc", n, init.astype(np.complex128).copy())
        m.AddField(f)
        m.AddScalar(mod.NewScalar("D", 0.8))
        m.AddEquation("dconc/dt = D*LAP conc - conc^3")
        for pos, fn in sources:
            m.AddSource(0, (gpf if mod is gpf else oterms).NewSource(pos, fn))
        out.append((m, f))
    return out


@pytest.mark.parametrize("dims", [[16, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("stepper", ["euler", "rk4"])
def test_point_sources_vs_oracle(dims, stepper):
    # pf/model.go:291-294: every source of the equation is added to its right-hand side at
    # t = Stepper.GetTime(); two sources with time-dependent amplitudes
    n = opfutil.prod_int(dims)
    rank = len(dims)
    rng = np.random.default_rng(3)
    init = 0.05 * rng.standard_normal(n)
    sources = [([3.0, 5.0, 2.0][:rank], lambda t: 2.0 * t + 0.5), ([10.5, 1.25, 7.0][:rank], lambda t: math.cos(3.0 * t))]
    (gm, gf), (om, of) = _source_models(dims, init, sources)
    dt = 0.05
    gs = gpf.NewSolver(gm, dims, dt)
    osolver = opf.NewSolver(om, dims, dt)
    if stepper == "rk4":
        gs.SetStepper("rk4")
        osolver.SetStepper("rk4")
    gs.Solve(2, 5)
    osolver.Solve(2, 5)
    assert not gs.IsFused
    assert rel_l2(gf.Data, of.Data) <= TOL
    # the sources actually acted: without them the field stays O(0.05)
    assert np.max(np.abs(of.Data)) > 0.0 and rel_l2(of.Data, init) > 1e-3


def test_source_position_shorter_than_rank_is_an_error():
    # Dot(freq, Pos) indexes Pos[k] for every frequency component: a Go panic in the reference
    dims = [16, 16]
    (gm, gf), _ = _source_models(dims, np.zeros(256), [([3.0], lambda t: 1.0)])
    gs = gpf.NewSolver(gm, dims, 0.1)
    with pytest.raises(gpf.GopfError, match="coordinates"):
        gs.Propagate(1)


def test_add_source_after_new_solver_is_refused():
    # the reference reads Model.AllSources live at every GetRHS (pf/model.go:291-294); the device
    # program is compiled once by NewSolver, so a later AddSource must fail loudly, not be ignored
    dims = [16, 16]
    (gm, gf), _ = _source_models(dims, np.zeros(256), [([3.0, 2.0], lambda t: 1.0)])
    gs = gpf.NewSolver(gm, dims, 0.1)
    with pytest.raises(gpf.GopfError, match="before NewSolver"):
        gm.AddSource(0, gpf.NewSource([1.0, 1.0], lambda t: 1.0))
    gs.Propagate(2)
    assert np.max(np.abs(gf.Data)) > 0.0
    gs.close()
    gm.AddSource(0, gpf.NewSource([1.0, 1.0], lambda t: 1.0))  # no solver attached any more


def test_exception_in_source_time_function_is_reported():
    dims = [16, 16]

    def bad(t):
        raise ValueError("boom")

    (gm, gf), _ = _source_models(dims, np.zeros(256), [([3.0, 2.0], bad)])
    gs = gpf.NewSolver(gm, dims, 0.1)
    with pytest.raises(gpf.GopfError, match="boom"):
        gs.Propagate(1)
