"""pf.ImplicitEuler on the device (gopf_b200/csrc/implicit_euler.cu) against the reference's own
tests (analytic tolerances) and against the oracle's restatement.  The non-linear solve is
third-party in the reference: parity holds to the solver tolerance (max|F| < 1e-7), so the
device-vs-oracle tolerance here is 1e-6 relative, not the 1e-10 of the explicit steppers."""
import math

import numpy as np
import pytest

from gopf_b200 import pf as gpf
from gopf_b200 import synthetic
from oracle import pf as opf
from oracle import pfutil as opfutil

pytestmark = pytest.mark.gpu

CASES = [  # pf/implicitEuler_test.go:22-69
    (["conc"], ["dconc/dt = -conc"], [1.0], lambda t: [math.exp(-t)]),
    (["conc"], ["dconc/dt = -conc^2"], [1.0], lambda t: [1.0 / (1.0 + t)]),
    (["conc"], ["dconc/dt = conc - conc^2"], [0.5], lambda t: [math.exp(t) / (1.0 + math.exp(t))]),
    (["conc1", "conc2"], ["dconc1/dt = -conc1*conc2", "dconc2/dt = -conc2"], [1.0, 1.0],
     lambda t: [math.exp(math.exp(-t) - 1.0), math.exp(-t)]),
]


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("fields,eqns,init,solution", CASES, ids=["linear", "nonlinear", "both", "coupled"])
def test_implicit_euler(fields, eqns, init, solution):
    # pf/implicitEuler_test.go:10-107
    N = 8
    m = gpf.NewModel()
    fs = []
    for j, name in enumerate(fields):
        f = gpf.NewField(name, N * N)
        f.Data[:] = init[j]
        m.AddField(f)
        fs.append(f)
    for e in eqns:
        m.AddEquation(e)
    s = gpf.NewSolver(m, [N, N], 0.01)
    s.Stepper = gpf.ImplicitEuler(0.01)
    s.Upload()
    s.StepDevice(100)
    s.Download()
    assert s.Stepper.Converged and s.Stepper.ResidualEvaluations > 0
    assert abs(s.Stepper.GetTime() - 1.0) < 1e-12
    expect = solution(1.0)
    for j, f in enumerate(fs):
        assert np.max(np.abs(f.Data.real - expect[j])) < 0.005 and np.max(np.abs(f.Data.imag)) < 0.005


def test_dissipating_heat_equation():
    # pf/implicitEuler_test.go:166-223
    N = 128
    i = np.arange(N * N)
    x, y = (i // N) / float(N), (i % N) / float(N)
    field = gpf.NewField("temperature", N * N, (np.sin(2.0 * x * math.pi) * np.sin(2.0 * y * math.pi)).astype(np.complex128))
    m = gpf.NewModel()
    m.AddField(field)
    m.RegisterFunction("DISSIPATE", "-0.2*temperature")
    m.AddEquation("dtemperature/dt = LAP temperature + DISSIPATE")
    dt = 0.005
    s = gpf.NewSolver(m, [N, N], dt)
    s.Stepper = gpf.ImplicitEuler(dt)
    s.Solve(1, 10)
    L = float(N)
    expect = np.exp(-(4.0 * math.pi / (L * L) + 0.2) * 10 * dt) * np.sin(2.0 * math.pi * y) * np.sin(2.0 * math.pi * x)
    assert np.max(np.abs(field.Data.real - expect)) < 1e-3 and np.max(np.abs(field.Data.imag)) < 1e-3


@pytest.mark.parametrize("stencil", [2, 6])
@pytest.mark.parametrize("dims", [[32, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_cahn_hilliard_vs_oracle(dims, stencil):
    # a spatially varying, stiff problem: Cahn-Hilliard with the implicit stepper, both sides with
    # the same Newton-Krylov settings
    n = opfutil.prod_int(dims)
    init = 0.1 * synthetic.cahn_hilliard_initial(n, 3)
    res = []
    for mod in (gpf, opf):
        m = mod.NewModel()
        f = mod.NewField("conc", n, init.copy())
        m.AddScalar(mod.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
        m.AddScalar(mod.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
        m.AddField(f)
        m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
        s = mod.NewSolver(m, dims, 0.05)
        if mod is gpf:
            s.Stepper = gpf.ImplicitEuler(0.05, NonlinSolver=gpf.NewtonKrylov(Stencil=stencil))
        else:
            s.Stepper = opf.ImplicitEuler(0.05, opfutil.NewFFTW(dims), NonlinSolver=opf.NewtonKrylov(Stencil=stencil))
        s.Solve(1, 5)
        res.append(f.Data.copy())
        if mod is gpf:
            assert s.Stepper.Converged
        else:
            assert s.Stepper.last_converged
    assert rel_l2(res[0], res[1]) <= 1e-6
    # and the implicit stepper did something different from a no-op: the field moved
    assert rel_l2(res[0], init) > 1e-3
