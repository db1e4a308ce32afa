"""Oracle restatement of pf.ImplicitEuler (oracle/pf.py) against the reference's own tests
(pf/implicitEuler_test.go).  The Newton-Krylov solve is third-party in the reference
(gononlin + gonum/exp, not in its tree): PARITY UNPINNED beyond these analytic tolerances.
No GPU needed."""
import math

import numpy as np
import pytest

from oracle import pf, pfutil

CASES = [  # pf/implicitEuler_test.go:22-69
    (["conc"], ["dconc/dt = -conc"], [1.0], lambda t: [math.exp(-t)]),
    (["conc"], ["dconc/dt = -conc^2"], [1.0], lambda t: [1.0 / (1.0 + t)]),
    (["conc"], ["dconc/dt = conc - conc^2"], [0.5], lambda t: [math.exp(t) / (1.0 + math.exp(t))]),
    (["conc1", "conc2"], ["dconc1/dt = -conc1*conc2", "dconc2/dt = -conc2"], [1.0, 1.0],
     lambda t: [math.exp(math.exp(-t) - 1.0), math.exp(-t)]),
]


@pytest.mark.parametrize("fields,eqns,init,solution", CASES, ids=["linear", "nonlinear", "both", "coupled"])
def test_implicit_euler(fields, eqns, init, solution):
    # pf/implicitEuler_test.go:10-107: N = 8, dt = 0.01, 100 steps, tolerance 0.005
    N = 8
    m = pf.NewModel()
    for j, name in enumerate(fields):
        f = pf.NewField(name, N * N)
        f.Data[:] = init[j]
        m.AddField(f)
    for e in eqns:
        m.AddEquation(e)
    m.Init()
    st = pf.ImplicitEuler(0.01, pfutil.NewFFTW([N, N]))
    for _ in range(100):
        st.Step(m)
        assert st.last_converged
    expect = solution(0.01 * 100)
    for j, f in enumerate(m.Fields):
        assert np.max(np.abs(f.Data.real - expect[j])) < 0.005 and np.max(np.abs(f.Data.imag)) < 0.005


def test_dissipating_heat_equation():
    # pf/implicitEuler_test.go:166-223 (tolerance 1e-3; the analytic form is the test's own)
    N = 128
    i = np.arange(N * N)
    x, y = (i // N) / float(N), (i % N) / float(N)
    field = pf.NewField("temperature", N * N, (np.sin(2.0 * x * math.pi) * np.sin(2.0 * y * math.pi)).astype(np.complex128))
    m = pf.NewModel()
    gamma = 0.2
    m.AddField(field)
    m.RegisterFunction("DISSIPATE", lambda idx, b: -complex(gamma, 0.0) * b["temperature"].Get(idx))
    m.AddEquation("dtemperature/dt = LAP temperature + DISSIPATE")
    dt = 0.005
    s = pf.NewSolver(m, [N, N], dt)
    s.Stepper = pf.ImplicitEuler(dt, pfutil.NewFFTW([N, N]))
    s.Solve(1, 10)
    L = float(N)
    expect = np.exp(-(4.0 * math.pi / (L * L) + gamma) * 10 * dt) * np.sin(2.0 * math.pi * y) * np.sin(2.0 * math.pi * x)
    assert np.max(np.abs(field.Data.real - expect)) < 1e-3 and np.max(np.abs(field.Data.imag)) < 1e-3


def test_nonlinear_integral_limits():
    # nonlinearIntegral (:151-162): the small-|denum| branch is the limit of the general one
    st = pf.ImplicitEuler(0.1, None)
    rhs = np.array([0.3 + 0.1j, -0.2 + 0.0j])
    prev = np.array([0.25 + 0.05j, -0.1 + 0.0j])
    small = st.nonlinearIntegral(np.array([2e-6 + 0j, 2e-6 + 0j]), rhs, prev)
    big = st.nonlinearIntegral(np.array([2e-5 + 0j, 2e-5 + 0j]), rhs, prev)
    assert np.max(np.abs(small - big)) < 1e-6
    # constant rhs and zero linear part: integral = dt * rhs
    assert np.allclose(st.nonlinearIntegral(np.zeros(2, dtype=complex), rhs, rhs), 0.1 * rhs)


def test_stencils_agree_on_jacobian_vector_product():
    nk2, nk6 = pf.NewtonKrylov(Stencil=2), pf.NewtonKrylov(Stencil=6)
    F = lambda v: v ** 3 - 2.0 * v + np.roll(v, 1)
    rng = np.random.default_rng(0)
    x, v = rng.standard_normal(64), rng.standard_normal(64)
    exact = (3.0 * x * x - 2.0) * v + np.roll(v, 1)
    assert np.max(np.abs(nk2.jac_vec(F, x, v) - exact)) < 1e-4
    assert np.max(np.abs(nk6.jac_vec(F, x, v) - exact)) < 1e-8
