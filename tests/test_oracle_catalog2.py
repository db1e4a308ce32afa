"""Oracle restatements of the remaining catalog terms (GradientCalculator,
DivGrad, WeightedLaplacian, Advection, Source, NegativeValuePenalty) pinned to the reference's own tests.
The device side of the gradient-based ones is gopf_b200/csrc/gradient_terms.cu, checked against these
restatements in tests/test_terms_gradient_gpu.py.  No GPU needed."""
import math

import numpy as np
import pytest

from oracle import pf, pfutil, terms

SIGMA = 1.0 / 10.0


def test_gradient_calculator():
    # pf/gradientCalculator_test.go:11-41
    N = 16
    i = np.arange(N * N)
    x = (i % N) / float(N)
    data = (x * x - 2 * x ** 3 + x ** 4).astype(np.complex128)
    expect = (2.0 * x - 6.0 * x * x + 4.0 * x ** 3) / float(N)
    grad = terms.GradientCalculator(pfutil.NewFFTW([N, N]), 1)
    got = np.zeros(N * N, dtype=np.complex128)
    grad.Calculate(data, got)
    assert np.max(np.abs(got.real - expect)) < 1e-4 and np.max(np.abs(got.imag)) < 1e-4


def test_div_grad():
    # pf/gradientCalculator_test.go:64-160: div(c grad c) of a Gaussian
    dg = terms.DivGrad("myfield", lambda idx, b: b["myfield"].Get(idx))
    assert dg.FuncName() == "DivGrad_myfield_Func" and dg.GradName(1) == "GRAD_myfield_1"
    N = 64
    i = np.arange(N * N)
    x = (i % N) / float(N) - 0.5
    y = (i // N) / float(N) - 0.5
    data = np.exp(-0.5 * (x * x + y * y) / (SIGMA * SIGMA))
    want = ((2.0 * x * x + 2.0 * y * y) / (SIGMA * SIGMA) - 2.0) * data * data / (SIGMA * SIGMA)
    m = pf.NewModel()
    field = pf.NewField("myfield", N * N, data.astype(np.complex128))
    m.AddField(field)
    ft = pfutil.NewFFTW([N, N])
    dg.PrepareModel(N * N, m, ft)
    m.Init()
    rhs = dg.Construct(m.Bricks)
    for f in m.Fields:
        ft.FFT(f.Data)
    for d in m.DerivedFields:
        ft.FFT(d.Data)
    res = np.zeros(N * N, dtype=np.complex128)
    rhs(ft.Freq, 0.0, res)
    ft.IFFT(res)
    res /= N * N
    re = res.real * float(N * N)
    ok = (np.abs(re - want) < 1e-3) | (np.abs(re - want) < want * 1e-3)
    assert np.all(ok)
    assert np.max(np.abs(res.imag)) < 1e-10


def test_weighted_laplacian():
    # pf/gradientCalculator_test.go:150-196: sin(2 pi x) LAP cos(2 pi x) on 16 x 16
    N = 16
    i = np.arange(N * N)
    x = (i // N) / float(N)  # pfutil.Pos(...)[0]
    two_pi = 2.0 * math.pi
    prefactor = pf.NewField("prefactor", N * N, np.sin(two_pi * x).astype(np.complex128))
    field = pf.NewField("field", N * N, np.cos(two_pi * x).astype(np.complex128))
    expect = -two_pi ** 2 * np.sin(two_pi * x) * np.cos(two_pi * x) / float(N * N)
    ft = pfutil.NewFFTW([N, N])
    ft.FFT(field.Data)
    ft.FFT(prefactor.Data)
    wl = terms.WeightedLaplacian("field", "prefactor", ft)
    result = np.zeros(N * N, dtype=np.complex128)
    wl.Construct({"prefactor": prefactor, "field": field})(ft.Freq, 0.0, result)
    ft.IFFT(result)
    result /= float(N * N)
    assert np.max(np.abs(result - expect)) < 1e-10


def _gauss(N):
    i = np.arange(N * N)
    x = (i % N) / float(N) - 0.5
    y = (i // N) / float(N) - 0.5
    return x, y, np.exp(-0.5 * (x * x + y * y) / (SIGMA * SIGMA))


@pytest.mark.parametrize("case", ["vx", "vy", "linear"])
def test_advection(case):
    # pf/advection_test.go:94-170
    N = 64
    x, y, g = _gauss(N)
    vx, vy = np.zeros(N * N, dtype=np.complex128), np.zeros(N * N, dtype=np.complex128)
    if case == "vx":
        vx[:] = 1.0
        expect = y * g / (SIGMA * SIGMA)
    elif case == "vy":
        vy[:] = 1.0
        expect = x * g / (SIGMA * SIGMA)
    else:
        vx[:] = x
        expect = y * x * g / (SIGMA * SIGMA)
    m = pf.NewModel()
    m.AddField(pf.NewField("conc", N * N, g.astype(np.complex128)))
    m.AddField(pf.NewField("vx", N * N, vx))
    m.AddField(pf.NewField("vy", N * N, vy))
    adv = terms.Advection("conc", ["vx", "vy"])
    ft = pfutil.NewFFTW([N, N])
    adv.PrepareModel(N * N, m, ft)
    m.Init()
    res = np.zeros(N * N, dtype=np.complex128)
    adv.Construct(m.Bricks)(ft.Freq, 0.0, res)
    assert np.max(np.abs(res.real * N - expect)) < 1e-3 and np.max(np.abs(res.imag * N)) < 1e-3


def test_advection_panics():
    # pf/advection_test.go:172-216
    m = pf.NewModel()
    adv = terms.Advection("conc", ["vx"])
    ft = pfutil.NewFFTW([8, 8])
    with pytest.raises(RuntimeError):
        adv.PrepareModel(64, m, ft)
    for name in ("conc", "vx", "vy"):
        m.AddField(pf.NewField(name, 64))
    with pytest.raises(RuntimeError):
        adv.PrepareModel(64, m, ft)  # wrong number of velocity fields
    terms.Advection("conc", ["vx", "vy"]).PrepareModel(64, m, ft)


def test_source_term():
    # pf/sourceTerm_test.go:19-31
    src = terms.NewSource([2.5], lambda t: 2.0 * t)
    data = np.zeros(2, dtype=np.complex128)
    src.Eval(pf.Frequency(lambda i: [float(i)], lambda cnt: np.arange(cnt, dtype=np.float64).reshape(-1, 1)), 2.0, data)
    expect = 4.0 * np.exp(-1j * 2.0 * math.pi * 2.5 * np.array([0.0, 1.0]))
    assert np.max(np.abs(data - expect)) < 1e-10


def test_negative_value_penalty():
    # pf/negative_value_penalty_test.go:8-36
    nvp = terms.NewDefaultNegativeValuePenalty("myfield")
    for value, expect in [(1.0, 0.0), (0.0, 0.0), (-1.0, -2.0 * nvp.Prefactor * float(nvp.Exponent))]:
        assert abs(float(nvp.Penalty(value)) - expect) < 1e-6


@pytest.mark.parametrize("case", ["zero", "sine"])
def test_charge_transport(case):
    # pf/chargeTransport_test.go:10-133
    N = 32
    idx = np.arange(N * N)
    x = np.array([pfutil.pos([N, N], int(i))[0] for i in idx]) / float(N)
    if case == "zero":
        rho = np.zeros(N * N)
        want_div = np.zeros(N * N)
        want_cur = [np.ones(N * N), np.zeros(N * N)]
    else:
        rho = np.sin(2.0 * math.pi * x)
        want_div = -np.sin(2.0 * math.pi * x)
        want_cur = [1.0 - float(N) * np.cos(2.0 * math.pi * x) / (2.0 * math.pi), np.zeros(N * N)]
    ft = pfutil.NewFFTW([N, N])
    ct = terms.ChargeTransport(lambda i: np.array([1.0, 1.0, 0.0]), [1.0, 0.0], "density", ft)
    field = pf.NewField("density", N * N, rho.astype(np.complex128))
    orig = field.Data.copy()
    ft.FFT(field.Data)
    result = np.zeros(N * N, dtype=np.complex128)
    ct.Construct({"density": field})(ft.Freq, 0.0, result)
    ft.IFFT(result)
    result /= N * N
    assert np.max(np.abs(result.real - want_div)) < 1e-10 and np.max(np.abs(result.imag)) < 1e-10
    cur = ct.Current(field, N * N, False)
    for d in range(2):
        assert np.max(np.abs(cur[d] - want_cur[d])) < 1e-10
    field.Data[:] = orig
    cur_real = ct.Current(field, N * N, True)
    for d in range(2):
        assert np.max(np.abs(cur_real[d] - cur[d])) < 1e-10
