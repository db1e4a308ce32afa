"""Pins oracle/terms.py against the reference's own known-answer tests.
Each test cites the Go test it restates.  No GPU needed.
"""
import math

import numpy as np
import pytest

from oracle import pf, pfutil, terms


def sin_x(nx, ny):
    # pf/squareGradientTerm_test.go:11-21
    i = np.arange(nx * ny)
    x = (i % nx) / float(nx)
    data = np.sin(2.0 * math.pi * x)
    grad_sq = np.power(2.0 * math.pi * np.cos(2.0 * math.pi * x) / float(nx), 2.0)
    return data, grad_sq


def poly4(nx, ny):
    # pf/squareGradientTerm_test.go:24-41
    i = np.arange(nx * ny)
    x = (i % nx) / float(nx)
    y = (i // nx) / float(ny)
    vx = 16.0 * (x * x - 2 * x ** 3 + x ** 4)
    vy = 16.0 * (y * y - 2 * y ** 3 + y ** 4)
    ddx = 16.0 * (2 * x - 6 * x * x + 4 * x ** 3) * vy / float(nx)
    ddy = vx * 16.0 * (2 * y - 6 * y * y + 4 * y ** 3) / float(ny)
    return vx * vy, ddx * ddx + ddy * ddy


@pytest.mark.parametrize("F,N,tol", [(sin_x, 16, 1e-10), (poly4, 32, 1e-5)])
def test_square_gradient(F, N, tol):
    # pf/squareGradientTerm_test.go:43-82
    data, grad_sq = F(N, N)
    grad = terms.NewSquareGradient("height", [N, N])
    field = pf.NewField("height", N * N, data.astype(np.complex128))
    grad.FT.FFT(field.Data)
    bricks = {"height": field}
    res = np.zeros(N * N, dtype=np.complex128)
    grad.Construct(bricks)(grad.FT.Freq, 0.0, res)
    grad.FT.IFFT(res)
    res /= N * N
    assert np.max(np.abs(res.real - grad_sq)) < tol and np.max(np.abs(res.imag)) < tol


def test_square_grad_with_solver():
    # pf/squareGradientTerm_test.go:84-128
    N = 16
    model = pf.NewModel()
    field1 = pf.NewField("field1", N * N)
    field2 = pf.NewField("field2", N * N)
    data, grad_sq = sin_x(N, N)
    field2.Data[:] = data
    grad = terms.NewSquareGradient("field2", [N, N])
    model.AddField(field1)
    model.AddField(field2)
    model.AddScalar(pf.NewScalar("ZERO", 0j))
    model.RegisterExplicitTerm("GRAD_SQ_f2", grad, None)
    model.AddEquation("dfield1/dt = GRAD_SQ_f2")
    model.AddEquation("dfield2/dt = ZERO*field1")
    dt, nsteps = 0.1, 10
    solver = pf.NewSolver(model, [N, N], dt)
    solver.Solve(1, nsteps)
    f2 = model.Bricks["field2"].Data
    f1 = model.Bricks["field1"].Data
    assert np.max(np.abs(f2.real - data)) < 1e-10 and np.max(np.abs(f2.imag)) < 1e-10
    assert np.max(np.abs(f1.real - dt * nsteps * grad_sq)) < 1e-10 and np.max(np.abs(f1.imag)) < 1e-10


@pytest.mark.parametrize("order", [3, 5, 10])
def test_vandeven(order):
    # pf/vandeven_test.go:8-39
    filt = terms.NewVandeven(order)
    for x, y in [(0.0, 1.0), (0.5, 0.5), (1.0, 0.0)]:
        assert abs(filt.Eval(x) - y) < 1e-4
    xs = np.linspace(0.0, 1.3, 777)
    assert np.array_equal(filt.eval_array(xs), np.array([filt.Eval(float(x)) for x in xs]))


def test_interpolation():
    # pf/spectralViscosity_test.go:10-35
    for k, expect in [(0.19, 0.0), (0.66, 1.0), (0.4, 1.0 / 8.0)]:
        assert abs(terms.interpolant(k, 0.6) - expect) < 1e-10
        assert abs(terms.interpolant_array(np.array([k]), 0.6)[0] - expect) < 1e-10


def test_spectral_viscosity_term():
    # pf/spectralViscosity_test.go:37-75
    sv = terms.SpectralViscosity(1.0, 0.25, 2)
    model = pf.NewModel()
    N = 16
    model.AddField(pf.NewField("conc", N * N))
    model.RegisterImplicitTerm("SPECTRAL_VISC", sv, None)
    model.AddEquation("dconc/dt = SPECTRAL_VISC")
    model.Init()
    rhs = model.RHS[0]
    assert len(rhs.Denum) == 1 and len(rhs.Terms) == 0
    ft = pfutil.NewFFTW([N, N])
    res = np.zeros(N * N, dtype=np.complex128)
    rhs.Denum[0](ft.Freq, 0.0, res)
    for i in range(N * N):
        fv = ft.Freq(i)
        f = fv[0] * fv[0] + fv[1] * fv[1]
        expect = -sv.Eps * terms.interpolant(math.sqrt(f), sv.DissipationThreshold) * f
        assert abs(res[i].real - expect) < 1e-10 and abs(res[i].imag) < 1e-10


def test_white_noise_variance():
    # pf/noise_test.go:23-48: std = sqrt(2*Strength) = 2.0 for Strength 2
    rng = np.random.default_rng(0)
    noise = terms.WhiteNoise(2.0, normal=rng.standard_normal)
    data = noise.Generate(np.arange(1_000_000), {}).real
    assert abs(np.std(data, ddof=1) - 2.0) < 0.005


def test_conservative_noise():
    # pf/noise_test.go:50-94
    rng = np.random.default_rng(3)
    model = pf.NewModel()
    N = 16
    field = pf.NewField("myfield", N * N)
    model.AddField(field)
    noise = terms.ConservativeNoise(1.0, 2, unique_prefix=1234, normal=rng.standard_normal)
    model.RegisterExplicitTerm("CONSERVATIVE_NOISE", noise, noise.RequiredDerivedFields(N * N))
    model.AddEquation("dmyfield/dt = CONSERVATIVE_NOISE")
    solver = pf.NewSolver(model, [N, N], 0.1)
    solver.Solve(10, 100)
    assert np.max(np.abs(field.Data.imag)) < 1e-10
    assert abs(field.Data.real.sum()) < 1e-10
    assert np.count_nonzero(np.abs(field.Data.real) > math.sqrt(2.0)) > 0


def test_vol_conserve():
    # pf/volumeConserving_test.go:15-64
    N, dt = 16, 0.01
    vol = terms.NewVolumeConservingLP("myfield", "indicator", dt, N * N)
    indicator = pf.NewField("indicator", N * N)
    myfield = pf.NewField("myfield", N * N)
    indicator.Data[N * N // 2 + 1:] = 1.0
    myfield.Data[:] = 1.0
    bricks = {"indicator": indicator, "myfield": myfield}
    ft = pfutil.NewFFTW([N, N])
    ft.FFT(indicator.Data)
    fn = vol.Construct(bricks)
    rate = 0.01
    corr = np.zeros(N * N, dtype=np.complex128)
    for _ in range(100):
        ft.IFFT(indicator.Data)
        pfutil.div_real_scalar(indicator.Data, float(N * N))
        fn(lambda i: [0.0, 0.0], 0.0, corr)
        myfield.Data += complex(dt, 0.0) * (complex(rate, 0.0) + corr)
        ft.FFT(indicator.Data)
        vol.OnStepFinished(0.0, bricks)
    volume = myfield.Data.real.sum()
    expect_change = 2.0 * dt * float(N * N) * rate
    assert abs(volume - expect_change - 256.0) < 1e-10


def _single_peak(loc, width=100.0):
    return terms.ReciprocalSpacePairCorrelation(0.0, [terms.Peak(1, loc, width, 1)])


def test_pair_correlation_term():
    # pf/pairCorrelationTerm_test.go:11-58
    pair = terms.PairCorrlationTerm(_single_peak(1.0), "myfield", 1.0, False)
    N = 16
    field = pf.NewField("myfield", N * N)
    field.Data[:] = 0.1 * np.arange(N * N)
    res = np.zeros(N * N, dtype=np.complex128)
    freq = lambda i: [float(i), float(2 * i)]
    pair.Construct({"myfield": field})(freq, 0.0, res)
    for i in range(N * N):
        f = freq(i)
        f_rad = 2.0 * math.pi * math.sqrt(f[0] * f[0] + f[1] * f[1])
        expect = -math.exp(-0.5 * (f_rad - 1.0) * (f_rad - 1.0) / 100.0 ** 2)
        assert abs(res[i].real - expect) < 1e-10


def test_pair_correlation_get_energy():
    # pf/pairCorrelationTerm_test.go:60-96
    wavenumber = 2.0 * math.pi / 4
    pair = terms.PairCorrlationTerm(_single_peak(wavenumber), "myfield", 1.0, False)
    N = 16
    field = pf.NewField("myfield", N * N)
    field.Data[:] = np.cos(wavenumber * (np.arange(N * N) % N))
    ft = pfutil.NewFFTW([N, N])
    energy = pair.GetEnergy({"myfield": field}, ft, [N, N])
    assert abs(energy - (-0.5 * N * N * 0.5)) < 1e-10


def test_ideal_mix_term():
    # pf/pairCorrelationTerm_test.go:98-135
    term = terms.IdealMixtureTerm(terms.IdealMix(1.0, 1.0), "eta", 1.0, False)
    N = 16
    data = (0.1 * np.arange(N * N)).astype(np.complex128)
    field = pf.NewField("eta", N * N, data)
    bricks = {"eta": field}
    val = term.Eval(np.arange(N * N), bricks)
    assert np.max(np.abs(val.real - term.IdealMix.Deriv(data.real))) < 1e-10
    field.Data[:] = 1.0
    assert abs(term.GetEnergy(bricks, N * N) - 5.0 * N * N / 12.0) < 1e-10


@pytest.mark.parametrize("lap", [False, True])
def test_ideal_mix_term_with_model(lap):
    # pf/pairCorrelationTerm_test.go:239-311
    N = 16
    field = pf.NewField("density", N * N)
    field.Data[:] = 0.5
    term = terms.IdealMixtureTerm(terms.IdealMix(1.0, 1.0), "density", 1.0, lap)
    model = pf.NewModel()
    model.AddField(field)
    model.RegisterMixedTerm("IDEAL_MIX", term, None)
    model.AddEquation("ddensity/dt = IDEAL_MIX")
    model.Init()
    freq = lambda i: [0.0, 0.6]
    with pytest.raises(RuntimeError):
        model.GetRHS(0, freq, 0.0)
    model.RegisterDerivedField(term.DerivedField(N * N, model.Bricks))
    model.Init()
    lapfac = -4.0 * math.pi * math.pi * 0.6 * 0.6 if lap else 1.0
    den = model.GetDenum(0, freq, 0.0)
    assert np.max(np.abs(den.real - 1.0 * lapfac)) < 1e-10 and np.max(np.abs(den.imag)) < 1e-10
    rhs = model.GetRHS(0, freq, 0.0)
    expect = (-0.5 * 0.5 ** 2 + 0.5 ** 3 / 3.0) * lapfac
    assert np.max(np.abs(rhs.real - expect)) < 1e-10 and np.max(np.abs(rhs.imag)) < 1e-10


@pytest.mark.parametrize("laplace", [False, True])
def test_explicit_pair_corr_term(laplace):
    # pf/pairCorrelationTerm_test.go:313-364
    term = terms.ExplicitPairCorrelationTerm(_single_peak(1.0), "myfield", 1.0, laplace)
    N = 4
    field = pf.NewField("myfield", N * N)
    field.Data[:] = 2.5
    freq = lambda i: [1.0 / (2.0 * math.pi), 0.0]
    result = np.zeros(N * N, dtype=np.complex128)
    term.Construct({"myfield": field})(freq, 0.0, result)
    expect = 2.5 * freq(0)[0] ** 2 * 4.0 * math.pi * math.pi if laplace else -2.5
    assert np.max(np.abs(result.real - expect)) < 1e-10 and np.all(result.imag == 0.0)


def test_tensorial_hessian():
    # pf/tensorialHessian_test.go:11-103: K picks d2/dx2, d2/dy2 or the mixed derivative of a polynomial bump
    N = 64
    i = np.arange(N * N)
    y, x = (i % N) / float(N), (i // N) / float(N)
    px = 16.0 * (x * x - 2.0 * x ** 3 + x ** 4)
    py = 16.0 * (y * y - 2.0 * y ** 3 + y ** 4)
    want = {
        "dx2": 16.0 * (2.0 - 12.0 * x + 12.0 * x * x) * py / float(N * N),
        "dy2": px * 16.0 * (2.0 - 12.0 * y + 12.0 * y * y) / float(N * N),
        "dxdy": 16.0 * (2.0 * x - 6.0 * x * x + 4.0 * x ** 3) * 16.0 * (2.0 * y - 6.0 * y * y + 4.0 * y ** 3) / float(N * N),
    }
    ft = pfutil.NewFFTW([N, N])
    data = (px * py).astype(np.complex128)
    ft.FFT(data)
    for K, key, tol in [([1.0, 0.0, 0.0, 0.0], "dx2", 1e-3), ([0.0, 0.0, 0.0, 1.0], "dy2", 1e-3), ([0.0, 0.5, 0.5, 0.0], "dxdy", 1e-6)]:
        res = np.zeros(N * N, dtype=np.complex128)
        terms.TensorialHessian(K).Construct({})(ft.Freq, 0.0, res)
        res *= data
        ft.IFFT(res)
        res /= N * N
        assert np.max(np.abs(res.real - want[key])) < tol and np.max(np.abs(res.imag)) < tol


def test_hessian_with_model():
    # pf/tensorialHessian_test.go:105-146: an implicit user term adds no explicit term
    N = 16
    m = pf.NewModel()
    m.AddField(pf.NewField("conc1", N * N, np.arange(N * N, dtype=np.float64).astype(np.complex128)))
    m.AddField(pf.NewField("conc2", N * N, np.arange(N * N, dtype=np.float64).astype(np.complex128)))
    m.RegisterImplicitTerm("HESSIAN", terms.TensorialHessian([1.0, 2.0, 2.0, 2.0]), None)
    m.AddEquation("dconc1/dt = HESSIAN")
    m.AddEquation("dconc2/dt = -conc2")
    m.Init()
    assert len(m.RHS[0].Terms) == 0 and len(m.RHS[0].Denum) == 1
