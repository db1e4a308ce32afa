"""Gradient-based catalog terms on the device (gopf_b200/csrc/gradient_terms.cu) through the C ABI, against the
oracle restatements of pf/gradientCalculator.go and pf/advection.go and against the reference's own known answers
(pf/gradientCalculator_test.go, pf/advection_test.go).  The reference reaches these types by calling PrepareModel /
Construct by hand (they are not registrable with a model as shipped); the ABI calls are that surface."""
import math

import numpy as np
import pytest

from gopf_b200 import pfutil as gpfutil
from oracle import pf as opf
from oracle import pfutil as opfutil
from oracle import terms as oterms

pytestmark = pytest.mark.gpu
SIGMA = 0.1


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _field(dims, seed):
    rng = np.random.default_rng(seed)
    n = int(np.prod(dims))
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-0.1, 0.1, n)).astype(np.complex128)


# power-of-two shapes take the in-pass multiplier (LK_GRADIENT_LINE), the others the literal-Freq kernel
SHAPES = [[16, 16], [32, 64], [12, 20], [16, 16, 16], [8, 8, 12]]


@pytest.mark.parametrize("dims", SHAPES, ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("keep", [False, True], ids=["zero-nyquist", "keep-nyquist"])
def test_gradient_calculator_vs_oracle(dims, keep):
    # pf/gradientCalculator.go:19-31
    data = _field(dims, 1)
    ft, oft = gpfutil.NewFFTW(dims), opfutil.NewFFTW(dims)
    for comp in range(len(dims)):
        want = np.zeros_like(data)
        oterms.GradientCalculator(oft, comp, keep).Calculate(data.copy(), want)
        got = np.zeros_like(data)
        ft.GradientCalculate(data.copy(), got, comp, keep)
        assert rel(got, want) <= 1e-12, (dims, comp)


def test_gradient_calculator_reference_known_answer():
    # pf/gradientCalculator_test.go:11-41
    N = 16
    i = np.arange(N * N)
    x = (i % N) / float(N)
    data = (x * x - 2 * x ** 3 + x ** 4).astype(np.complex128)
    expect = (2.0 * x - 6.0 * x * x + 4.0 * x ** 3) / float(N)
    got = np.zeros(N * N, dtype=np.complex128)
    gpfutil.NewFFTW([N, N]).GradientCalculate(data, got, 1)
    assert np.max(np.abs(got.real - expect)) < 1e-4 and np.max(np.abs(got.imag)) < 1e-4


def _oracle_model(dims, arrays):
    m = opf.NewModel()
    n = int(np.prod(dims))
    for name, a in arrays.items():
        m.AddField(opf.NewField(name, n, a.copy()))
    return m


@pytest.mark.parametrize("dims", SHAPES, ids=lambda d: "x".join(map(str, d)))
def test_advection_vs_oracle(dims):
    # pf/advection.go:50-96: PrepareModel's derived fields, then Construct
    n = int(np.prod(dims))
    names = ["vx", "vy", "vz"][:len(dims)]
    arrays = {"conc": _field(dims, 2)}
    for k, v in enumerate(names):
        arrays[v] = _field(dims, 10 + k)
    m = _oracle_model(dims, arrays)
    oft = opfutil.NewFFTW(dims)
    adv = oterms.Advection("conc", names)
    adv.PrepareModel(n, m, oft)
    m.Init()
    want = np.zeros(n, dtype=np.complex128)
    adv.Construct(m.Bricks)(oft.Freq, 0.0, want)
    ft = gpfutil.NewFFTW(dims)
    got = np.zeros(n, dtype=np.complex128)
    ft.AdvectionConstruct(arrays["conc"], [arrays[v] for v in names], got)
    assert rel(got, want) <= 1e-12
    # as a step sees it: the derived field transformed
    got_t = np.zeros(n, dtype=np.complex128)
    ft.AdvectionConstruct(arrays["conc"], [arrays[v] for v in names], got_t, transformed=True)
    assert rel(got_t, oft.FFT(want.copy())) <= 1e-12


def test_advection_reference_known_answer_and_velocity_count():
    # pf/advection_test.go:94-170 (case vx = x) and :172-216 (wrong number of velocity fields)
    from gopf_b200._lib import GopfError
    N = 64
    i = np.arange(N * N)
    x = (i % N) / float(N) - 0.5
    y = (i // N) / float(N) - 0.5
    g = np.exp(-0.5 * (x * x + y * y) / (SIGMA * SIGMA)).astype(np.complex128)
    ft = gpfutil.NewFFTW([N, N])
    res = np.zeros(N * N, dtype=np.complex128)
    ft.AdvectionConstruct(g, [x.astype(np.complex128), np.zeros(N * N, dtype=np.complex128)], res)
    expect = y * x * g.real / (SIGMA * SIGMA)
    assert np.max(np.abs(res.real * N - expect)) < 1e-3 and np.max(np.abs(res.imag * N)) < 1e-3
    with pytest.raises(GopfError, match="Inconsistent number of velocity fields"):
        ft.AdvectionConstruct(g, [g], res)


@pytest.mark.parametrize("dims", SHAPES, ids=lambda d: "x".join(map(str, d)))
def test_div_grad_vs_oracle(dims):
    # pf/gradientCalculator.go:72-108 with F = 1 + field^2
    n = int(np.prod(dims))
    data = _field(dims, 3)
    F = lambda idx, b: 1.0 + b["myfield"].Get(idx) ** 2
    m = _oracle_model(dims, {"myfield": data})
    oft = opfutil.NewFFTW(dims)
    dg = oterms.DivGrad("myfield", F)
    dg.PrepareModel(n, m, oft)
    m.Init()
    rhs = dg.Construct(m.Bricks)
    for d in m.DerivedFields:
        oft.FFT(d.Data)
    want = np.zeros(n, dtype=np.complex128)
    rhs(oft.Freq, 0.0, want)
    got = np.zeros(n, dtype=np.complex128)
    gpfutil.NewFFTW(dims).DivGradConstruct(data, (1.0 + data ** 2).astype(np.complex128), got)
    assert rel(got, want) <= 1e-12


def test_div_grad_reference_known_answer():
    # pf/gradientCalculator_test.go:64-148: div(c grad c) of a Gaussian
    N = 64
    i = np.arange(N * N)
    x = (i % N) / float(N) - 0.5
    y = (i // N) / float(N) - 0.5
    data = np.exp(-0.5 * (x * x + y * y) / (SIGMA * SIGMA))
    want = ((2.0 * x * x + 2.0 * y * y) / (SIGMA * SIGMA) - 2.0) * data * data / (SIGMA * SIGMA)
    ft = gpfutil.NewFFTW([N, N])
    res = np.zeros(N * N, dtype=np.complex128)
    ft.DivGradConstruct(data.astype(np.complex128), data.astype(np.complex128), res)
    ft.IFFT(res)
    res /= N * N
    re = res.real * float(N * N)
    assert np.all((np.abs(re - want) < 1e-3) | (np.abs(re - want) < want * 1e-3))
    assert np.max(np.abs(res.imag)) < 1e-10


@pytest.mark.parametrize("dims", SHAPES, ids=lambda d: "x".join(map(str, d)))
def test_weighted_laplacian_vs_oracle(dims):
    # pf/gradientCalculator.go:131-172; both bricks hold spectra
    n = int(np.prod(dims))
    oft = opfutil.NewFFTW(dims)
    field_hat = oft.FFT(_field(dims, 4))
    pre_hat = oft.FFT(_field(dims, 5))
    wl = oterms.WeightedLaplacian("field", "prefactor", oft)
    want = np.zeros(n, dtype=np.complex128)
    wl.Construct({"field": opf.NewField("field", n, field_hat.copy()),
                  "prefactor": opf.NewField("prefactor", n, pre_hat.copy())})(oft.Freq, 0.0, want)
    got = np.zeros(n, dtype=np.complex128)
    gpfutil.NewFFTW(dims).WeightedLaplacianConstruct(field_hat, pre_hat, got)
    assert rel(got, want) <= 1e-12


def test_weighted_laplacian_reference_known_answer():
    # pf/gradientCalculator_test.go:150-196
    N = 16
    i = np.arange(N * N)
    x = (i // N) / float(N)
    two_pi = 2.0 * math.pi
    ft = gpfutil.NewFFTW([N, N])
    field_hat = ft.FFT(np.cos(two_pi * x).astype(np.complex128))
    pre_hat = ft.FFT(np.sin(two_pi * x).astype(np.complex128))
    res = np.zeros(N * N, dtype=np.complex128)
    ft.WeightedLaplacianConstruct(field_hat, pre_hat, res)
    ft.IFFT(res)
    res /= float(N * N)
    expect = -two_pi ** 2 * np.sin(two_pi * x) * np.cos(two_pi * x) / float(N * N)
    assert np.max(np.abs(res - expect)) < 1e-10
