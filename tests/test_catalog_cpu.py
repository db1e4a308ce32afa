"""CPU checks of the ChargeTransport / Source host logic behind the C ABI: the per-k factors
the kernels apply (the same __host__ __device__ code, gopf_b200/csrc/catalog_terms.cuh) against
the oracle restatement of pf/chargeTransport.go and pf/sourceTerm.go, and the model-side
registration rules.  No compute entry point runs (no GPU here)."""
import ctypes
import math

import numpy as np
import pytest

from gopf_b200 import _lib
from gopf_b200 import pf as gpf
from oracle import pfutil as opfutil
from oracle import terms as oterms
from oracle.pf import as_frequency

PD = ctypes.POINTER(ctypes.c_double)


def k_table(dims):
    ft = opfutil.NewFFTW(dims)
    n = opfutil.prod_int(dims)
    return np.ascontiguousarray(as_frequency(ft.Freq).table(n), dtype=np.float64)


@pytest.mark.parametrize("dims", [[8, 16], [9, 9], [8, 8, 8], [5, 5, 5]], ids=lambda d: "x".join(map(str, d)))
def test_charge_transport_multipliers_match_oracle_formula(dims):
    # pf/chargeTransport.go:64-73 and :106-112, every node of the grid (Nyquist planes included)
    k = k_table(dims)
    n, rank = k.shape
    fm = np.empty((rank, n))
    dm = np.empty((rank, n))
    _lib.check(_lib.lib().gopf_charge_transport_multipliers(rank, k.ctypes.data_as(PD), ctypes.c_int64(n),
                                                            fm.ctypes.data_as(PD), dm.ctypes.data_as(PD)))
    k_sq = np.zeros(n)
    for c in range(rank):  # pfutil.Dot accumulates in component order
        k_sq = k_sq + k[:, c] * k[:, c]
    for c in range(rank):
        keep = np.abs(np.abs(k[:, c]) - 0.5) > 1e-10
        expect_f = np.where(keep, k[:, c] / (2.0 * math.pi * k_sq + 1e-16), 0.0)
        expect_d = np.where(keep, 2.0 * math.pi * k[:, c], 0.0)
        assert np.array_equal(fm[c], expect_f)
        assert np.array_equal(dm[c], expect_d)
        if dims[0] % 2 == 0:
            assert (~keep).any() and np.all(fm[c][~keep] == 0.0)


def test_voigt_index_matches_reference_tables():
    # pf/chargeTransport.go:151-171
    for dim in (2, 3):
        for i in range(dim):
            for j in range(dim):
                assert gpf.voigtIndex(i, j, dim) == oterms.voigtIndex(i, j, dim)
    assert [[gpf.voigtIndex(i, j, 3) for j in range(3)] for i in range(3)] == [[0, 5, 4], [5, 1, 3], [4, 3, 2]]
    assert [[gpf.voigtIndex(i, j, 2) for j in range(2)] for i in range(2)] == [[0, 2], [2, 1]]
    out = ctypes.c_int(0)
    assert _lib.lib().gopf_charge_transport_voigt_index(2, 0, 2, ctypes.byref(out)) != 0


def test_source_eval_reference_kat():
    # pf/sourceTerm_test.go:19-31: freq(i) = [i], Pos = [2.5], f(t) = 2t at t = 2
    freq = np.array([[0.0], [1.0]])
    pos = np.array([2.5])
    out = np.empty(2, dtype=np.complex128)
    _lib.check(_lib.lib().gopf_source_eval(1, freq.ctypes.data_as(PD), ctypes.c_int64(2), pos.ctypes.data_as(PD),
                                           ctypes.c_double(4.0), out.ctypes.data_as(PD)))
    expect = np.array([4.0 * np.exp(-1j * 2.0 * math.pi * 2.5 * 0.0), 4.0 * np.exp(-1j * 2.0 * math.pi * 2.5 * 1.0)])
    assert np.max(np.abs(out - expect)) < 1e-10


@pytest.mark.parametrize("dims", [[8, 16], [6, 6, 6]], ids=lambda d: "x".join(map(str, d)))
def test_source_eval_matches_oracle(dims):
    k = k_table(dims)
    n, rank = k.shape
    pos = [1.5, 3.25, 2.0][:rank]
    src = oterms.NewSource(pos, lambda t: 2.0 * t)
    expect = np.empty(n, dtype=np.complex128)
    src.Eval(opfutil.NewFFTW(dims).Freq, 0.7, expect)
    out = np.empty(n, dtype=np.complex128)
    p = np.array(pos)
    _lib.check(_lib.lib().gopf_source_eval(rank, k.ctypes.data_as(PD), ctypes.c_int64(n), p.ctypes.data_as(PD),
                                           ctypes.c_double(1.4), out.ctypes.data_as(PD)))
    assert np.max(np.abs(out - expect)) < 1e-14


def _model(n=64):
    m = gpf.NewModel()
    f = gpf.NewField("density", n)
    m.AddField(f)
    return m, f


def test_add_source_follows_reference_rules():
    m, _ = _model()
    with pytest.raises(gpf.GopfError, match="index out of range"):  # AllSources[0] does not exist yet (model.go:153)
        m.AddSource(0, gpf.NewSource([1.0, 2.0], lambda t: 1.0))
    m.AddEquation("ddensity/dt = LAP density")
    m.AddSource(0, gpf.NewSource([1.0, 2.0], lambda t: 1.0))
    m.AddSource(0, gpf.NewSource([3.0, 2.0], lambda t: t))
    m.Init()
    rhs = m.RHS[0]  # sources are not RHS terms (model.go:283-294): the bilinear term went to Denum
    assert (len(rhs.Terms), len(rhs.Denum)) == (0, 1)
    with pytest.raises(gpf.GopfError, match="index out of range"):
        m.AddSource(1, gpf.NewSource([1.0, 2.0], lambda t: 1.0))
    for _ in range(6):
        m.AddSource(0, gpf.NewSource([0.0, 0.0], lambda t: 0.0))
    with pytest.raises(gpf.GopfError, match="at most"):
        m.AddSource(0, gpf.NewSource([0.0, 0.0], lambda t: 0.0))


def test_charge_transport_registration_checks():
    m, _ = _model(64)
    with pytest.raises(gpf.GopfError, match="Voigt"):
        m.RegisterExplicitTerm("CT", gpf.ChargeTransport(lambda i: np.ones((len(i), 4)), [1.0, 0.0], "density"))
    with pytest.raises(gpf.GopfError, match="explicit"):
        m.RegisterImplicitTerm("CT", gpf.ChargeTransport(lambda i: [1.0, 1.0, 0.0], [1.0, 0.0], "density"))
    with pytest.raises(gpf.GopfError, match="ExternalField"):
        m.RegisterExplicitTerm("CT", gpf.ChargeTransport(lambda i: [1.0, 1.0, 1.0, 0.0, 0.0, 0.0], [1.0, 0.0], "density"))
    m.RegisterExplicitTerm("MINUS_DIV_CURRENT", gpf.ChargeTransport(lambda i: [1.0, 1.0, 0.0], [1.0, 0.0], "density"))
    m.AddEquation("ddensity/dt = MINUS_DIV_CURRENT")
    m.Init()
    rhs = m.RHS[0]
    assert (len(rhs.Terms), len(rhs.Denum)) == (1, 0)
    # unknown field is reported at Init (the reference would panic on bricks[ct.Field] at the first step)
    m2, _ = _model(64)
    m2.RegisterExplicitTerm("CT", gpf.ChargeTransport(lambda i: [1.0, 1.0, 0.0], [1.0, 0.0], "nofield"))
    m2.AddEquation("ddensity/dt = CT")
    with pytest.raises(gpf.GopfError, match="unknown field"):
        m2.Init()
