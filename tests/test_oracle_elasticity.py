"""Pins oracle/elasticity.py and the oracle's HomogeneousModulusLinElast against the
reference's own known-answer tests.  No GPU needed."""
import math

import numpy as np

from oracle import elasticity as el
from oracle import pf, pfutil, terms


def test_isotropic_rotation_invariance():
    # elasticity/rank4_test.go:11-40
    t = el.Isotropic(61.4, 0.3)
    data = t.Data.copy()
    for axis, angle in [(0, 14.0), (1, 56.0), (2, -56.0)]:
        t.Rotate(el.RotationMatrix(angle * math.pi / 180.0, axis))
        assert np.allclose(t.Data, data, atol=1e-10)


def test_cubic_material_symmetry():
    # elasticity/rank4_test.go:42-75: invariant under the identity and a 180 degree rotation,
    # not invariant under a generic rotation
    t = el.CubicMaterial(110.0, 60.0, 30.0)
    data = t.Data.copy()
    for rot in (np.eye(3), np.diag([-1.0, -1.0, 1.0])):
        c = el.Rank4(data)
        c.Rotate(rot)
        assert np.allclose(c.Data, data, atol=1e-10)
    c = el.Rank4(data)
    c.Rotate(el.RotationMatrix(0.3, 2))
    assert not np.allclose(c.Data, data, atol=1e-6)


def test_strain():
    # elasticity/linearElasticity_test.go:51-107
    N = 8
    i = np.arange(N * N)
    x = (i // N) / float(N)
    ux = np.power(x * (1.0 - x), 2).astype(np.complex128)
    expect = (2.0 * x * (1 - x) * (1 - x) - 2 * x * x * (1 - x)) / float(N)
    ft = pfutil.NewFFTW([N, N])
    ft.FFT(ux)
    disp = np.zeros((N * N, 2), dtype=np.complex128)
    disp[:, 0] = ux
    f = ft.freq_table()
    for (a, b), exp in [((0, 0), expect), ((0, 1), np.zeros(N * N)), ((1, 1), np.zeros(N * N))]:
        s = np.ascontiguousarray(el.Strain(disp, f, a, b))
        ft.IFFT(s)
        s /= N * N
        assert np.max(np.abs(s.real - exp)) < 1e-3 and np.max(np.abs(s.imag)) < 1e-3


def test_dilatational_misfit_eshelby_sphere():
    # elasticity/linearElasticity_test.go:11-49, first case (sphere), 5 % tolerance
    N, poisson, bulk, eps = 64, 0.3, 50.0, 0.05
    shear = el.Shear(bulk, poisson)
    mat_prop = el.Isotropic(bulk, poisson)
    misfit = np.diag([eps, eps, eps])
    energy = el.HomogeneousModulusEnergy(el.Ellipsoid(N, 10.0, 10.0, 10.0), [N, N, N], misfit, mat_prop)
    expect = el.EshelbyEnergyDensityDilatational(poisson, shear, eps)
    assert abs(energy - expect) < 0.05 * expect


def test_indicator_deriv():
    # pf/homoLinElast_test.go:12-25
    dx = 0.01
    for i in range(100):
        x = dx * i
        d = (terms.Indicator(x + dx / 2.0) - terms.Indicator(x - dx / 2.0)) / dx
        assert abs(d - terms.IndicatorDeriv(x)) < 1e-4


def test_homogeneous_rhs():
    # pf/homoLinElast_test.go:27-88 (displacements stubbed with a known u_x)
    N = 16
    mat_prop = el.Isotropic(60.0, 0.3)
    eps = 0.01
    misfit = np.zeros((3, 3))
    misfit[0, 0] = eps
    i = np.arange(N * N)
    x = (i // N) / float(N)
    ux = np.power(x * (1.0 - x), 2).astype(np.complex128)
    exx = 2 * x * (1 - x) * (1 - 2 * x) / float(N)
    h = terms.NewHomogeneousModolus("x", [N, N], mat_prop, misfit)
    h.FT.FFT(ux)

    def disps(force, freq, mp):
        res = np.zeros((N * N, 3), dtype=np.complex128)
        res[:, 0] = ux
        return res

    h.Disps = disps
    eta = 0.5
    f = pf.NewField("elasticity", N * N)
    f.Data[:] = eta
    h.Field[:] = eta
    C = mat_prop.At(0, 0, 0, 0)
    expect = C * eps * (exx - eps * terms.Indicator(eta)) * terms.IndicatorDeriv(eta)
    res = np.zeros(N * N, dtype=np.complex128)
    h.Construct({"elasticity": f})(h.FT.Freq, 0.0, res)
    h.FT.IFFT(res)
    res /= N * N
    assert np.max(np.abs(res.real - expect)) < 1e-4 and np.max(np.abs(res.imag)) < 1e-4
