"""Host-side pieces of the elastic term (no GPU): the elasticity helpers of the C ABI against
the oracle's restatement of elasticity/rank4.go, and the collapsed Khachaturyan multiplier
M(k) (gopf_b200/csrc/elastic.cuh) against the oracle's literal Force -> Displacements ->
Strain chain (elasticity/effectiveForce.go:26-35, linearElasticity.go:33-83,
pf/homoLinElast.go:59-86)."""
import math

import numpy as np
import pytest

from gopf_b200 import elasticity as gel
from oracle import elasticity as oel
from oracle import pfutil as opfutil


def test_tensor_helpers_match_oracle():
    for o, g in ((oel.CubicMaterial(110.0, 60.0, 30.0), gel.CubicMaterial(110.0, 60.0, 30.0)),
                 (oel.Isotropic(61.4, 0.3), gel.Isotropic(61.4, 0.3))):
        assert np.array_equal(o.Data, g.Data)
        t = np.array([[0.05, 0.01, 0.0], [0.01, -0.01, 0.002], [0.0, 0.002, 0.03]])
        assert np.allclose(o.ContractLast(t), g.ContractLast(t), rtol=0, atol=1e-14)
        assert abs(oel.EnergyDensity(o, t) - gel.EnergyDensity(g, t)) < 1e-14
        rot = oel.RotationMatrix(0.3, 2)
        o.Rotate(rot)
        g.Rotate(rot)
        assert np.allclose(o.Data, g.Data, rtol=0, atol=1e-12)


def test_isotropic_rotation_invariance():
    # elasticity/rank4_test.go:11-40 through the C ABI
    t = gel.Isotropic(61.4, 0.3)
    data = t.Data.copy()
    for axis, angle in [(0, 14.0), (1, 56.0), (2, -56.0)]:
        t.Rotate(oel.RotationMatrix(angle * math.pi / 180.0, axis))
        assert np.allclose(t.Data, data, atol=1e-10)


@pytest.mark.parametrize("dims", [[16, 16], [8, 16], [8, 8, 8]], ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("material", ["cubic", "isotropic", "rotated"])
def test_khachaturyan_multiplier_equals_literal_chain(dims, material):
    dim = len(dims)
    n = opfutil.prod_int(dims)
    f3 = oel.pad3(opfutil.NewFFTW(dims).freq_table())
    if dim == 3:
        misfit = np.array([[0.05, 0.01, 0.0], [0.01, -0.01, 0.002], [0.0, 0.002, 0.03]])
    else:
        misfit = np.array([[0.05, 0.01, 0.0], [0.01, -0.01, 0.0], [0.0, 0.0, 0.0]])
    if material == "cubic":
        co, cg = oel.CubicMaterial(110.0, 60.0, 30.0), gel.CubicMaterial(110.0, 60.0, 30.0)
    elif material == "isotropic":
        co, cg = oel.Isotropic(60.0, 0.3), gel.Isotropic(60.0, 0.3)
    else:
        co, cg = oel.CubicMaterial(110.0, 60.0, 30.0), gel.CubicMaterial(110.0, 60.0, 30.0)
        rot = oel.RotationMatrix(0.4, 2)
        co.Rotate(rot)
        cg.Rotate(rot)
    rng = np.random.default_rng(3)
    H = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    eff = oel.EffectiveForce(co, misfit)
    force = np.zeros((n, 3), dtype=np.complex128)
    for i in range(dim):  # pf/homoLinElast.go:114-127: components i < Dim only
        force[:, i] = eff.Get(i, f3, H)
    disp = oel.Displacements(force, f3, co)
    A = co.ContractLast(misfit)
    total = np.zeros(n, dtype=np.complex128)
    for i in range(dim):
        for j in range(i, dim):
            total += (1.0 if i == j else 2.0) * A[i, j] * oel.Strain(disp, f3, i, j)
    M = gel.KhachaturyanMultiplier(cg, misfit, dim, f3)
    assert M[0] == 0.0  # zero mode (linearElasticity.go:43-46)
    assert np.max(np.abs(M * H - total)) <= 1e-13 * np.max(np.abs(total))


@pytest.mark.parametrize("dims", [[8, 8, 8], [8, 16]], ids=lambda d: "x".join(map(str, d)))
def test_strain_factor_matches_literal_displacement_strain_chain(dims):
    # s_ij(k) H^ (csrc/elastic_energy.cu) against EffectiveForce -> Displacements -> Strain
    # (elasticity/effectiveForce.go:26-35, linearElasticity.go:16-83) on a random H^, all six
    # components, force components looped over comp < 3 as HomogeneousModulusEnergy does (:114-119)
    from gopf_b200 import elasticity as gel
    f = opfutil.NewFFTW(dims).freq_table()
    f3 = oel.pad3(f)
    n = f3.shape[0]
    C = oel.CubicMaterial(110.0, 60.0, 30.0)
    mis = np.array([[0.05, 0.01, 0.0], [0.01, -0.01, 0.02], [0.0, 0.02, 0.03]])
    rng = np.random.default_rng(9)
    H = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    eff = oel.EffectiveForce(C, mis)
    force = np.stack([eff.Get(c, f, H) for c in range(3)], axis=1)
    disp = oel.Displacements(force, f3, C)
    gC = gel.CubicMaterial(110.0, 60.0, 30.0)
    for i in range(3):
        for j in range(i, 3):
            want = oel.Strain(disp, f3, i, j)
            got = gel.StrainFactor(gC, mis, f3, i, j) * H
            assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want))), (i, j)
