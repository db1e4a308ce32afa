"""GPU parity of the epoch-boundary observers that read the device-resident state (SURVEY.md 8f
ranks 3 and 4): Uint8IO's payload (pf/fileIO.go:29-44, pf/util.go:108-117 -- the callback of
examples/cahnHilliard, config 1) and the free-energy observer of examples/pfcPhases (config 5):
IdealMixtureTerm.GetEnergy and PairCorrlationTerm.GetEnergy (pf/pairCorrelationTerm.go:58-84,
185-193).  Bytes are compared exactly, energies to 1e-10 relative."""
import numpy as np
import pytest

from gopf_b200 import pf as gpf
from gopf_b200 import synthetic, workloads
from oracle import pf as opf
from oracle import pfutil as opfutil
from oracle import terms as oterms

pytestmark = pytest.mark.gpu


def ch_model(mod, dims, seed=0):
    n = opfutil.prod_int(dims)
    m = mod.NewModel()
    f = mod.NewField("conc", n, synthetic.cahn_hilliard_initial(n, seed).copy())
    m.AddScalar(mod.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(mod.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    return m, f


@pytest.mark.parametrize("dims", [[128, 128], [32, 32, 32]], ids=lambda d: "x".join(map(str, d)))
def test_uint8_payload_is_byte_exact(dims, tmp_path):
    # examples/cahnHilliard/main.go:40-44: Uint8IO.SaveFields after every epoch
    m, f = ch_model(gpf, dims)
    s = gpf.NewSolver(m, dims, synthetic.CAHN_HILLIARD_DT)
    s.Upload()
    s.StepDevice(10)
    got, mn, mx = s.DownloadUint8(0)
    s.Download()  # the same device state on the host
    re = f.Data.real
    assert mn == float(re.min()) and mx == float(re.max())  # pfutil.MinReal / MaxReal
    expect = opf.uint8_payload(f.Data)
    assert got.dtype == np.uint8 and np.array_equal(got, expect)
    assert got.min() == 0 and got.max() == 255
    # the file written from the device equals the file written from host Field.Data
    gpf.NewUint8IO(str(tmp_path / "dev"), from_device=True).SaveFields(s, 3)
    gpf.NewUint8IO(str(tmp_path / "host")).SaveFields(s, 3)
    a = (tmp_path / "dev_conc_3.bin").read_bytes()
    b = (tmp_path / "host_conc_3.bin").read_bytes()
    assert a == b == expect.tobytes()


def test_uint8_constant_field_uses_unit_range():
    # pf/util.go:110-112: max - min < 1e-10 -> max = min + 1, every byte 0
    dims = [32, 32]
    m = gpf.NewModel()
    f = gpf.NewField("conc", 1024)
    f.Data[:] = 0.3
    m.AddField(f)
    m.AddEquation("dconc/dt = LAP conc")
    s = gpf.NewSolver(m, dims, 0.1)
    s.Upload()
    got, mn, mx = s.DownloadUint8(0)
    assert abs(mn - 0.3) < 1e-12 and abs(mx - 0.3) < 1e-12
    assert np.array_equal(got, np.zeros(1024, dtype=np.uint8))


@pytest.mark.parametrize("dims", [[32, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_pfc_energy_observer_vs_oracle(dims):
    # examples/pfcPhases/main.go:24-31: (ideal + excess) / N after every epoch
    n = opfutil.prod_int(dims)
    gm, gf, gs = workloads.build_pfc(gpf, gpf, dims, noise=None, filt_order=None)
    om, of, osolver = workloads.build_pfc(opf, oterms, dims, noise=None, filt_order=None)
    for epoch in range(2):
        gs.Solve(1, 5)
        osolver.Solve(1, 5)
        ideal_g = gm.MixedTerms["IDEAL"].GetEnergy(gm.Bricks, n)
        excess_g = gm.ImplicitTerms["EXCESS"].GetEnergy(gm.Bricks, None, dims)
        ideal_o = om.MixedTerms["IDEAL"].GetEnergy(om.Bricks, n)
        excess_o = om.ImplicitTerms["EXCESS"].GetEnergy(om.Bricks, osolver.FT, dims)
        assert abs(ideal_g - ideal_o) <= 1e-10 * max(1.0, abs(ideal_o))
        assert abs(excess_g - excess_o) <= 1e-10 * max(1.0, abs(excess_o))
        assert ideal_o != 0.0 and excess_o != 0.0


def test_energy_of_other_terms_is_refused():
    dims = [16, 16]
    m, f = ch_model(gpf, dims)
    m.RegisterExplicitTerm("SG", gpf.NewSquareGradient("conc", dims))
    s = gpf.NewSolver(m, dims, 0.1)
    s.Upload()
    import ctypes
    from gopf_b200._lib import lib
    e = ctypes.c_double(0.0)
    assert lib().gopf_solver_term_energy(s._h, b"SG", ctypes.byref(e)) != 0
    assert b"no energy" in lib().gopf_last_error()
    assert lib().gopf_solver_term_energy(s._h, b"NOPE", ctypes.byref(e)) != 0


def test_homogeneous_modulus_energy_vs_oracle_and_eshelby():
    # elasticity/linearElasticity.go:101-165; elasticity/linearElasticity_test.go:11-49 (Eshelby's
    # dilatational sphere within 5 % on a 64^3 grid)
    from gopf_b200 import elasticity as gel
    from oracle import elasticity as oel
    N = 32
    mis = np.array([[0.05, 0.01, 0.0], [0.01, -0.01, 0.02], [0.0, 0.02, 0.03]])
    ind = oel.Ellipsoid(N, 6.0, 4.0, 5.0)
    want = oel.HomogeneousModulusEnergy(ind, [N, N, N], mis, oel.CubicMaterial(110.0, 60.0, 30.0))
    got = gel.HomogeneousModulusEnergy(ind, [N, N, N], mis, gel.CubicMaterial(110.0, 60.0, 30.0))
    assert abs(got - want) <= 1e-10 * abs(want)
    # 2-D: the function still loops the force over three components and pads the frequencies (:114-132)
    i = np.arange(N * N)
    ind2 = ((((i // N) - N // 2) / 7.0) ** 2 + (((i % N) - N // 2) / 4.0) ** 2 <= 1.0).astype(np.complex128)
    want = oel.HomogeneousModulusEnergy(ind2, [N, N], mis, oel.Isotropic(60.0, 0.3))
    got = gel.HomogeneousModulusEnergy(ind2, [N, N], mis, gel.Isotropic(60.0, 0.3))
    assert abs(got - want) <= 1e-10 * abs(want)
    # the reference's own known answer
    N = 64
    poisson, bulk, eps = 0.3, 50.0, 0.05
    shear = oel.Shear(bulk, poisson)
    energy = gel.HomogeneousModulusEnergy(oel.Ellipsoid(N, 10.0, 10.0, 10.0), [N, N, N], np.diag([eps, eps, eps]),
                                          gel.Isotropic(bulk, poisson))
    expect = oel.EshelbyEnergyDensityDilatational(poisson, shear, eps)
    assert abs(energy - expect) < 0.05 * expect
