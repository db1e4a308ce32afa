"""The CUDA path (through the C ABI) against the committed golden vectors of tests/golden/ --
no oracle code runs in these tests.  Tolerance: relative L2 <= 1e-10 (BASELINE.json north_star);
the k-table is bit-exact."""
import os

import numpy as np
import pytest

from gopf_b200 import elasticity as gel
from gopf_b200 import pf as gpf
from gopf_b200 import pfutil as gpfutil
from gopf_b200 import synthetic, workloads

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("dims", [[8, 16], [8, 8, 8]], ids=lambda d: "x".join(map(str, d)))
def test_device_k_table_bit_exact(dims):
    g = load("ktable.npz")
    n = int(np.prod(dims))
    got = gpfutil.NewFFTW(dims).freq_device(np.arange(n))
    assert np.array_equal(got.reshape(n, -1)[:, :len(dims)], g[f"freq_{'x'.join(map(str, dims))}"])


def test_fft_ramp():
    g = load("fft_ramp_8x16.npz")
    x = g["input"].copy()
    gpfutil.NewFFTW([8, 16]).FFT(x)
    assert rel_l2(x, g["forward"]) < 1e-13


@pytest.mark.parametrize("generic", [False, True], ids=["fused", "generic"])
@pytest.mark.parametrize("name,dims,stepper", [("ch_2d_32x32_euler.npz", [32, 32], "euler"), ("ch_3d_16_euler.npz", [16, 16, 16], "euler"),
                                               ("ch_2d_32x32_rk4.npz", [32, 32], "rk4")])
def test_cahn_hilliard_trajectories(name, dims, stepper, generic):
    g = load(name)
    n = int(np.prod(dims))
    m = gpf.NewModel()
    f = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
    m.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    s = gpf.NewSolver(m, dims, synthetic.CAHN_HILLIARD_DT)
    s.SetStepper(stepper)
    if generic:
        s.ForceGeneric(True)
    done = 0
    for k in sorted(int(key.split("_")[1]) for key in g.files):
        s.Propagate(k - done)
        done = k
        assert rel_l2(f.Data, g[f"after_{k}"]) <= TOL


@pytest.mark.parametrize("dims", [[32, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_precipitate(dims):
    g = load(f"precipitate_{'x'.join(map(str, dims))}.npz")
    m, conc, phase, s, vol = workloads.build_precipitate(gpf, gpf, gel, dims, expressions=True)
    s.Solve(2, 5)
    assert rel_l2(conc.Data, g["conc"]) <= TOL and rel_l2(phase.Data, g["phase"]) <= TOL
    assert abs(s.LPMultiplier(0) - float(g["multiplier"][0])) <= 1e-9


def test_pfc():
    g = load("pfc_32x32_vandeven5.npz")
    m, f, s = workloads.build_pfc(gpf, gpf, [32, 32], noise=None, filt_order=5)
    s.Solve(2, 5)
    assert rel_l2(f.Data, g["density"]) <= TOL


# ---- ChargeTransport, point sources, SDD, epoch observers ---------------------
def test_charge_transport():
    g = load("charge_transport_32x32.npz")
    m, f, term, s = workloads.build_charge(gpf, gpf, [32, 32])
    s.Solve(2, 3)
    assert rel_l2(f.Data, g["density"]) <= TOL
    assert rel_l2(np.stack(term.Current()), g["current"]) <= TOL


@pytest.mark.parametrize("dims", [[16, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_point_sources(dims):
    g = load(f"sources_{'x'.join(map(str, dims))}.npz")
    m, f, s = workloads.build_sourced_diffusion(gpf, gpf, dims)
    s.Solve(2, 5)
    assert rel_l2(f.Data, g["conc"]) <= TOL


def test_sdd_nucleation():
    g = load("sdd_nucleation_64x64.npz")
    m, phi, sdd, s = workloads.build_sdd_nucleation(gpf, gpf.NewSDD, 64, expressions=True)
    s.Solve(1, 60)
    assert rel_l2(phi.Data, g["phi"]) <= TOL
    assert rel_l2(sdd.orientation, g["orientation"]) <= 1e-7  # conditioning ~ 1 / MinDimerLength
    mon = np.array([sdd.Monitor.MaxForce, sdd.Monitor.ForcePowerSpectrum, sdd.Monitor.MaxTorque, sdd.Monitor.FieldNorm,
                    sdd.Monitor.FieldNormChange])
    assert np.allclose(mon, g["monitor"], rtol=1e-7, atol=1e-9)


def test_observers():
    g = load("observers.npz")
    n = 32 * 32
    m = gpf.NewModel()
    f = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
    m.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    s = gpf.NewSolver(m, [32, 32], synthetic.CAHN_HILLIARD_DT)
    s.Upload()
    s.StepDevice(10)
    got, mn, mx = s.DownloadUint8(0)
    # the trajectory agrees to 1e-10, so a cell within that of a quantisation boundary may land on
    # either side: at most one level, on a handful of cells
    assert np.max(np.abs(got.astype(int) - g["ch_uint8"].astype(int))) <= 1
    assert np.mean(got != g["ch_uint8"]) < 0.01
    m, f, s = workloads.build_pfc(gpf, gpf, [32, 32], noise=None, filt_order=None)
    s.Solve(1, 5)
    e = np.array([m.MixedTerms["IDEAL"].GetEnergy(), m.ImplicitTerms["EXCESS"].GetEnergy()])
    assert np.allclose(e, g["pfc_energy"], rtol=1e-10, atol=0.0)


# ---- benchmark-scale fingerprints (tests/golden/make_golden_large.py) --------------
def _fingerprint_check(got, g, k):
    idx = g["indices"]
    ref = g[f"after_{k}_samples"]
    # relative L2 over 16 384 fixed cells: an unbiased estimate of the full-field figure
    assert rel_l2(got[idx], ref) <= TOL
    re = got.real
    n = re.size
    assert abs(float(np.sqrt(np.sum(re.astype(np.longdouble) ** 2))) - float(g[f"after_{k}_l2"])) <= 1e-11 * float(g[f"after_{k}_l2"])
    assert abs(float(re.sum(dtype=np.longdouble)) - float(g[f"after_{k}_sum"])) <= 1e-9 * n ** 0.5
    assert abs(float(np.max(np.abs(re))) - float(g[f"after_{k}_max_abs"])) <= 1e-10
    assert float(np.max(np.abs(got.imag))) <= 1e-11


@pytest.mark.parametrize("edge", [128, 256])
def test_cahn_hilliard_benchmark_grid_100_steps_vs_oracle_fingerprint(edge):
    """BASELINE.json configs[1] at its own size: the fused kernels (k_fused_kspace<256,16,LATE>,
    k_fused_real<256>, k_pass_strided<256,*>) against the oracle's state after 10 and 100 steps
    (pf/euler.go:16-47), through the fingerprint frozen by make_golden_large.py."""
    g = load(f"ch_3d_{edge}_fingerprint.npz")
    n = edge ** 3
    m = gpf.NewModel()
    f = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
    m.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    s = gpf.NewSolver(m, [edge] * 3, synthetic.CAHN_HILLIARD_DT)
    assert s.IsFused
    s.Upload()
    s.StepDevice(10)
    s.Download()
    _fingerprint_check(f.Data, g, 10)
    s.StepDevice(90)   # device state untouched by the read-back
    s.Download()
    _fingerprint_check(f.Data, g, 100)
