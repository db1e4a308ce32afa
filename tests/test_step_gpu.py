"""GPU parity of the step-level path (pf.Model / pf.Solver over the C ABI) against the
oracle: same seeded inputs, sizes the oracle finishes in seconds.  Tolerance from
BASELINE.json north_star: relative L2 field error <= 1e-10 in fp64 after 100 steps.
Tests that restate a reference test cite it.
"""
import math

import numpy as np
import pytest

from gopf_b200 import pf as gpf
from gopf_b200 import synthetic
from oracle import pf as opf
from oracle import pfutil as opfutil
from oracle import terms as oterms

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def ch_models(dims, seed=0):
    n = opfutil.prod_int(dims)
    init = synthetic.cahn_hilliard_initial(n, seed)
    assert np.array_equal(init.real, 2.0 * opfutil.splitmix64_uniform(seed, n) - 1.0)
    out = []
    for mod in (gpf, opf):
        m = mod.NewModel()
        f = mod.NewField("conc", n, init.copy())
        m.AddScalar(mod.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
        m.AddScalar(mod.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
        m.AddField(f)
        m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
        out.append((m, f))
    return out


@pytest.mark.parametrize("dims", [[128, 128], [64, 256], [32, 32, 32], [64, 64, 64]], ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("generic", [False, True], ids=["fused", "generic"])
def test_cahn_hilliard_100_steps(dims, generic):
    # cfg 1 (examples/cahnHilliard/main.go) and small instances of cfg 2
    (gm, gf), (om, of) = ch_models(dims)
    gs = gpf.NewSolver(gm, dims, synthetic.CAHN_HILLIARD_DT)
    if generic:
        gs.ForceGeneric(True)
    assert gs.IsFused == (not generic)
    osolver = opf.NewSolver(om, dims, synthetic.CAHN_HILLIARD_DT)
    gs.Solve(10, 10)  # the reference example's 10 epochs x 10 steps
    osolver.Solve(10, 10)
    assert rel_l2(gf.Data, of.Data) <= TOL
    assert abs(gs.Stepper.GetTime() - osolver.Stepper.GetTime()) < 1e-12
    assert gs.KernelLaunches() > 0


def test_epoch_split_equals_single_run():
    # device-resident stepping between callbacks must equal host round trips every epoch
    dims = [64, 64]
    (gm1, gf1), _ = ch_models(dims)
    (gm2, gf2), _ = ch_models(dims)
    s1 = gpf.NewSolver(gm1, dims, 0.1)
    s1.Solve(5, 4)
    s2 = gpf.NewSolver(gm2, dims, 0.1)
    s2.Upload()
    s2.StepDevice(20)
    s2.Download()
    assert rel_l2(gf1.Data, gf2.Data) < 1e-13


def decay(mod, eq, N=8, c0=1.0):
    field = mod.NewField("field", N * N)
    field.Data[:] = c0
    model = mod.NewModel()
    model.AddField(field)
    model.AddScalar(mod.Scalar("rate", -1.0))
    model.AddEquation(eq)
    return model, field


def test_euler_exponential_decay():
    # pf/euler_test.go:10-49
    model, field = decay(gpf, "dfield/dt = rate*field")
    s = gpf.NewSolver(model, [8, 8], 0.001)
    s.Propagate(1000)
    assert np.all(np.abs(field.Data.real - math.exp(-1.0)) < 1e-3) and np.all(np.abs(field.Data.imag) < 1e-3)
    assert abs(s.Stepper.GetTime() - 1.0) < 1e-10


def test_euler_square_decay():
    # pf/euler_test.go:51-85
    model, field = decay(gpf, "dfield/dt = rate*field^2")
    s = gpf.NewSolver(model, [8, 8], 0.001)
    s.Propagate(1000)
    assert np.all(np.abs(field.Data.real - 0.5) < 1e-3) and np.all(np.abs(field.Data.imag) < 1e-3)


def test_rk4_simple_model_and_implicit():
    # pf/rk4_test.go:14-55
    model, field = decay(gpf, "dfield/dt = rate*field^2")
    s = gpf.NewSolver(model, [8, 8], 0.1)
    s.SetStepper("rk4")
    s.Propagate(10)
    assert np.all(np.abs(field.Data.real - 0.5) < 1e-6) and np.all(np.abs(field.Data.imag) < 1e-6)
    # pf/rk4_test.go:62-100
    c0 = 0.5
    model, field = decay(gpf, "dfield/dt = field + rate*field^2", c0=c0)
    s = gpf.NewSolver(model, [8, 8], 0.01)
    s.SetStepper("rk4")
    s.Propagate(100)
    expect = math.exp(1.0) / ((1.0 / c0 - 1.0) + math.exp(1.0))
    assert np.all(np.abs(field.Data.real - expect) < 1e-3)
    with pytest.raises(gpf.GopfError, match="Unknown stepper scheme"):
        s.SetStepper("leapfrog")  # pf/solver.go:100-102


@pytest.mark.parametrize("dims", [[32, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_rk4_cahn_hilliard_vs_oracle(dims):
    (gm, gf), (om, of) = ch_models(dims, seed=3)
    gs = gpf.NewSolver(gm, dims, 0.05)
    gs.SetStepper("rk4")
    osolver = opf.NewSolver(om, dims, 0.05)
    osolver.SetStepper("rk4")
    gs.Solve(2, 10)
    osolver.Solve(2, 10)
    assert rel_l2(gf.Data, of.Data) <= TOL


def test_solver_diffusion():
    # pf/solver_test.go:9-34
    m = gpf.NewModel()
    conc = gpf.NewField("conc", 16 * 16)
    conc.Data[128] = 1.0
    m.AddField(conc)
    m.AddEquation("dconc/dt = LAP conc")
    solver = gpf.NewSolver(m, [16, 16], 0.1)
    solver.Solve(10, 10)
    assert abs(conc.Data.real.sum() - 1.0) < 1e-4
    assert np.all(conc.Data.real < 1.0) and np.all(conc.Data.real >= -1e-15)
    with pytest.raises(gpf.GopfError, match="Inconsistent domain size"):
        gpf.NewSolver(m, [16, 8], 0.1)  # pf/solver.go:55-60


def test_gauss_seidel_field_ordering():
    # pf/euler.go:27-39: equation i+1 sees field i's UPDATED spectrum
    N = 8
    outs = []
    for mod in (gpf, opf):
        a, b = mod.NewField("aa", N * N), mod.NewField("bb", N * N)
        a.Data[:] = 1.0
        m = mod.NewModel()
        m.AddField(a)
        m.AddField(b)
        m.AddScalar(mod.NewScalar("rate", -1.0))
        m.AddScalar(mod.NewScalar("one", 1.0))
        m.AddEquation("daa/dt = rate*aa")
        m.AddEquation("dbb/dt = one*aa")
        s = mod.NewSolver(m, [N, N], 0.5)
        s.Propagate(1)
        outs.append((a.Data.copy(), b.Data.copy()))
    assert np.allclose(outs[0][0].real, 1.0 / 1.5, atol=1e-13)
    assert np.allclose(outs[0][1].real, 0.5 / 1.5, atol=1e-13)
    assert rel_l2(outs[0][0], outs[1][0]) < 1e-13 and rel_l2(outs[0][1], outs[1][1]) < 1e-13


@pytest.mark.parametrize("dims", [[32, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_reaction_diffusion_three_fields(dims):
    # the model of pf/model_test.go:45-108 run through both solvers
    n = opfutil.prod_int(dims)
    outs = []
    for mod in (gpf, opf):
        m = mod.NewModel()
        fields = []
        for k, name in enumerate(["concA", "concB", "concC"]):
            f = mod.NewField(name, n, (0.5 + 0.4 * opfutil.splitmix64_uniform(10 + k, n)).astype(np.complex128))
            m.AddField(f)
            fields.append(f)
        m.AddScalar(mod.NewScalar("kf", 2.0))
        m.AddScalar(mod.NewScalar("kr", 0.2))
        m.AddEquation("dconcA/dt = LAP concA - kf*concA^2*concB^3 + kr*concC")
        m.AddEquation("dconcB/dt = LAP concB - kf*concA^2*concB^3 + kr*concC")
        m.AddEquation("dconcC/dt = LAP concC - kr*concC + kf*concA^2*concB^3")
        s = mod.NewSolver(m, dims, 0.01)
        s.Solve(2, 25)
        outs.append([f.Data.copy() for f in fields])
    for a, b in zip(*outs):
        assert rel_l2(a, b) <= TOL


def sin_x(nx, ny):
    i = np.arange(nx * ny)
    x = (i % nx) / float(nx)
    return np.sin(2.0 * math.pi * x), np.power(2.0 * math.pi * np.cos(2.0 * math.pi * x) / float(nx), 2.0)


def test_square_grad_with_solver():
    # pf/squareGradientTerm_test.go:84-128
    N = 16
    model = gpf.NewModel()
    field1, field2 = gpf.NewField("field1", N * N), gpf.NewField("field2", N * N)
    data, grad_sq = sin_x(N, N)
    field2.Data[:] = data
    grad = gpf.NewSquareGradient("field2", [N, N])
    model.AddField(field1)
    model.AddField(field2)
    model.AddScalar(gpf.NewScalar("ZERO", 0.0))
    model.RegisterExplicitTerm("GRAD_SQ_f2", grad, None)
    model.AddEquation("dfield1/dt = GRAD_SQ_f2")
    model.AddEquation("dfield2/dt = ZERO*field1")
    dt, nsteps = 0.1, 10
    solver = gpf.NewSolver(model, [N, N], dt)
    solver.Solve(1, nsteps)
    assert np.max(np.abs(field2.Data.real - data)) < 1e-10 and np.max(np.abs(field2.Data.imag)) < 1e-10
    assert np.max(np.abs(field1.Data.real - dt * nsteps * grad_sq)) < 1e-10
    assert np.max(np.abs(field1.Data.imag)) < 1e-10


@pytest.mark.parametrize("dims", [[32, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_cahn_hilliard_plus_squared_gradient_vs_oracle(dims):
    # cfg-2 variant of SURVEY 8d: SquaredGradient (Factor = 1) added as explicit term
    n = opfutil.prod_int(dims)
    init = 0.1 * synthetic.cahn_hilliard_initial(n, 5)
    outs = []
    for mod, tmod in ((gpf, gpf), (opf, oterms)):
        m = mod.NewModel()
        f = mod.NewField("conc", n, init.copy())
        m.AddScalar(mod.NewScalar("gamma", 2.0))
        m.AddScalar(mod.NewScalar("m1", -1.0))
        m.AddField(f)
        m.RegisterExplicitTerm("GRAD_SQ", tmod.NewSquareGradient("conc", dims), None)
        m.AddEquation("dconc/dt = LAP conc^3 + m1*LAP conc + m1*gamma*LAP^2 conc + GRAD_SQ")
        s = mod.NewSolver(m, dims, 0.05)
        s.Solve(2, 10)
        outs.append(f.Data.copy())
    assert rel_l2(outs[0], outs[1]) <= TOL


def pfc_models(dims, lap, filt_order=None, noise_table=None):
    # cfg 5: examples/pfcPhases/main.go:63-89 (pair correlation + ideal mixture), optional
    # Vandeven filter and prescribed white noise shared by oracle and device
    n = opfutil.prod_int(dims)
    a = 16.0
    init = (0.3 * (2.0 * opfutil.splitmix64_uniform(11, n) - 1.0)).astype(np.complex128)
    res = []
    for mod, tmod in ((gpf, gpf), (opf, oterms)):
        m = mod.NewModel()
        f = mod.NewField("density", n, init.copy())
        m.AddField(f)
        peaks = [tmod.Peak(1.0, 2.0 * math.pi / a, 0.02, 4),
                 tmod.Peak(1.0 / math.sqrt(2.0), 2.0 * math.pi / (a / math.sqrt(2.0)), 0.02, 4)]
        term = tmod.PairCorrlationTerm(tmod.ReciprocalSpacePairCorrelation(0.1, peaks), "density", 1.0, lap)
        ideal = tmod.IdealMixtureTerm(tmod.IdealMix(1.0, 1.0), "density", 1.0, lap)
        m.RegisterImplicitTerm("EXCESS", term, None)
        m.RegisterMixedTerm("IDEAL", ideal, [ideal.DerivedField(n, m.Bricks)])
        eq = "ddensity/dt = IDEAL + EXCESS"
        if noise_table is not None:
            if mod is gpf:
                m.RegisterTableField("NOISE", noise_table)
            else:
                # Model.Init evaluates Calc once before the first step (model.go:249)
                m.RegisterDerivedField(opf.DerivedField(np.zeros(n, dtype=np.complex128), "NOISE", None))
                d = m.DerivedFields[-1]
                d._k = -1

                def calc(out, d=d):
                    out[:] = noise_table[d._k % noise_table.shape[0]] if d._k >= 0 else 0.0
                    d._k += 1

                d.Calc = calc
            eq += " + NOISE"
        m.AddEquation(eq)
        s = mod.NewSolver(m, dims, 0.1)
        if filt_order is not None:
            s.Stepper.SetFilter(tmod.NewVandeven(filt_order))
        res.append((m, f, s))
    return res


@pytest.mark.parametrize("lap", [False, True], ids=["nolap", "lap"])
def test_pfc_pair_correlation_ideal_mixture_vs_oracle(lap):
    dims = [64, 64]
    (gm, gf, gs), (om, of, osolver) = pfc_models(dims, lap)
    # without the Laplacian the linear part grows like (1 - dt)^-n: keep that run short
    nsteps = 25 if lap else 5
    gs.Solve(2, nsteps)
    osolver.Solve(2, nsteps)
    assert np.all(np.isfinite(of.Data))
    assert rel_l2(gf.Data, of.Data) <= TOL


@pytest.mark.parametrize("dims", [[64, 64], [32, 32, 32]], ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("knoise", [False, True], ids=["nonoise", "knoise"])
def test_pfc_tabulated_fused_form_vs_general_path_and_oracle(dims, knoise):
    """cfg 5 on the fused kernels: implicit pair-correlation + ideal-mixture terms and the Vandeven filter tabulated
    once per k-point, the ideal-mixture polynomial evaluated by Horner, white noise drawn at the k-point.  The
    general path interprets the same program and draws the same Philox stream, so the two must agree to rounding;
    without noise the oracle (pf/euler.go:16-47, pf/pairCorrelationTerm.go:37-51, pf/vandeven.go:30-40) pins both."""
    from gopf_b200 import workloads
    if knoise and not gpf.HasKSpaceNoise():
        pytest.skip("library built without -DGOPF_KNOISE")
    kw = dict(noise="device" if knoise else None, filt_order=5, kspace_noise=knoise)
    m, f, s = workloads.build_pfc(gpf, gpf, dims, **kw)
    assert s.IsFused and s.FusedForm() == (2, 2)
    s.Solve(2, 5)
    fused = f.Data.copy()
    m2, f2, s2 = workloads.build_pfc(gpf, gpf, dims, **kw)
    s2.ForceGeneric(True)
    assert not s2.IsFused
    s2.Solve(2, 5)
    assert rel_l2(fused, f2.Data) <= 1e-11
    if not knoise:
        om, of, osolver = workloads.build_pfc(opf, oterms, dims, noise=None, filt_order=5)
        osolver.Solve(2, 5)
        assert rel_l2(fused, of.Data) <= TOL


def test_pfc_with_vandeven_filter_and_prescribed_noise_vs_oracle():
    # cfg 5 additions: white noise injected from a shared array, Vandeven(5)
    dims = [32, 32, 32]
    n = 32 ** 3
    rng = np.random.default_rng(42)
    noise = math.sqrt(2.0 * 1e-4) * rng.standard_normal((20, n))
    (gm, gf, gs), (om, of, osolver) = pfc_models(dims, True, filt_order=5, noise_table=noise)
    assert not gs.IsFused  # two derived fields -> general path
    gs.Solve(2, 10)
    osolver.Solve(2, 10)
    assert rel_l2(gf.Data, of.Data) <= TOL


def test_spectral_viscosity_vs_oracle():
    dims = [32, 32]
    n = 1024
    init = synthetic.cahn_hilliard_initial(n, 2)
    outs = []
    for mod, tmod in ((gpf, gpf), (opf, oterms)):
        m = mod.NewModel()
        f = mod.NewField("conc", n, init.copy())
        m.AddField(f)
        m.AddScalar(mod.NewScalar("m1", -1.0))
        m.RegisterImplicitTerm("SPECTRAL_VISC", tmod.SpectralViscosity(0.5, 0.25, 2), None)
        m.AddEquation("dconc/dt = LAP conc^3 + m1*LAP conc + SPECTRAL_VISC")
        s = mod.NewSolver(m, dims, 0.01)
        s.Solve(1, 50)
        outs.append(f.Data.copy())
    assert rel_l2(outs[0], outs[1]) <= TOL


def test_white_noise_statistics():
    # pf/noise_test.go:23-48 on the device stream: std = sqrt(2*Strength) = 2 for Strength 2
    N = 512
    m = gpf.NewModel()
    f = gpf.NewField("price", N * N)
    m.AddField(f)
    m.RegisterFunction("WHITE_NOISE", gpf.WhiteNoise(2.0, seed=7).Generate)
    m.AddEquation("dprice/dt = WHITE_NOISE")
    s = gpf.NewSolver(m, [N, N], 1.0)
    s.Propagate(1)  # price = dt * noise
    x = f.Data.real
    assert abs(np.std(x, ddof=1) - 2.0) < 0.01 and abs(np.mean(x)) < 0.02
    assert np.max(np.abs(f.Data.imag)) < 1e-9
    first = x.copy()
    s.Propagate(1)
    inc = f.Data.real - first
    assert abs(np.corrcoef(first, inc)[0, 1]) < 0.01  # fresh draws every step


def test_two_white_noise_fields_are_independent():
    # the reference draws every WhiteNoise from the shared math/rand stream (pf/noise.go:20-23): two noise
    # fields registered with the same (default) seed must not be the same field
    N = 256
    m = gpf.NewModel()
    a = gpf.NewField("a", N * N)
    b = gpf.NewField("b", N * N)
    m.AddField(a)
    m.AddField(b)
    m.RegisterFunction("NOISE_A", gpf.WhiteNoise(0.5).Generate)
    m.RegisterFunction("NOISE_B", gpf.WhiteNoise(0.5).Generate)
    m.AddEquation("da/dt = NOISE_A")
    m.AddEquation("db/dt = NOISE_B")
    s = gpf.NewSolver(m, [N, N], 1.0)
    s.Propagate(1)
    x, y = a.Data.real, b.Data.real
    assert abs(np.std(x, ddof=1) - 1.0) < 0.02 and abs(np.std(y, ddof=1) - 1.0) < 0.02
    assert abs(np.corrcoef(x, y)[0, 1]) < 0.02


def test_conservative_noise_properties():
    # pf/noise_test.go:50-94: field stays real and its integral stays zero
    N = 16
    m = gpf.NewModel()
    field = gpf.NewField("myfield", N * N)
    m.AddField(field)
    noise = gpf.NewConservativeNoise(1.0, 2, unique_prefix=1234, seed=3)
    m.RegisterExplicitTerm("CONSERVATIVE_NOISE", noise, noise.RequiredDerivedFields(N * N))
    m.AddEquation("dmyfield/dt = CONSERVATIVE_NOISE")
    solver = gpf.NewSolver(m, [N, N], 0.1)
    solver.Solve(10, 100)
    assert np.max(np.abs(field.Data.imag)) < 1e-10
    assert abs(field.Data.real.sum()) < 1e-10
    assert np.count_nonzero(np.abs(field.Data.real) > math.sqrt(2.0)) > 0


def test_conservative_noise_prescribed_currents_vs_oracle():
    N = 16
    n = N * N
    rng = np.random.default_rng(9)
    tabs = [rng.standard_normal((8, n)) for _ in range(2)]
    outs = []
    for mod in (gpf, opf):
        m = mod.NewModel()
        field = mod.NewField("myfield", n)
        m.AddField(field)
        if mod is gpf:
            for c in range(2):
                m.RegisterTableField(f"77_current_{c}", tabs[c])
            m.RegisterExplicitTerm("CONSERVATIVE_NOISE", gpf.NewConservativeNoise(1.0, 2, unique_prefix=77), None)
        else:
            noise = oterms.ConservativeNoise(1.0, 2, unique_prefix=77)
            dfs = []
            for c in range(2):
                d = opf.DerivedField(np.zeros(n, dtype=np.complex128), noise.GetCurrentName(c), None)
                d._k = -1

                def calc(out, d=d, c=c):
                    out[:] = tabs[c][d._k % 8] if d._k >= 0 else 0.0
                    d._k += 1

                d.Calc = calc
                dfs.append(d)
            m.RegisterExplicitTerm("CONSERVATIVE_NOISE", noise, dfs)
        m.AddEquation("dmyfield/dt = CONSERVATIVE_NOISE")
        s = mod.NewSolver(m, [N, N], 0.1)
        s.Solve(1, 8)
        outs.append(field.Data.copy())
    assert rel_l2(outs[0], outs[1]) <= TOL


@pytest.mark.parametrize("stepper", ["euler", "rk4"])
def test_two_phase_functions_and_volume_constraint_vs_oracle(stepper):
    # the non-elastic part of cfg 4 (examples/strain_single_precipitate/main.go:45-111):
    # registered functions, kappa*LAP terms and the VolumeConservingLP constraint.  Under RK4 the
    # OnStepFinished hooks still run after every Stepper.Step (pf/solver.go:70-84).
    M = 32
    dims = [M, M]
    n = M * M
    A = B = W = 0.1
    dt = 0.1
    H = lambda x: 3.0 * x * x - 2.0 * x * x * x
    dH = lambda x: 6.0 * x - 6.0 * x * x
    dL = lambda x: 2.0 * x - 6.0 * x * x + 4.0 * x * x * x
    idx = np.arange(n)
    r, c = idx // M, idx % M
    inside = (r > 3 * M // 8) & (r < 5 * M // 8) & (c > 3 * M // 8) & (c < 5 * M // 8)
    outs = []
    mult = []
    for mod in (gpf, opf):
        m = mod.NewModel()
        conc = mod.NewField("conc", n, inside.astype(np.complex128))
        phase = mod.NewField("phase", n, inside.astype(np.complex128))
        m.AddScalar(mod.NewScalar("kappa", 0.1))
        m.AddField(conc)
        m.AddField(phase)
        if mod is gpf:
            m.RegisterFunction("CHEMICALPOT", "-((0.1*conc*(1.0-H(phase))-0.1*(1.0-conc)*H(phase))*1.0)")
            m.RegisterFunction("DERIV_PHASE_ORDER",
                               "-(-0.5*0.1*conc^2*dH(phase)+0.5*0.1*(1.0-conc)^2*dH(phase)+0.1*dLandau(phase))")
            m.RegisterFunction("SMEARING_DERIV", "dH(phase)")
            vol = gpf.NewVolumeConservingLP("phase", "SMEARING_DERIV", dt, n)
        else:
            cc = lambda i, b: np.real(b["conc"].Get(i))
            xx = lambda i, b: np.real(b["phase"].Get(i))
            m.RegisterFunction("CHEMICALPOT", lambda i, b: -((A * cc(i, b) * (1.0 - H(xx(i, b))) - B * (1.0 - cc(i, b)) * H(xx(i, b))) * 1.0) + 0j)
            m.RegisterFunction("DERIV_PHASE_ORDER", lambda i, b: -(-0.5 * A * cc(i, b) ** 2 * dH(xx(i, b)) + 0.5 * B * (1.0 - cc(i, b)) ** 2 * dH(xx(i, b)) + W * dL(xx(i, b))) + 0j)
            m.RegisterFunction("SMEARING_DERIV", lambda i, b: dH(xx(i, b)) + 0j)
            vol = oterms.NewVolumeConservingLP("phase", "SMEARING_DERIV", dt, n)
        m.RegisterExplicitTerm("CONSERVE_PREC_VOL", vol, None)
        m.AddEquation("dconc/dt = CHEMICALPOT + kappa*LAP conc")
        m.AddEquation("dphase/dt = DERIV_PHASE_ORDER + kappa*LAP phase + CONSERVE_PREC_VOL")
        s = mod.NewSolver(m, dims, dt)
        if stepper != "euler":
            s.SetStepper(stepper)
        if mod is gpf:
            s.Upload()
            s.StepDevice(30)  # hooks run on the device between steps
            s.Download()
            mult.append(s.LPMultiplier(0))
        else:
            s.Solve(3, 10)
            mult.append(vol.Multiplier)
        outs.append((conc.Data.copy(), phase.Data.copy()))
    assert rel_l2(outs[0][0], outs[1][0]) <= TOL and rel_l2(outs[0][1], outs[1][1]) <= TOL
    assert abs(mult[0] - mult[1]) <= 1e-9 * max(1.0, abs(mult[1]))
    # the multiplier is doing work: the constrained volume drifts less than the free one would
    assert mult[0] != 0.0


def test_cahn_hilliard_256_cubed_properties():
    # BASELINE.json cfg 2 at full size: size-independent properties instead of the oracle
    dims = [256, 256, 256]
    n = 256 ** 3
    m = gpf.NewModel()
    f = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
    m.AddScalar(gpf.NewScalar("gamma", 2.0))
    m.AddScalar(gpf.NewScalar("m1", -1.0))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    mass0 = f.Data.real.sum()
    s = gpf.NewSolver(m, dims, 0.1)
    assert s.IsFused
    s.Upload()
    s.StepDevice(20)
    s.Download()
    fused = f.Data.copy()
    # mass conservation: every RHS term carries a Laplacian, so the DC mode is invariant
    assert abs(fused.real.sum() - mass0) < 1e-6 * n ** 0.5
    assert np.max(np.abs(fused.imag)) < 1e-12 and np.all(np.isfinite(fused.real))
    assert np.max(np.abs(fused.real)) < 1.5  # bounded: the dynamics are dissipative
    # the general (unfused) path is an independent implementation of the same step
    f.Data[:] = synthetic.cahn_hilliard_initial(n, 0)
    s.ForceGeneric(True)
    s.Upload()
    s.StepDevice(20)
    s.Download()
    assert rel_l2(f.Data, fused) < 1e-12


@pytest.mark.parametrize("dims,steps", [([256, 256, 256], 10), ([512, 512, 512], 2), ([1024, 1024], 100), ([2048, 2048], 10),
                                        ([1024, 256], 20)], ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else str(v))
def test_cahn_hilliard_benchmark_scale_vs_oracle(dims, steps):
    """the fused kernels at the sizes that are benchmarked, against the oracle itself
    (pf/euler.go:16-47; scipy.fft with every host thread).  256^3 = BASELINE.json configs[1];
    512^3 = the cfg 4 / cfg 5 line length; 1024^2 and 2048^2 run k_fused_kspace<1024|2048> and
    k_fused_real<1024|2048>, the line kernels of the 1024^3 sharded arm (configs[2])."""
    import os
    n = opfutil.prod_int(dims)
    init = synthetic.cahn_hilliard_initial(n, 0)
    gm = gpf.NewModel()
    gf = gpf.NewField("conc", n, init.copy())
    gm.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    gm.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    gm.AddField(gf)
    gm.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    gs = gpf.NewSolver(gm, dims, synthetic.CAHN_HILLIARD_DT)
    assert gs.IsFused
    gs.Upload()
    gs.StepDevice(steps)
    gs.Download()
    got = gf.Data.copy()
    gs.close()
    del gs, gm, gf
    om = opf.NewModel()
    of = opf.NewField("conc", n, init)
    om.AddScalar(opf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    om.AddScalar(opf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    om.AddField(of)
    om.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    opf.NewSolver(om, dims, synthetic.CAHN_HILLIARD_DT, workers=os.cpu_count() or 1).Propagate(steps)
    assert rel_l2(got, of.Data) <= TOL


# ---- cfg 4: examples/strain_single_precipitate with HomogeneousModulusLinElast -----------
def _precipitate_pair(dims, **kw):
    from gopf_b200 import elasticity as gel
    from gopf_b200 import workloads
    from oracle import elasticity as oel
    g = workloads.build_precipitate(gpf, gpf, gel, dims, expressions=True, **kw)
    o = workloads.build_precipitate(opf, oterms, oel, dims, expressions=False, **kw)
    return g, o


@pytest.mark.parametrize("dims", [[64, 64], [32, 32, 32]], ids=lambda d: "x".join(map(str, d)))
def test_strain_single_precipitate_vs_oracle(dims):
    # the shipped example (2-D 64^2, main.go:66-127) and its 3-D extension (BASELINE.json cfg 4)
    (gm, gconc, gphase, gs, gvol), (om, oconc, ophase, osolver, ovol) = _precipitate_pair(dims)
    assert not gs.IsFused
    gs.Solve(4, 5)
    osolver.Solve(4, 5)
    assert rel_l2(gconc.Data, oconc.Data) <= TOL and rel_l2(gphase.Data, ophase.Data) <= TOL
    assert abs(gs.LPMultiplier(0) - ovol.Multiplier) <= 1e-9 * max(1.0, abs(ovol.Multiplier))


def test_elastic_term_vanishes_in_first_step_then_acts():
    # HomogeneousModulusLinElast.Field starts as zeros and is refreshed in OnStepFinished only
    # (pf/homoLinElast.go:130-134,145): step 0 must equal the run without the term, step 1 not
    dims = [32, 32]
    (gm, gconc, gphase, gs, _), _ = _precipitate_pair(dims, volume=False)
    (_, _, gphase0, gs0, _), _ = _precipitate_pair(dims, volume=False, elastic=False)
    gs.Solve(1, 1)
    gs0.Solve(1, 1)
    assert np.array_equal(gphase.Data, gphase0.Data)
    gs.Solve(1, 1)
    gs0.Solve(1, 1)
    assert rel_l2(gphase.Data, gphase0.Data) > 1e-6


def test_elastic_term_rk4_vs_oracle():
    dims = [32, 32]
    (gm, gconc, gphase, gs, _), (om, oconc, ophase, osolver, _) = _precipitate_pair(dims, volume=False)
    gs.SetStepper("rk4")
    osolver.SetStepper("rk4")
    gs.Solve(2, 3)
    osolver.Solve(2, 3)
    assert rel_l2(gconc.Data, oconc.Data) <= TOL and rel_l2(gphase.Data, ophase.Data) <= TOL


@pytest.mark.parametrize("dims,K", [([64, 64], [1.0, 0.3, 0.3, 2.0]), ([16, 16, 16], [1.0, 0.2, 0.1, 0.2, 2.0, 0.3, 0.1, 0.3, 0.5])],
                         ids=["2d", "3d"])
def test_tensorial_hessian_anisotropic_diffusion_vs_oracle(dims, K):
    # pf.TensorialHessian (pf/tensorialHessian.go:17-74) as the implicit part of a diffusion equation
    n = opfutil.prod_int(dims)
    init = synthetic.cahn_hilliard_initial(n, 5)
    outs = []
    for mod, tmod in ((gpf, gpf), (opf, oterms)):
        m = mod.NewModel()
        f = mod.NewField("conc", n, init.copy())
        m.AddField(f)
        m.RegisterImplicitTerm("HESSIAN", tmod.TensorialHessian(K), None)
        m.AddEquation("dconc/dt = HESSIAN - conc^3")
        s = mod.NewSolver(m, dims, 0.05)
        s.Solve(2, 10)
        outs.append(f.Data.copy())
    assert rel_l2(outs[0], outs[1]) <= TOL
    assert rel_l2(outs[0], init) > 0.1


@pytest.mark.parametrize("dims", [[128, 128], [32, 32, 32]], ids=lambda d: "x".join(map(str, d)))
def test_graph_replay_equals_eager_launches(dims):
    # small grids replay a captured CUDA graph of 8 fused steps; one step per call never does
    (gm1, gf1), _ = ch_models(dims)
    (gm2, gf2), _ = ch_models(dims)
    s1 = gpf.NewSolver(gm1, dims, 0.1)
    s1.Upload()
    s1.StepDevice(35)   # 1 eager + 4 graph launches + 2 eager
    s1.StepDevice(20)   # the graph is reused
    s1.Download()
    s2 = gpf.NewSolver(gm2, dims, 0.1)
    s2.Upload()
    for _ in range(55):
        s2.StepDevice(1)
    s2.Download()
    assert np.array_equal(gf1.Data, gf2.Data)
    assert abs(s1.Stepper.GetTime() - 5.5) < 1e-12 and abs(s2.Stepper.GetTime() - 5.5) < 1e-12
    assert s1.KernelLaunches() == s2.KernelLaunches()


def test_float64_io_from_device_matches_host_path(tmp_path):
    # pf.Float64IO.SaveFields (pf/fileIO.go:57-62): big-endian float64 of the real part; the
    # device-resident epoch loop writes the same bytes as the host path after the same steps
    dims = [32, 32, 32]
    (gm1, gf1), (om, of) = ch_models(dims)
    (gm2, gf2), _ = ch_models(dims)
    s1 = gpf.NewSolver(gm1, dims, 0.1)
    s1.AddCallback(gpf.NewFloat64IO(str(tmp_path / "host")).SaveFields)
    s1.Solve(3, 4)
    s2 = gpf.NewSolver(gm2, dims, 0.1)
    s2.AddCallback(gpf.NewFloat64IO(str(tmp_path / "dev"), from_device=True).SaveFields)
    s2.SolveOnDevice(3, 4)
    osolver = opf.NewSolver(om, dims, 0.1)
    osolver.Solve(3, 4)
    for epoch in range(3):
        a = gpf.LoadFloat64(str(tmp_path / f"host_conc_{epoch}.bin"))
        b = gpf.LoadFloat64(str(tmp_path / f"dev_conc_{epoch}.bin"))
        assert a.shape[0] == 32 ** 3 and rel_l2(a, b) < 1e-13
    raw = (tmp_path / "dev_conc_2.bin").read_bytes()
    assert raw == np.ascontiguousarray(s2.DownloadReal(0)).astype(">f8").tobytes()  # byte order: big endian
    assert rel_l2(gpf.LoadFloat64(str(tmp_path / "dev_conc_2.bin")), of.Data.real) <= TOL
    assert rel_l2(gf2.Data, of.Data) <= TOL


def test_negative_value_penalty_vs_oracle():
    # pf.NegativeValuePenalty (pf/negative_value_penalty.go) as a registered function: acts on the
    # negative part of the field only
    dims = [32, 32]
    n = 32 * 32
    init = synthetic.cahn_hilliard_initial(n, 9) * 0.2   # values in [-0.2, 0.2)
    outs = []
    for mod, tmod in ((gpf, gpf), (opf, oterms)):
        m = mod.NewModel()
        f = mod.NewField("density", n, init.copy())
        m.AddField(f)
        nvp = tmod.NegativeValuePenalty(5.0, 3, "density")
        m.RegisterFunction("PENALTY", nvp.Evaluate)
        m.AddEquation("ddensity/dt = LAP density - PENALTY")
        s = mod.NewSolver(m, dims, 0.01)
        s.Solve(2, 10)
        outs.append(f.Data.copy())
    assert rel_l2(outs[0], outs[1]) <= TOL
    assert np.mean(outs[0].real) > np.mean(init.real) + 1e-4   # the penalty pushed negative values up
