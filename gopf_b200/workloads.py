"""Workload definitions of BASELINE.json configs 4 and 5, written once against the ``pf``
surface so the same builder drives this package (``gopf_b200.pf``) and -- in tests and the CPU
baseline only -- the oracle restatement.  The caller passes the modules; nothing here imports
``oracle``.

cfg 4: examples/strain_single_precipitate/main.go:45-127 (two fields, three registered
functions, VolumeConservingLP, HomogeneousModulusLinElast), extended to 3-D as SURVEY.md 8d
states.
"""
from __future__ import annotations

import numpy as np

PRECIPITATE_DT = 0.1          # main.go:67
PRECIPITATE_KAPPA = 0.1       # main.go:86
PRECIPITATE_A = PRECIPITATE_B = PRECIPITATE_W = 0.1  # main.go:93-98
PRECIPITATE_M = 1.0
PRECIPITATE_MISFIT = np.array([[0.05, 0.0, 0.0], [0.0, -0.01, 0.0], [0.0, 0.0, 0.0]])  # main.go:109
PRECIPITATE_CUBIC = (110.0, 60.0, 30.0)  # main.go:110

# Device-expressible forms of ModelFunctions.ChemicalPotential / DerivPhase / SmearingDeriv
# (main.go:45-64); H, dH, dLandau are main.go:18-32.
CHEMICALPOT_EXPR = "-((0.1*conc*(1.0-H(phase))-0.1*(1.0-conc)*H(phase))*1.0)"
DERIV_PHASE_EXPR = "-(-0.5*0.1*conc^2*dH(phase)+0.5*0.1*(1.0-conc)^2*dH(phase)+0.1*dLandau(phase))"
SMEARING_EXPR = "dH(phase)"


def _H(x):
    return 3.0 * x * x - 2.0 * x * x * x


def _dH(x):
    return 6.0 * x - 6.0 * x * x


def _dLandau(x):
    return 2.0 * x - 6.0 * x * x + 4.0 * x * x * x


def precipitate_initial(dims) -> np.ndarray:
    """Cube (square in 2-D) 3M/8 < r, c[, d] < 5M/8 set to one (main.go:76-84)."""
    M = dims[0]
    n = int(np.prod(dims))
    idx = np.arange(n, dtype=np.int64)
    c = idx % dims[1]
    r = (idx // dims[1]) % dims[0]
    inside = (r > 3 * M // 8) & (r < 5 * M // 8) & (c > 3 * M // 8) & (c < 5 * M // 8)
    if len(dims) == 3:
        d = idx // (dims[0] * dims[1])
        inside &= (d > 3 * M // 8) & (d < 5 * M // 8)
    return inside.astype(np.complex128)


def build_precipitate(pf, terms, elasticity, dims, *, expressions: bool, elastic: bool = True, volume: bool = True,
                      pinned: bool = False):
    """Returns (model, conc, phase, solver, volume_term).  ``expressions`` selects the
    device-expressible RegisterFunction form (this package) or Python closures (the oracle)."""
    n = int(np.prod(dims))
    dt = PRECIPITATE_DT
    init = precipitate_initial(dims)
    m = pf.NewModel()
    if pinned:
        conc = pf.NewField("conc", n, None, pinned=True)
        phase = pf.NewField("phase", n, None, pinned=True)
        conc.Data[:] = init
        phase.Data[:] = init
    else:
        conc = pf.NewField("conc", n, init.copy())
        phase = pf.NewField("phase", n, init.copy())
    m.AddScalar(pf.NewScalar("kappa", PRECIPITATE_KAPPA))
    m.AddField(conc)
    m.AddField(phase)
    if expressions:
        m.RegisterFunction("CHEMICALPOT", CHEMICALPOT_EXPR)
        m.RegisterFunction("DERIV_PHASE_ORDER", DERIV_PHASE_EXPR)
        m.RegisterFunction("SMEARING_DERIV", SMEARING_EXPR)
    else:
        A, B, W, Mm = PRECIPITATE_A, PRECIPITATE_B, PRECIPITATE_W, PRECIPITATE_M
        cc = lambda i, b: np.real(b["conc"].Get(i))
        xx = lambda i, b: np.real(b["phase"].Get(i))
        m.RegisterFunction("CHEMICALPOT", lambda i, b: -((A * cc(i, b) * (1.0 - _H(xx(i, b))) - B * (1.0 - cc(i, b)) * _H(xx(i, b))) * Mm) + 0j)
        m.RegisterFunction("DERIV_PHASE_ORDER", lambda i, b: -(-0.5 * A * cc(i, b) ** 2 * _dH(xx(i, b)) + 0.5 * B * (1.0 - cc(i, b)) ** 2 * _dH(xx(i, b)) + W * _dLandau(xx(i, b))) + 0j)
        m.RegisterFunction("SMEARING_DERIV", lambda i, b: _dH(xx(i, b)) + 0j)
    vol = None
    eq_phase = "dphase/dt = DERIV_PHASE_ORDER"
    if volume:
        vol = terms.NewVolumeConservingLP("phase", "SMEARING_DERIV", dt, n)
        m.RegisterExplicitTerm("CONSERVE_PREC_VOL", vol, None)
    if elastic:
        mat_prop = elasticity.CubicMaterial(*PRECIPITATE_CUBIC)
        lin = terms.NewHomogeneousModolus("phase", dims, mat_prop, PRECIPITATE_MISFIT.copy())
        m.RegisterExplicitTerm("LIN_ELAST", lin, None)
        eq_phase += " + LIN_ELAST"
    eq_phase += " + kappa*LAP phase"
    if volume:
        eq_phase += " + CONSERVE_PREC_VOL"
    m.AddEquation("dconc/dt = CHEMICALPOT + kappa*LAP conc")
    m.AddEquation(eq_phase)
    solver = pf.NewSolver(m, dims, dt)
    return m, conc, phase, solver, vol


# ---- cfg 5: examples/pfcPhases/main.go:38-89 scaled to 3-D, + white noise + Vandeven filter ----
PFC_DT = 0.1
PFC_LATTICE = 16.0          # main.go:30 uses a = 16 lattice spacing units (flag default)
PFC_PEAK_WIDTH = 0.02       # main.go:66
PFC_EFF_TEMP = 0.1
PFC_NOISE_STRENGTH = 1e-4
PFC_SEED = 11


def pfc_initial(n_nodes: int, seed: int = PFC_SEED, mean_density: float = 0.0) -> np.ndarray:
    """density = 0.3 (2u - 1) + mean (main.go:57-60 with the synthetic stream)."""
    from . import synthetic
    out = np.empty(n_nodes, dtype=np.complex128)
    chunk = 1 << 24
    for s in range(0, n_nodes, chunk):
        m = min(chunk, n_nodes - s)
        out[s:s + m] = 0.3 * (2.0 * synthetic.splitmix64_uniform(seed, m, s) - 1.0) + mean_density
    return out


def build_pfc(pf, terms, dims, *, noise=None, filt_order=5, pinned: bool = False, noise_seed: int = 7, kspace_noise: bool = False):
    """Phase-field crystal: implicit pair-correlation term (two-peak set, main.go:63-72), mixed
    ideal-mixture term (main.go:75-83), optional white noise (``noise`` = "device": the device
    Philox stream through WhiteNoise.Generate; None: no noise) and Vandeven(filt_order) filter.
    ``kspace_noise``: Model.SetKSpaceNoise (needs a library built with -DGOPF_KNOISE).
    Returns (model, density, solver)."""
    import math
    n = int(np.prod(dims))
    m = pf.NewModel()
    if pinned:
        f = pf.NewField("density", n, None, pinned=True)
        f.Data[:] = pfc_initial(n)
    else:
        f = pf.NewField("density", n, pfc_initial(n))
    m.AddField(f)
    a = PFC_LATTICE
    peaks = [terms.Peak(1.0, 2.0 * math.pi / a, PFC_PEAK_WIDTH, 4),
             terms.Peak(1.0 / math.sqrt(2.0), 2.0 * math.pi / (a / math.sqrt(2.0)), PFC_PEAK_WIDTH, 4)]
    term = terms.PairCorrlationTerm(terms.ReciprocalSpacePairCorrelation(PFC_EFF_TEMP, peaks), "density", 1.0, True)
    ideal = terms.IdealMixtureTerm(terms.IdealMix(1.0, 1.0), "density", 1.0, True)
    m.RegisterImplicitTerm("EXCESS", term, None)
    m.RegisterMixedTerm("IDEAL", ideal, [ideal.DerivedField(n, m.Bricks)])
    eq = "ddensity/dt = IDEAL + EXCESS"
    if noise == "device":
        m.RegisterFunction("NOISE", terms.WhiteNoise(PFC_NOISE_STRENGTH, seed=noise_seed).Generate)
        eq += " + NOISE"
    elif noise is not None:
        m.RegisterFunction("NOISE", noise)
        eq += " + NOISE"
    m.AddEquation(eq)
    if kspace_noise:  # draw the noise spectrum at the k-point (this package only)
        m.SetKSpaceNoise(True)
    solver = pf.NewSolver(m, dims, PFC_DT)
    if filt_order is not None:
        solver.Stepper.SetFilter(terms.NewVandeven(filt_order))
    return m, f, solver


# ---- examples/electricConductivity/main.go:60-160: charge relaxation in a polycrystal -------------
CHARGE_DT = 0.1                     # main.go:93
CHARGE_EXTERNAL_FIELD = [1.0, 0.0]  # main.go:134
CHARGE_REF_CONDUCTIVITY = (1.0, 3.0)  # main.go:119: diag(1, 3) rotated per grain


def charge_conductivity(dims) -> np.ndarray:
    """(N, 3) Voigt conductivity [s_xx, s_yy, s_xy] of a 2-D polycrystal: vertical stripes of grains,
    grain g rotated by g*pi/9 (main.go:122-125), smooth blend at the boundaries (the example blurs
    the grain indicator fields, main.go:62-84)."""
    import math
    n0, n1 = dims
    idx = np.arange(n0 * n1)
    c = idx % n1
    n_grains = 4
    pos = c / float(n1) * n_grains
    g0 = np.floor(pos).astype(int) % n_grains
    g1 = (g0 + 1) % n_grains
    w = np.clip((pos - np.floor(pos) - 0.8) / 0.2, 0.0, 1.0)  # blend over the last fifth of a grain
    out = np.zeros((n0 * n1, 3))
    s1, s2 = CHARGE_REF_CONDUCTIVITY
    for g, weight in ((g0, 1.0 - w), (g1, w)):
        ang = g * math.pi / 9.0
        ca, sa = np.cos(ang), np.sin(ang)
        out[:, 0] += weight * (s1 * ca * ca + s2 * sa * sa)
        out[:, 1] += weight * (s1 * sa * sa + s2 * ca * ca)
        out[:, 2] += weight * ((s1 - s2) * ca * sa)
    return out


def build_charge(pf, terms, dims, ft=None):
    """density field with d rho/dt = MINUS_DIV_CURRENT (main.go:131-140).  ``ft`` is the
    FourierTransform the reference's struct carries (needed by the oracle, unused on the device).
    Returns (model, density, term, solver)."""
    from . import synthetic
    n = int(np.prod(dims))
    sigma = charge_conductivity(dims)
    m = pf.NewModel()
    init = (0.05 * (2.0 * synthetic.splitmix64_uniform(3, n) - 1.0)).astype(np.complex128)
    f = pf.NewField("density", n, init)
    m.AddField(f)
    term = terms.ChargeTransport(lambda i: sigma[i], CHARGE_EXTERNAL_FIELD, "density", ft)
    m.RegisterExplicitTerm("MINUS_DIV_CURRENT", term, None)
    m.AddEquation("ddensity/dt = MINUS_DIV_CURRENT")
    return m, f, term, pf.NewSolver(m, dims, CHARGE_DT)


# ---- diffusion with two time-dependent point sources (pf/sourceTerm.go, pf/model.go:151-154) ------
def build_sourced_diffusion(pf, terms, dims):
    import math
    from . import synthetic
    n = int(np.prod(dims))
    rank = len(dims)
    m = pf.NewModel()
    f = pf.NewField("conc", n, (0.05 * (2.0 * synthetic.splitmix64_uniform(5, n) - 1.0)).astype(np.complex128))
    m.AddField(f)
    m.AddScalar(pf.NewScalar("D", 0.8))
    m.AddEquation("dconc/dt = D*LAP conc - conc^3")
    m.AddSource(0, terms.NewSource([3.0, 5.0, 2.0][:rank], lambda t: 2.0 * t + 0.5))
    m.AddSource(0, terms.NewSource([10.5, 1.25, 7.0][:rank], lambda t: math.cos(3.0 * t)))
    return m, f, pf.NewSolver(m, dims, 0.05)


# ---- pf/sdd_test.go:163-246 (TestClassicalNucleation): critical nucleus by shrinking dimer ---------
SDD_GAMMA, SDD_RHO, SDD_DT = 0.5, 0.05, 0.7


def _disc(n: int, radius: int) -> np.ndarray:
    """-1 outside / +1 inside a centred disc, blurred by the 5 x 5 periodic box (insertCircleAtCenter +
    pfutil.Blur with BoxKernel{Width: 2}, sdd_test.go:151-161, 199-205)."""
    i = np.arange(n * n)
    dx, dy = i // n - n // 2, i % n - n // 2
    v = np.where(dx * dx + dy * dy <= radius * radius, 1.0, -1.0).reshape(n, n)
    acc = np.zeros_like(v)
    for dr in range(-2, 3):
        for dc in range(-2, 3):
            acc += np.roll(np.roll(v, dr, axis=0), dc, axis=1)
    return (acc / 25.0).reshape(-1)


def build_sdd_nucleation(pf, new_sdd, n: int, *, expressions: bool):
    """Returns (model, phi, sdd, solver) with the stepper assigned, as the reference test does."""
    rho = SDD_RHO
    phi = pf.NewField("phi", n * n, _disc(n, 12).astype(np.complex128))
    m = pf.NewModel()
    m.AddField(phi)
    m.AddScalar(pf.NewScalar("gamma", SDD_GAMMA))
    if expressions:
        m.RegisterFunction("MINUS_CHEM_POT", f"(1.0 - phi*phi)*(phi + {3.0 * rho / 4.0!r})")
    else:
        m.RegisterFunction("MINUS_CHEM_POT", lambda i, b: (1.0 - b["phi"].Get(i) ** 2) * (b["phi"].Get(i) + 3.0 * rho / 4.0))
    m.AddEquation("dphi/dt = MINUS_CHEM_POT + gamma*LAP phi")
    sdd = new_sdd([n, n], m)
    sdd.InitDimerLength = 1.0
    sdd.MinDimerLength = 5e-6
    sdd.Dt = SDD_DT
    sdd.Init([pf.NewField("a", n * n, _disc(n, 10).astype(np.complex128))],
             [pf.NewField("b", n * n, _disc(n, 15).astype(np.complex128))])
    solver = pf.NewSolver(m, [n, n], SDD_DT)
    solver.Stepper = sdd
    return m, phi, sdd, solver
