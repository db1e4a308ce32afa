"""The device evaluators of the general path, compiled for the HOST, against the oracle -- no GPU.

`gopf_b200/csrc/kupdate.cuh`, `step_program.h` and `cplx.cuh` are header-only device code: the k-space
update of pf/euler.go:27-39 with every catalog term, the derived-field evaluators of
pf/model.go:237-241 (monomials with Go's cmplx.Pow, registered functions, Philox noise, tables), the
literal `Freq`, the modal filter.  tests/host_emul/ compiles exactly these headers with g++ behind a
small CUDA shim; `EmulatedSolver` below wires them to the programs the C++ model compiles
(`gopf_model_program_image`, `gopf_model_derived_image`), with scipy's FFT standing where the CUDA
transforms are.  The step-level device tests of tests/test_step_gpu.py that stay on the
semi-implicit Euler general path are then run unchanged against it, so the oracle checks the
product's own evaluator source on CPU.  This is test infrastructure: nothing in gopf_b200/ can reach
it, and the transforms, the fused kernels and the launch code remain covered by the device tests only.
"""
import ctypes
import os
import shutil
import subprocess
import types

import numpy as np
import pytest
import scipy.fft

from gopf_b200 import pf as gpf
from gopf_b200._lib import check, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAX_SPECTRA, MAX_FIELDS = 16, 4
DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("emul") / "emul.so"
    # -DGOPF_KNOISE: the k-space noise generator, which the library's device build does not carry yet
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-w", "-DGOPF_KNOISE", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I",
                    os.path.join(ROOT, "gopf_b200", "csrc"), "-o", str(so), os.path.join(ROOT, "tests", "host_emul", "emul.cpp")],
                   check=True)
    dll = ctypes.CDLL(str(so))
    dll.emul_sizeof_program.restype = ctypes.c_int
    dll.emul_sizeof_derived.restype = ctypes.c_int
    return dll


def _image(fn, *args, tail=()):
    need = ctypes.c_int64(0)
    check(fn(*args, None, ctypes.c_int64(0), ctypes.byref(need), *tail))
    buf = ctypes.create_string_buffer(need.value)
    return buf, need


class _Term(ctypes.Structure):  # DevTerm (step_program.h)
    _fields_ = [("cre", ctypes.c_double), ("cim", ctypes.c_double), ("brick", ctypes.c_int), ("lap", ctypes.c_int),
                ("kind", ctypes.c_int), ("param", ctypes.c_int)]


class _Equation(ctypes.Structure):  # DevEquation: GOPF_MAX_TERMS = 8
    _fields_ = [("n_rhs", ctypes.c_int), ("n_den", ctypes.c_int), ("rhs", _Term * 8), ("den", _Term * 8)]


class _ProgramHead(ctypes.Structure):  # the leading members of DevKProgram: GOPF_MAX_FIELDS = 4
    _fields_ = [("rank", ctypes.c_int), ("n_fields", ctypes.c_int), ("dt", ctypes.c_double), ("eq", _Equation * MAX_FIELDS)]


class ProgramView:
    def __init__(self, buf):
        self.head = _ProgramHead.from_buffer(buf)
        self.n_fields = self.head.n_fields

    def den_terms(self, i):
        q = self.head.eq[i]
        return [q.den[j] for j in range(q.n_den)]

    def rhs_terms(self, i):
        q = self.head.eq[i]
        return [q.rhs[j] for j in range(q.n_rhs)]


class EmulatedSolver:
    """gopf_b200.pf.Solver's surface, semi-implicit Euler on the general path, on the host."""

    IsFused = False

    def __init__(self, dll, m, dims, dt):
        self.dll, self.Model, self.dims, self.Dt = dll, m, [int(d) for d in dims], dt
        self.Callbacks, self.Monitors, self.StartEpoch = [], [], 0
        self.filter = None
        self.current_step = 0  # Euler.CurrentStep; RK4.Step never advances it (rk4.go:130-135)
        self.steps_taken = 0   # counter of the noise stream
        self.launches = 0
        self.Stepper = types.SimpleNamespace(SetFilter=self._set_filter, GetTime=lambda: self.current_step * self.Dt, Dt=dt)
        self.rank = len(dims)
        self.N = int(np.prod(dims))
        for f in m.Fields:
            if f.Data.shape[0] != self.N:  # solver.go:55-60
                raise gpf.GopfError("solver: Inconsistent domain size and number of grid points")
        h = m._h
        buf, need = _image(lib().gopf_model_program_image, h, self.rank, ctypes.c_double(dt))
        assert need.value == dll.emul_sizeof_program(), "emulation and library disagree on sizeof(DevKProgram)"
        check(lib().gopf_model_program_image(h, self.rank, ctypes.c_double(dt), buf, need, None))
        self.program = buf
        nd = ctypes.c_int(0)
        check(lib().gopf_model_num_derived_fields(h, ctypes.byref(nd)))
        self.derived = []
        for d in range(nd.value):
            used = ctypes.c_int(0)
            dbuf, dneed = _image(lib().gopf_model_derived_image, h, d, tail=(None,))
            assert dneed.value == dll.emul_sizeof_derived()
            check(lib().gopf_model_derived_image(h, d, dbuf, dneed, None, ctypes.byref(used)))
            self.derived.append((dbuf, bool(used.value)))
        if len(m.Fields) + nd.value > MAX_SPECTRA:
            raise gpf.GopfError("too many spectra")
        self.tables = getattr(m, "_emul_tables", [])
        self.stepper = "euler"
        self.lp_state = np.array([0.0, 0.0, 1.0] * 2)  # per slot: Multiplier, CurrentIntegral, IsFirstUpdate (solver.cu)
        # VolumeConservingLP terms: slots in name order (std::map in model.cu), field / indicator spectrum
        names = m.AllFieldNames()
        self.lp_terms = [(names.index(t.Field), names.index(t.Indicator), t.Dt)
                         for _, t in sorted(m.ExplicitTerms.items()) if isinstance(t, gpf.VolumeConservingLP)]
        self.tabs = None

    # -- the surface the tests use
    def SetStepper(self, name):
        if name not in ("euler", "rk4"):
            raise gpf.GopfError("Unknown stepper scheme")
        self.stepper = name

    def LPMultiplier(self, slot=0):
        return float(self.lp_state[3 * slot])

    def _set_filter(self, filt):
        self.filter = None if filt is None else np.ascontiguousarray(filt.Data, dtype=np.float64)
        self.tabs = None

    def ForceGeneric(self, on=True):
        pass

    def KernelLaunches(self, reset=False):
        return self.launches

    def Solve(self, nepochs, nsteps):
        for i in range(nepochs):
            self.Propagate(nsteps)
            for cb in self.Callbacks:
                cb(self, i + self.StartEpoch)
            for mon in self.Monitors:
                mon.Add(self.Model.Bricks)

    def Propagate(self, nsteps):
        self.Upload()
        self.StepDevice(nsteps)
        self.Download()

    def Upload(self):  # euler.go:19-21
        self.S = [self._fft(f.Data) for f in self.Model.Fields]

    def StepDevice(self, nsteps):
        for _ in range(nsteps):
            if self.stepper == "rk4":
                self._rk4_step(self.S)
            else:
                self._step(self.S)

    def Download(self):  # euler.go:42-45
        for f, s in zip(self.Model.Fields, self.S):
            f.Data[:] = self._ifft(s)

    # -- one Solver::euler_update_generic
    def _fft(self, x):
        return np.ascontiguousarray(scipy.fft.fftn(x.reshape(self.dims)).reshape(-1))

    def _ifft(self, x):
        return np.ascontiguousarray(scipy.fft.ifftn(x.reshape(self.dims)).reshape(-1))  # includes the 1/N

    def _ptrs(self, arrays, count):
        return (DP * count)(*[a.ctypes.data_as(DP) if a is not None else None for a in arrays])

    def _geom(self):
        d = self.dims
        return self.rank, d[0], d[1], d[2] if self.rank > 2 else 1, ctypes.c_longlong(self.N)

    def _filter_args(self):
        if self.filter is None:
            return None, 0
        return self.filter.ctypes.data_as(DP), self.filter.shape[0]

    def _lp_args(self):
        base = self.lp_state.ctypes.data
        return ctypes.cast(base, DP), ctypes.cast(base + 24, DP)

    def _implicit_tabs(self):
        """Solver::launch_update: tabulate filter / (1 - dt den) when it is expensive per k and depends on k only."""
        if self.tabs is None:
            P = ProgramView(self.program)
            self.tabs = [None] * MAX_FIELDS
            for i in range(P.n_fields):
                den = P.den_terms(i)
                expensive = self.filter is not None or any(t.kind in (1, 2) for t in den)  # TK_SPECTRAL_VISC, TK_PAIR_CORR
                k_only = all(t.brick < 0 and t.kind not in (3, 4) for t in den)            # TK_CONS_NOISE, TK_VOLUME_LP
                if expensive and k_only:
                    out = np.empty(self.N, dtype=np.complex128)
                    self.dll.emul_implicit_table(self.program, *self._filter_args(), i, out.ctypes.data_as(DP), *self._geom())
                    self.tabs[i] = out
        return self.tabs

    def _derived_spectra(self, S):
        F = len(S)
        spectra = list(S) + [None] * (MAX_SPECTRA - F)
        if any(used for _, used in self.derived):
            R = [self._ifft(s) for s in S]
            rp = (DP * MAX_FIELDS)(*[R[i].ctypes.data_as(DP) if i < F else None for i in range(MAX_FIELDS)])
            table_no = 0
            for d, (dbuf, used) in enumerate(self.derived):
                kind = ctypes.c_int.from_buffer(dbuf).value
                table, tn = None, 0
                if kind == 3:  # DK_TABLE
                    tab = self.tables[table_no]
                    table_no += 1
                    table, tn = tab.ctypes.data_as(DP), tab.shape[1]
                if not used:
                    continue
                out = np.empty(self.N, dtype=np.complex128)
                self.dll.emul_derived(dbuf, rp, table, ctypes.c_longlong(tn), out.ctypes.data_as(DP),
                                      ctypes.c_ulonglong(self.steps_taken), ctypes.c_longlong(self.N))
                spectra[F + d] = self._fft(out)
                self.launches += 1
        return spectra

    def _step(self, S):  # Solver::euler_step_generic
        self.dll.emul_stamp_noise_step(self.program, ctypes.c_ulonglong(self.steps_taken))
        spectra = self._derived_spectra(S)
        tabs = self._implicit_tabs() if self.use_tabs else [None] * MAX_FIELDS
        self.dll.emul_update(self.program, self._ptrs(spectra, MAX_SPECTRA), self._ptrs(tabs, MAX_FIELDS), *self._filter_args(),
                             *self._lp_args(), *self._geom())
        self.launches += 1
        for slot, (fi, ii, dt) in enumerate(self.lp_terms):  # Solver::volume_lp_hooks
            state = ctypes.cast(self.lp_state.ctypes.data + 24 * slot, DP)
            self.dll.emul_volume_lp_update(state, spectra[fi].ctypes.data_as(DP), spectra[ii].ctypes.data_as(DP), ctypes.c_double(dt))
        self.current_step += 1
        self.steps_taken += 1

    use_tabs = True

    def _rk4_step(self, S):  # Solver::rk4_step (pf/rk4.go:29-74); S is updated in place
        F = len(S)
        init = [s.copy() for s in S]
        fin = [s.copy() for s in S]
        kf = [np.zeros(self.N, dtype=np.complex128) for _ in S]
        pad = [None] * (MAX_SPECTRA - F)

        def sync_and_rhs():
            spectra = self._derived_spectra(S)
            self.dll.emul_rk4_rhs(self.program, self._ptrs(spectra, MAX_SPECTRA), self._ptrs(kf + pad, MAX_SPECTRA),
                                  *self._lp_args(), *self._geom())
            self.launches += 1
            return spectra

        def point(mode, fdt, spectra):
            self.dll.emul_rk4_point(self.program, *self._filter_args(), *self._lp_args(), mode, ctypes.c_double(fdt),
                                    self._ptrs(spectra, MAX_SPECTRA), self._ptrs(init + pad, MAX_SPECTRA),
                                    self._ptrs(fin + pad, MAX_SPECTRA), self._ptrs(kf + pad, MAX_SPECTRA), *self._geom())
            self.launches += 1

        self.dll.emul_stamp_noise_step(self.program, ctypes.c_ulonglong(self.steps_taken))
        dt = self.Dt
        for first, second in (((0, dt / 6.0), (1, 0.5 * dt)), ((0, dt / 3.0), (1, 0.5 * dt)), ((0, dt / 3.0), (1, 1.0 * dt)),
                              ((0, dt / 6.0), (2, dt))):
            spectra = sync_and_rhs()
            point(*first, spectra)
            point(*second, spectra)
        for slot, (fi, ii, dt) in enumerate(self.lp_terms):  # Solver::volume_lp_hooks (solver.go:74-82)
            state = ctypes.cast(self.lp_state.ctypes.data + 24 * slot, DP)
            self.dll.emul_volume_lp_update(state, spectra[fi].ctypes.data_as(DP), spectra[ii].ctypes.data_as(DP), ctypes.c_double(dt))
        self.steps_taken += 1


class _RecordingModel(gpf.Model):
    """gopf_b200.pf.Model that also keeps the host copy of prescribed table fields for the emulation."""

    def RegisterTableField(self, name, values):
        super().RegisterTableField(name, values)
        self.__dict__.setdefault("_emul_tables", []).append(np.ascontiguousarray(values, dtype=np.float64))


@pytest.fixture()
def T(emul, monkeypatch):
    """tests/test_step_gpu.py with its `gpf` bound to the host emulation."""
    import test_step_gpu as mod

    shim = types.SimpleNamespace(**{k: getattr(gpf, k) for k in dir(gpf) if not k.startswith("__")})
    shim.NewModel = _RecordingModel
    shim.NewSolver = lambda m, dims, dt, device=-1: EmulatedSolver(emul, m, dims, dt)
    monkeypatch.setattr(mod, "gpf", shim)
    return mod


# ---- the device tests that stay on the Euler general path, run against the emulation ------------------
@pytest.mark.parametrize("dims", [[128, 128], [64, 256], [32, 32, 32]], ids=lambda d: "x".join(map(str, d)))
def test_cahn_hilliard_100_steps(T, dims):
    T.test_cahn_hilliard_100_steps(dims, True)


def test_euler_decays(T):
    T.test_euler_exponential_decay()
    T.test_euler_square_decay()


def test_solver_diffusion(T):
    T.test_solver_diffusion()


def test_gauss_seidel_field_ordering(T):
    T.test_gauss_seidel_field_ordering()


@pytest.mark.parametrize("dims", [[32, 32], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_reaction_diffusion_three_fields(T, dims):
    T.test_reaction_diffusion_three_fields(dims)


@pytest.mark.parametrize("lap", [False, True], ids=["nolap", "lap"])
def test_pfc_pair_correlation_ideal_mixture(T, lap):
    T.test_pfc_pair_correlation_ideal_mixture_vs_oracle(lap)


def test_pfc_with_vandeven_filter_and_prescribed_noise(T):
    T.test_pfc_with_vandeven_filter_and_prescribed_noise_vs_oracle()


def test_spectral_viscosity(T):
    T.test_spectral_viscosity_vs_oracle()


def test_white_noise_statistics(T):
    T.test_white_noise_statistics()


def test_two_white_noise_fields_are_independent(T):
    T.test_two_white_noise_fields_are_independent()


def test_conservative_noise(T):
    T.test_conservative_noise_properties()
    T.test_conservative_noise_prescribed_currents_vs_oracle()


@pytest.mark.parametrize("dims,K", [([64, 64], [1.0, 0.3, 0.3, 2.0]), ([16, 16, 16], [1.0, 0.2, 0.1, 0.2, 2.0, 0.3, 0.1, 0.3, 0.5])],
                         ids=["2d", "3d"])
def test_tensorial_hessian(T, dims, K):
    T.test_tensorial_hessian_anisotropic_diffusion_vs_oracle(dims, K)


def test_negative_value_penalty(T):
    T.test_negative_value_penalty_vs_oracle()


def test_rk4(T):
    T.test_rk4_simple_model_and_implicit()
    T.test_rk4_cahn_hilliard_vs_oracle([32, 32])
    T.test_rk4_cahn_hilliard_vs_oracle([16, 16, 16])


@pytest.mark.parametrize("stepper", ["euler", "rk4"])
def test_two_phase_functions_and_volume_constraint(T, stepper):
    T.test_two_phase_functions_and_volume_constraint_vs_oracle(stepper)


def test_tabulated_and_literal_implicit_side_agree(emul, monkeypatch):
    """pfc with pair correlation + Vandeven filter: launch_update tabulates filter / (1 - dt den); the
    literal evaluation of the same factor must give the same trajectory."""
    import test_step_gpu as mod

    out = []
    for use_tabs in (True, False):
        shim = types.SimpleNamespace(**{k: getattr(gpf, k) for k in dir(gpf) if not k.startswith("__")})
        shim.NewModel = _RecordingModel
        cls = type("S", (EmulatedSolver,), {"use_tabs": use_tabs})
        shim.NewSolver = lambda m, dims, dt, device=-1, cls=cls: cls(emul, m, dims, dt)
        monkeypatch.setattr(mod, "gpf", shim)
        (gm, gf, gs), _ = mod.pfc_models([32, 32], True, filt_order=5)
        gs.Solve(2, 10)
        if use_tabs:
            assert gs.tabs is not None and gs.tabs[0] is not None
        out.append(gf.Data.copy())
    assert np.linalg.norm(out[0] - out[1]) <= 1e-13 * np.linalg.norm(out[1])


# ---- white noise drawn in k-space (TK_WHITE_NOISE_K, -DGOPF_KNOISE) -----------------------------------------
def _noise_model(dims, strength, equation, seed=5, kspace=True):
    n = int(np.prod(dims))
    m = _RecordingModel()
    f = gpf.NewField("conc", n, np.zeros(n, dtype=np.complex128))
    m.AddField(f)
    m.AddScalar(gpf.NewScalar("m1", -1.0))
    m.RegisterFunction("NOISE", gpf.WhiteNoise(strength, seed=seed).Generate)
    m.AddEquation(equation)
    m.SetKSpaceNoise(kspace)
    return m, f


# 2-D and cubic 3-D only: elsewhere the reference's Freq does not follow the transform's layout and
# the solver refuses k-space noise
@pytest.mark.parametrize("dims", [[64, 64], [16, 16, 16], [12, 20], [27, 9], [12, 12, 12]], ids=lambda d: "x".join(map(str, d)))
def test_kspace_noise_is_real_white_noise_of_the_reference_variance(emul, dims):
    """dconc/dt = NOISE from conc = 0 with dt = 1: after one step conc IS the noise field.  WhiteNoise.Generate
    (pf/noise.go:20-23) draws N(0, 2 Strength) per node; drawn in k-space the field must come out real
    (Hermitian spectrum, also on grids whose k-table is inexact), with that variance, uncorrelated, Gaussian,
    and fresh at every step."""
    strength = 0.125
    n = int(np.prod(dims))
    m, f = _noise_model(dims, strength, "dconc/dt = NOISE")
    s = EmulatedSolver(emul, m, dims, 1.0)
    assert [used for _, used in s.derived] == [False]  # the noise field is never evaluated nor transformed
    s.Propagate(1)
    x1 = f.Data.copy()
    assert np.max(np.abs(x1.imag)) <= 1e-14 * np.max(np.abs(x1.real))
    f.Data[:] = 0.0
    s.Propagate(1)  # the second step of the same stream
    x2 = f.Data.copy()
    var = 2.0 * strength
    for x in (x1.real, x2.real):
        assert abs(np.mean(x)) < 5.0 * np.sqrt(var / n)
        assert abs(np.var(x) / var - 1.0) < 5.0 * np.sqrt(2.0 / n)
        assert abs(np.mean(x ** 4) / var ** 2 - 3.0) < 5.0 * np.sqrt(96.0 / n)
        grid = x.reshape(dims)
        for axis in range(len(dims)):  # nearest-neighbour correlation along every axis
            assert abs(np.mean(grid * np.roll(grid, 1, axis=axis)) / var) < 5.0 / np.sqrt(n)
    assert abs(np.mean(x1.real * x2.real) / var) < 5.0 / np.sqrt(n)
    # same seed, same step: the same field (the stream is keyed by frequency, seed and step)
    m2, f2 = _noise_model(dims, strength, "dconc/dt = NOISE")
    EmulatedSolver(emul, m2, dims, 1.0).Propagate(1)
    assert np.array_equal(f2.Data, x1)
    # another seed: another field
    m3, f3 = _noise_model(dims, strength, "dconc/dt = NOISE", seed=6)
    EmulatedSolver(emul, m3, dims, 1.0).Propagate(1)
    assert abs(np.mean(f3.Data.real * x1.real) / var) < 5.0 / np.sqrt(n)


def test_kspace_noise_spectrum_is_flat_and_takes_prefactors(emul):
    dims, strength = [32, 32], 0.5
    n = 1024
    m, f = _noise_model(dims, strength, "dconc/dt = NOISE")
    EmulatedSolver(emul, m, dims, 1.0).Propagate(1)
    spec = scipy.fft.fftn(f.Data.reshape(dims))
    power = np.abs(spec) ** 2 / (n * 2.0 * strength)  # E = 1 at every k
    assert abs(np.mean(power) - 1.0) < 0.15
    ky, kx = np.meshgrid(np.fft.fftfreq(32), np.fft.fftfreq(32), indexing="ij")
    low = np.hypot(kx, ky) < 0.2
    assert abs(np.mean(power[low]) - np.mean(power[~low])) < 0.3
    # m1 * LAP NOISE: coefficient and Laplacian apply to the drawn spectrum (rhsBuilder.go:158-188)
    m2, f2 = _noise_model(dims, strength, "dconc/dt = m1*LAP NOISE")
    EmulatedSolver(emul, m2, dims, 1.0).Propagate(1)
    spec2 = scipy.fft.fftn(f2.Data.reshape(dims))
    L = -(2.0 * np.pi) ** 2 * (kx ** 2 + ky ** 2)
    assert np.allclose(spec2, -1.0 * L * spec, rtol=1e-12, atol=1e-12 * np.max(np.abs(spec)))


def test_kspace_noise_off_keeps_the_real_space_field(emul):
    m, f = _noise_model([16, 16], 0.1, "dconc/dt = NOISE", kspace=False)
    s = EmulatedSolver(emul, m, [16, 16], 1.0)
    assert [used for _, used in s.derived] == [True]
    assert not any(t.kind == 6 for t in ProgramView(s.program).rhs_terms(0))
    m2, f2 = _noise_model([16, 16], 0.1, "dconc/dt = NOISE", kspace=True)
    s2 = EmulatedSolver(emul, m2, [16, 16], 1.0)
    assert [t.kind for t in ProgramView(s2.program).rhs_terms(0)] == [6]


# ---- seeded random models: parser -> program -> evaluators against the oracle, numerically ---------------
@pytest.mark.parametrize("seed", range(80))
def test_random_models_step_like_the_oracle(emul, seed):
    """The equations of tests/test_parser_fuzz_cpu.py (the reference's grammar: signs, scalar powers,
    LAP^n prefixes, products of field powers, names of fields absent from the model), stepped."""
    import random

    from oracle import pf as opf
    from oracle import pfutil as opfutil
    from test_parser_fuzz_cpu import FIELDS, SCALARS, random_equation

    rng = random.Random(1000 + seed)
    dims = rng.choice([[8, 8], [4, 16], [4, 4, 4]])
    n = int(np.prod(dims))
    names = FIELDS[:rng.choice([1, 2, 3])]
    eqs = [random_equation(rng, f) for f in names]
    dt = 1e-3
    res = []
    for mod in (gpf, opf):
        m = _RecordingModel() if mod is gpf else mod.NewModel()
        fields = []
        for k, name in enumerate(names):
            f = mod.NewField(name, n, (0.6 + 0.3 * opfutil.splitmix64_uniform(seed * 7 + k, n)).astype(np.complex128))
            m.AddField(f)
            fields.append(f)
        for sname, val in SCALARS:
            m.AddScalar(mod.NewScalar(sname, val))
        try:
            for eq in eqs:
                m.AddEquation(eq)
            solver = EmulatedSolver(emul, m, dims, dt) if mod is gpf else mod.NewSolver(m, dims, dt)
            solver.Solve(1, 3)
            res.append([f.Data.copy() for f in fields])
        except Exception as e:  # noqa: BLE001 -- the reference panics; either side raises its own type
            res.append(e)
    failed = [isinstance(r, Exception) for r in res]
    assert failed[0] == failed[1], (eqs, res)
    if failed[0]:
        return
    for a, b in zip(*res):
        if not np.all(np.isfinite(b)) or np.max(np.abs(b)) > 1e8:
            assert not np.all(np.isfinite(a)) or np.max(np.abs(a)) > 1e6, eqs  # blown up on both sides
            continue
        assert np.linalg.norm(a - b) <= 1e-9 * max(np.linalg.norm(b), 1e-300), (eqs, dims)


# ---- the specialised k-space update (jit.cu), the very unit NVRTC receives, compiled for the host -----
class _Spectra(ctypes.Structure):
    _fields_ = [("s", DP * MAX_SPECTRA)]


class _Tabs(ctypes.Structure):
    _fields_ = [("t", DP * MAX_FIELDS)]


class SpecialisedEmulatedSolver(EmulatedSolver):
    """EmulatedSolver whose update is `gopf_jit_kupdate`: program image as constant words, grid geometry,
    node count and the (host) address of the filter table as literals."""

    use_tabs = False

    build_dir = None
    calls = 0

    def _step(self, S):
        self._with_unit(super()._step, S)

    def _rk4_step(self, S):
        self._with_unit(super()._rk4_step, S)

    def _with_unit(self, step, S):
        if not hasattr(self, "_kernel") or self._filter_id != id(self.filter):
            self._filter_id = id(self.filter)
            addr = self.filter.ctypes.data if self.filter is not None else 0
            fn = self.filter.shape[0] if self.filter is not None else 0
            src = self.Model.KUpdateSource(self.dims, self.Dt, 0, filter_addr=addr, filter_n=fn, lp_addr=self.lp_state.ctypes.data)
            assert "jit_prog_words" in src
            SpecialisedEmulatedSolver.serial = getattr(SpecialisedEmulatedSolver, "serial", 0) + 1
            stem = os.path.join(self.build_dir, f"unit_{SpecialisedEmulatedSolver.serial}")
            with open(stem + ".cpp", "w") as f:
                f.write('#include "cuda_shim.h"\n' + src)
            subprocess.run(["g++", "-O1", "-ffp-contract=off", "-w", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I", os.path.join(ROOT, "gopf_b200", "csrc"),
                            "-I", os.path.join(ROOT, "tests", "host_emul"), "-o", stem + ".so", stem + ".cpp"], check=True)
            unit = ctypes.CDLL(stem + ".so")
            self._kernel = unit.gopf_jit_kupdate
            self._kernel.argtypes = [_Spectra, _Tabs]
            self._kernel.restype = None
            self._rk4_rhs = unit.gopf_jit_rk4_rhs
            self._rk4_rhs.argtypes = [_Spectra, _Spectra]
            self._rk4_rhs.restype = None
            self._rk4_point = unit.gopf_jit_rk4_point
            self._rk4_point.argtypes = [ctypes.c_int, ctypes.c_double, _Spectra, _Spectra, _Spectra, _Spectra]
            self._rk4_point.restype = None
        def specialised(program, sp, tabs, filt, fn, lp0, lp1, rank, d0, d1, d2, n):
            self._kernel(_Spectra(sp), _Tabs())  # tab_mask 0: the literal implicit side
            SpecialisedEmulatedSolver.calls += 1

        def rk4_rhs(program, sp, kout, lp0, lp1, rank, d0, d1, d2, n):
            self._rk4_rhs(_Spectra(sp), _Spectra(kout))
            SpecialisedEmulatedSolver.calls += 1

        def rk4_point(program, filt, fn, lp0, lp1, mode, fdt, field, initial, final_, kf, rank, d0, d1, d2, n):
            self._rk4_point(mode, fdt, _Spectra(field), _Spectra(initial), _Spectra(final_), _Spectra(kf))
            SpecialisedEmulatedSolver.calls += 1

        real = self.dll
        self.dll = types.SimpleNamespace(emul_update=specialised, emul_rk4_rhs=rk4_rhs, emul_rk4_point=rk4_point,
                                         emul_derived=real.emul_derived, emul_volume_lp_update=real.emul_volume_lp_update,
                                         emul_stamp_noise_step=real.emul_stamp_noise_step)
        try:
            step(S)
        finally:
            self.dll = real


@pytest.fixture()
def TS(emul, monkeypatch, tmp_path):
    import test_step_gpu as mod

    SpecialisedEmulatedSolver.build_dir = str(tmp_path)
    shim = types.SimpleNamespace(**{k: getattr(gpf, k) for k in dir(gpf) if not k.startswith("__")})
    shim.NewModel = _RecordingModel
    shim.NewSolver = lambda m, dims, dt, device=-1: SpecialisedEmulatedSolver(emul, m, dims, dt)
    monkeypatch.setattr(mod, "gpf", shim)
    return mod


def test_specialised_update_unit_on_the_host(TS):
    SpecialisedEmulatedSolver.calls = 0
    TS.test_cahn_hilliard_100_steps([64, 256], True)  # non-square: d0, d1 literals in the right order
    TS.test_reaction_diffusion_three_fields([16, 16, 16])
    TS.test_pfc_with_vandeven_filter_and_prescribed_noise_vs_oracle()  # filter address as a literal
    TS.test_spectral_viscosity_vs_oracle()
    TS.test_conservative_noise_prescribed_currents_vs_oracle()
    TS.test_tensorial_hessian_anisotropic_diffusion_vs_oracle([16, 16, 16], [1.0, 0.2, 0.1, 0.2, 2.0, 0.3, 0.1, 0.3, 0.5])
    assert SpecialisedEmulatedSolver.calls > 200  # every step above went through gopf_jit_kupdate


def test_specialised_rk4_and_volume_constraint_units_on_the_host(TS):
    SpecialisedEmulatedSolver.calls = 0
    TS.test_rk4_simple_model_and_implicit()
    TS.test_rk4_cahn_hilliard_vs_oracle([16, 16, 16])
    assert SpecialisedEmulatedSolver.calls > 100
    TS.test_two_phase_functions_and_volume_constraint_vs_oracle("euler")  # the multiplier's address as a literal
    TS.test_two_phase_functions_and_volume_constraint_vs_oracle("rk4")


def test_emulated_freq_is_the_reference_k_table(emul):
    from oracle import pfutil as opfutil
    for dims in ([8, 16], [6, 10], [8, 8, 8]):
        n = int(np.prod(dims))
        out = np.zeros((n, 3))
        emul.emul_freq(len(dims), dims[0], dims[1], dims[2] if len(dims) > 2 else 1, ctypes.c_longlong(n), out.ctypes.data_as(DP))
        ft = opfutil.NewFFTW(dims)
        want = np.array([ft.Freq(i) for i in range(n)])
        assert np.array_equal(out[:, :len(dims)], want)
