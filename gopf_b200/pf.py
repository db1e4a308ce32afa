"""Mirror of the reference's ``pf`` hot-path surface over the C ABI
(include/gopf_cuda.h).  Same names and argument meaning as the Go package so the
parity tests read like the reference's own tests:

    model = pf.NewModel()
    conc = pf.NewField("conc", N, None)
    model.AddScalar(pf.NewScalar("gamma", 2.0)); model.AddField(conc)
    model.AddEquation("dconc/dt = LAP conc^3 + m1*LAP conc + m1*gamma*LAP^2 conc")
    solver = pf.NewSolver(model, [nx, ny], dt)
    solver.Solve(10, 10)          # conc.Data holds the result, as in Go

Differences forced by "no Go closures on the device" (hard parts):
``RegisterFunction`` takes an expression string, and user terms come from the
catalog below (the reference's own PureTerm/MixedTerm implementations).  Failures
raise ``GopfError`` where the reference panics.
"""
from __future__ import annotations

import ctypes
import math
from typing import Callable, List, Optional

import numpy as np

from ._lib import GopfError, check, int_array, lib


def _s(x: str) -> bytes:
    return x.encode("utf-8")


def pinned_empty(n: int) -> np.ndarray:
    """complex128 array of n cells in page-locked host memory (gopf_host_alloc)."""
    p = ctypes.c_void_p()
    check(lib().gopf_host_alloc(ctypes.c_int64(16 * n), ctypes.byref(p)))
    buf = (ctypes.c_double * (2 * n)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.complex128, count=n)
    return arr, _PinnedOwner(p)


class _PinnedOwner:
    def __init__(self, p):
        self.p = p

    def __del__(self):
        try:
            lib().gopf_host_free(self.p)
        except Exception:
            pass


class Field:
    """pf.Field (pf/model.go:16-57): Data is the host []complex128."""

    def __init__(self, name: str, N: int, data: Optional[np.ndarray] = None, pinned: bool = False):
        if data is None:
            if pinned:
                self.Data, self._pin = pinned_empty(N)  # _pin keeps the allocation alive
                self.Data[:] = 0
            else:
                self.Data = np.zeros(N, dtype=np.complex128)
        else:
            if data.shape[0] != N:
                raise GopfError("model: Inconsistent length of data")
            if data.dtype != np.complex128 or not data.flags.c_contiguous:
                raise TypeError("Field data must be a C-contiguous complex128 array")
            self.Data = data
        self.Name = name

    # The C model keeps the host pointer of the array registered by Model.AddField for its lifetime (pf.Field.Data
    # is a Go slice header there).  Rebinding Data afterwards would leave the library with a dangling pointer, so
    # it is refused: write into the array (f.Data[:] = ...) instead.
    def __setattr__(self, key, value):
        if key == "Data" and getattr(self, "_registered", False):
            raise GopfError("Field.Data is registered with a model: assign into it (f.Data[:] = ...) instead of rebinding it")
        object.__setattr__(self, key, value)

    def Get(self, i):
        return self.Data[i]


def NewField(name: str, N: int, data: Optional[np.ndarray] = None, pinned: bool = False) -> Field:
    return Field(name, N, data, pinned)


class Scalar:
    """pf.Scalar (pf/model.go:86-111)."""

    def __init__(self, Name: str, Value: complex):
        self.Name = Name
        self.Value = complex(Value)


def NewScalar(name: str, value: complex) -> Scalar:
    return Scalar(name, value)


# ---- term catalog (device-expressible PureTerm / MixedTerm implementations) ----------
class SpectralViscosity:
    """pf.SpectralViscosity (pf/spectralViscosity.go:23-56)."""

    def __init__(self, Eps: float, DissipationThreshold: float, Power: int):
        self.Eps, self.DissipationThreshold, self.Power = Eps, DissipationThreshold, Power

    def _register(self, m: "Model", name: str, cls: str):
        if cls != "implicit":
            raise GopfError("SpectralViscosity is an implicit term")
        check(lib().gopf_model_register_spectral_viscosity(m._h, _s(name), ctypes.c_double(self.Eps),
                                                            ctypes.c_double(self.DissipationThreshold), int(self.Power)))


def HasKSpaceNoise() -> bool:
    """True when libgopfcuda's device code was built with -DGOPF_KNOISE (Model.SetKSpaceNoise usable on the GPU)."""
    return bool(lib().gopf_has_kspace_noise())


class NegativeValuePenalty:
    """pf.NegativeValuePenalty (pf/negative_value_penalty.go:5-38).  ``Evaluate`` is what the
    reference passes to RegisterFunction; here it is the equivalent device expression:
    Penalty(x) = x > 0 ? 0 : -2 P p x^(p-1) = -2 P p negpart(x)^(p-1) for p > 1."""

    def __init__(self, Prefactor: float, Exponent: int, Field: str):
        if Exponent < 2:
            raise GopfError("NegativeValuePenalty: the device expression needs Exponent >= 2")
        self.Prefactor, self.Exponent, self.Field = Prefactor, Exponent, Field

    @property
    def Evaluate(self) -> str:
        return f"-2.0*({self.Prefactor!r})*{float(self.Exponent)!r}*negpart({self.Field})^{self.Exponent - 1}"


def NewDefaultNegativeValuePenalty(field: str) -> NegativeValuePenalty:
    return NegativeValuePenalty(1500.0, 3, field)


class TensorialHessian:
    """pf.TensorialHessian (pf/tensorialHessian.go:17-74), implicit."""

    def __init__(self, K, Field: str = ""):
        self.K, self.Field = [float(v) for v in K], Field

    def _register(self, m: "Model", name: str, cls: str):
        if cls != "implicit":
            raise GopfError("TensorialHessian is an implicit term")
        k = (ctypes.c_double * len(self.K))(*self.K)
        check(lib().gopf_model_register_tensorial_hessian(m._h, _s(name), _s(self.Field), k, len(self.K)))


class Peak:
    """pfc.Peak (pfc/pairCorrelation.go:8-13)."""

    def __init__(self, PlaneDensity: float, Location: float, Width: float, NumPlanes: int):
        self.PlaneDensity, self.Location, self.Width, self.NumPlanes = PlaneDensity, Location, Width, NumPlanes


class ReciprocalSpacePairCorrelation:
    """pfc.ReciprocalSpacePairCorrelation (pfc/pairCorrelation.go:18-25)."""

    def __init__(self, EffTemp: float, Peaks: List[Peak]):
        self.EffTemp, self.Peaks = EffTemp, Peaks


def _term_energy(term) -> float:
    m = getattr(term, "_model", None)
    if m is None or not m._solvers:
        raise GopfError("GetEnergy: the term is not registered with a model that has a solver")
    e = ctypes.c_double(0.0)
    check(lib().gopf_solver_term_energy(m._solvers[-1]._h, _s(term._name), ctypes.byref(e)))
    return e.value


class PairCorrlationTerm:
    """pf.PairCorrlationTerm (pf/pairCorrelationTerm.go:22-51), implicit."""

    explicit = False

    def __init__(self, PairCorrFunc: ReciprocalSpacePairCorrelation, Field: str, Prefactor: float, Laplacian: bool = False):
        self.PairCorrFunc, self.Field, self.Prefactor, self.Laplacian = PairCorrFunc, Field, Prefactor, Laplacian

    def _register(self, m: "Model", name: str, cls: str):
        want = "explicit" if self.explicit else "implicit"
        if cls != want:
            raise GopfError(f"{type(self).__name__} is an {want} term")
        pk = self.PairCorrFunc.Peaks
        dbl = lambda vals: (ctypes.c_double * len(vals))(*vals)
        check(lib().gopf_model_register_pair_correlation(
            m._h, _s(name), 1 if self.explicit else 0, _s(self.Field), ctypes.c_double(self.Prefactor),
            1 if self.Laplacian else 0, ctypes.c_double(self.PairCorrFunc.EffTemp), len(pk),
            dbl([p.PlaneDensity for p in pk]), dbl([p.Location for p in pk]), dbl([p.Width for p in pk]),
            int_array([p.NumPlanes for p in pk])))
        self._model, self._name = m, name

    def GetEnergy(self, bricks=None, ft=None, domainSize=None) -> float:
        """PairCorrlationTerm.GetEnergy (pf/pairCorrelationTerm.go:58-84) on the device-resident state
        of the model's solver; the arguments are accepted for signature parity."""
        return _term_energy(self)


class ExplicitPairCorrelationTerm(PairCorrlationTerm):
    """pf.ExplicitPairCorrelationTerm (pf/pairCorrelationTerm.go:89-110)."""

    explicit = True


class IdealMix:
    """pfc.IdealMix (pfc/ideal.go:17-50)."""

    def __init__(self, C3: float, C4: float):
        self.C3, self.C4 = C3, C4


class IdealMixtureTerm:
    """pf.IdealMixtureTerm (pf/pairCorrelationTerm.go:116-193), mixed term."""

    def __init__(self, IdealMix_: IdealMix, Field: str, Prefactor: float, Laplacian: bool = False):
        self.IdealMix, self.Field, self.Prefactor, self.Laplacian = IdealMix_, Field, Prefactor, Laplacian

    def DerivedField(self, num_nodes: int = 0, bricks=None):
        """Token for RegisterMixedTerm's dFields / RegisterDerivedField."""
        return ("ideal_mixture_derived", self)

    def eval_expression(self) -> str:
        """IdealMixtureTerm.Eval as a RegisterFunction expression: Prefactor * IdealMix.Deriv(re(field))."""
        v = f"re({self.Field})"
        c3, c4 = -self.IdealMix.C3 / 6.0, self.IdealMix.C4 / 12.0
        return f"({self.Prefactor!r})*(2.0*0.5*{v}+3.0*({c3!r})*{v}*{v}+4.0*({c4!r})*{v}*{v}*{v})"

    def _register(self, m: "Model", name: str, cls: str, with_derived: bool = False):
        if cls != "mixed":
            raise GopfError("IdealMixtureTerm is a mixed term")
        check(lib().gopf_model_register_ideal_mixture(
            m._h, _s(name), _s(self.Field), ctypes.c_double(self.IdealMix.C3), ctypes.c_double(self.IdealMix.C4),
            ctypes.c_double(self.Prefactor), 1 if self.Laplacian else 0, 1 if with_derived else 0))
        self._model, self._name = m, name

    def GetEnergy(self, bricks=None, nodes=None) -> float:
        """IdealMixtureTerm.GetEnergy (pf/pairCorrelationTerm.go:185-193) on the device-resident state."""
        return _term_energy(self)


class WhiteNoise:
    """pf.WhiteNoise (pf/noise.go:11-23); pass ``noise.Generate`` to RegisterFunction."""

    def __init__(self, Strength: float, seed: int = 0):
        self.Strength, self.seed = Strength, seed

    @property
    def Generate(self):
        return ("white_noise", self)


class ConservativeNoise:
    """pf.ConservativeNoise (pf/noise.go:25-100)."""

    def __init__(self, strength: float, dim: int, unique_prefix: int = 0, seed: int = 0):
        self.Strength, self.Dim, self.UniquePrefix, self.seed = strength, dim, unique_prefix, seed

    def GetCurrentName(self, comp: int) -> str:
        return f"{self.UniquePrefix}_current_{comp}"

    def RequiredDerivedFields(self, num_nodes: int = 0):
        return ("cons_noise_derived", self)

    def _register(self, m: "Model", name: str, cls: str, with_derived: bool = False):
        if cls != "explicit":
            raise GopfError("ConservativeNoise is an explicit term")
        if with_derived:
            check(lib().gopf_model_register_conservative_noise(
                m._h, _s(name), ctypes.c_double(self.Strength), int(self.Dim), ctypes.c_uint32(self.UniquePrefix),
                ctypes.c_uint64(self.seed)))
        else:
            check(lib().gopf_model_register_conservative_noise_term(m._h, _s(name), int(self.Dim),
                                                                    ctypes.c_uint32(self.UniquePrefix)))


def NewConservativeNoise(strength: float, dim: int, unique_prefix: int = 0, seed: int = 0) -> ConservativeNoise:
    return ConservativeNoise(strength, dim, unique_prefix, seed)


class VolumeConservingLP:
    """pf.VolumeConservingLP (pf/volumeConserving.go:3-61)."""

    def __init__(self, fieldName: str, indicator: str, dt: float, numNodes: int):
        self.Field, self.Indicator, self.Dt, self.NumNodes = fieldName, indicator, dt, numNodes

    def _register(self, m: "Model", name: str, cls: str):
        if cls != "explicit":
            raise GopfError("VolumeConservingLP is registered as an explicit term")
        check(lib().gopf_model_register_volume_conserving_lp(m._h, _s(name), _s(self.Field), _s(self.Indicator),
                                                              ctypes.c_double(self.Dt)))


def NewVolumeConservingLP(fieldName: str, indicator: str, dt: float, numNodes: int) -> VolumeConservingLP:
    return VolumeConservingLP(fieldName, indicator, dt, numNodes)


class SquaredGradient:
    """pf.SquaredGradient (pf/squareGradientTerm.go:14-68)."""

    def __init__(self, field: str, domainSize):
        if len(domainSize) not in (2, 3):
            raise GopfError("squaregradient: Domain size has to be of length 2 or 3")
        self.Field, self.Factor = field, 1.0

    def _register(self, m: "Model", name: str, cls: str):
        if cls != "explicit":
            raise GopfError("SquaredGradient is an explicit term")
        check(lib().gopf_model_register_squared_gradient(m._h, _s(name), _s(self.Field), ctypes.c_double(self.Factor)))


def NewSquareGradient(field: str, domainSize) -> SquaredGradient:
    return SquaredGradient(field, domainSize)


class HomogeneousModulusLinElast:
    """pf.HomogeneousModulusLinElast (pf/homoLinElast.go:30-150).  MatProp is an
    ``elasticity.Rank4`` (or 81 doubles), Misfit the 3x3 misfit strain."""

    def __init__(self, fieldName: str, domainSize, matProp, misfit):
        if len(domainSize) not in (2, 3):
            raise GopfError("HomogeneousModulusLinElast: domain size has to be of length 2 or 3")
        self.FieldName = fieldName
        self.Dim = len(domainSize)
        self.MatProp = np.ascontiguousarray(getattr(matProp, "Data", matProp), dtype=np.float64).reshape(81)
        self.Misfit = np.ascontiguousarray(misfit, dtype=np.float64).reshape(9)

    def _register(self, m: "Model", name: str, cls: str):
        if cls != "explicit":
            raise GopfError("HomogeneousModulusLinElast is an explicit term")
        pd = ctypes.POINTER(ctypes.c_double)
        check(lib().gopf_model_register_homogeneous_modulus_lin_elast(
            m._h, _s(name), _s(self.FieldName), self.MatProp.ctypes.data_as(pd), self.Misfit.ctypes.data_as(pd)))


def NewHomogeneousModolus(fieldName: str, domainSize, matProp, misfit) -> HomogeneousModulusLinElast:
    return HomogeneousModulusLinElast(fieldName, domainSize, matProp, misfit)


def voigtIndex(i: int, j: int, dim: int) -> int:
    """pf/chargeTransport.go:151-171."""
    out = ctypes.c_int(0)
    check(lib().gopf_charge_transport_voigt_index(int(i), int(j), int(dim), ctypes.byref(out)))
    return out.value


class ChargeTransport:
    """pf.ChargeTransport (pf/chargeTransport.go:29-146).  ``Conductivity`` is the reference's
    ``func(i int) []float64``; a Go closure cannot run on the device, so it is tabulated once at
    registration (a callable of a node-index array returning (N, n_voigt) or one constant tensor,
    or the (N, n_voigt) array itself).  ``FT`` is accepted for signature parity and unused: the
    device solver's own transform is the term's transform."""

    def __init__(self, Conductivity, ExternalField, Field: str, FT=None):
        self.Conductivity, self.ExternalField, self.Field, self.FT = Conductivity, [float(e) for e in ExternalField], Field, FT
        self._model: Optional["Model"] = None
        self._name = ""

    def _table(self, N: int) -> np.ndarray:
        s = self.Conductivity(np.arange(N)) if callable(self.Conductivity) else self.Conductivity
        s = np.asarray(s, dtype=np.float64)
        if s.ndim == 1:
            s = np.broadcast_to(s, (N, s.shape[0]))
        if s.ndim != 2 or s.shape[0] != N or s.shape[1] not in (3, 6):
            raise GopfError("ChargeTransport: Conductivity must give 3 (2-D) or 6 (3-D) Voigt components per node")
        return np.ascontiguousarray(s.T)  # component-major [n_voigt][N]

    def _register(self, m: "Model", name: str, cls: str):
        if cls != "explicit":
            raise GopfError("ChargeTransport is an explicit term")
        tab = self._table(m.NumNodes())
        ext = np.ascontiguousarray(self.ExternalField, dtype=np.float64)
        pd = ctypes.POINTER(ctypes.c_double)
        check(lib().gopf_model_register_charge_transport(m._h, _s(name), _s(self.Field), tab.ctypes.data_as(pd),
                                                         int(tab.shape[0]), ctypes.c_int64(tab.shape[1]),
                                                         ext.ctypes.data_as(pd), int(ext.shape[0])))
        self._model, self._name = m, name

    def Current(self, density=None, N: Optional[int] = None, realspace: bool = True):
        """ChargeTransport.Current (:121-146) evaluated on the device-resident state of the solver
        the term's model belongs to: a list of ``dim`` real arrays, minus the real part of the current.
        ``density`` / ``realspace`` are accepted for signature parity; the device spectrum of the
        term's field is the density."""
        if self._model is None or not self._model._solvers:
            raise GopfError("ChargeTransport.Current: the term is not registered with a model that has a solver")
        solver = self._model._solvers[-1]
        n = self._model.NumNodes()
        dim = len(self.ExternalField)
        out = np.empty((dim, n), dtype=np.float64)
        check(lib().gopf_solver_charge_current(solver._h, _s(self._name), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return [out[d] for d in range(dim)]

    def OnStepFinished(self, t, bricks=None):
        pass


_TIME_FN = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_double, ctypes.c_void_p)


class Source:
    """pf.Source (pf/sourceTerm.go:13-22): a point source at ``Pos`` with amplitude ``f(t)``."""

    def __init__(self, pos, f: Callable[[float], float]):
        self.Pos = [float(p) for p in pos]
        self.f = f
        self._error: Optional[BaseException] = None

        def call(t, _user):  # an exception cannot cross the C frame: keep it, poison the amplitude
            try:
                return float(f(t))
            except BaseException as e:  # noqa: BLE001 -- re-raised by Solver after the C call returns
                self._error = e
                return float("nan")

        self._cb = _TIME_FN(call)  # kept alive with the Source


def NewSource(pos, f) -> Source:
    return Source(pos, f)


class Vandeven:
    """pf.Vandeven (pf/vandeven.go:8-40): Data is the 1000-point table."""

    def __init__(self, order: int):
        self.Data = np.zeros(1000, dtype=np.float64)
        check(lib().gopf_vandeven_table(int(order), self.Data.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1000))


def NewVandeven(order: int) -> Vandeven:
    return Vandeven(order)


class TableFilter:
    """Any ModalFilter (pf/util.go:120-122) sampled on a uniform table over [0, 1]."""

    def __init__(self, data):
        self.Data = np.ascontiguousarray(data, dtype=np.float64)


# ---- Model ------------------------------------------------------------------------------
class _TermCounts:
    def __init__(self, n_terms, n_denum):
        self.Terms = [None] * n_terms
        self.Denum = [None] * n_denum


class Model:
    """pf.Model (pf/model.go:119-484) over ``gopf_model``."""

    def __init__(self):
        self._h = ctypes.c_void_p()
        check(lib().gopf_model_create(ctypes.byref(self._h)))
        self.Fields: List[Field] = []
        self.Bricks = {}
        self.Equations: List[str] = []
        self._solvers = []
        self._sources: List["Source"] = []  # keeps the ctypes callbacks alive
        self._host_buffers: list = []       # registered Field.Data arrays (+ pin owners): the C model holds their pointers
        self.ImplicitTerms, self.ExplicitTerms, self.MixedTerms = {}, {}, {}  # model.go:120-123

    def AddField(self, f: Field):
        check(lib().gopf_model_add_field(self._h, _s(f.Name), ctypes.c_int64(f.Data.shape[0]),
                                         f.Data.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        self.Fields.append(f)
        self.Bricks[f.Name] = f
        # the registered buffer (and the owner of a pinned allocation) must outlive the C model
        self._host_buffers.append((f.Data, getattr(f, "_pin", None)))
        object.__setattr__(f, "_registered", True)

    def AddScalar(self, s: Scalar):
        check(lib().gopf_model_add_scalar(self._h, _s(s.Name), ctypes.c_double(s.Value.real), ctypes.c_double(s.Value.imag)))
        self.Bricks[s.Name] = s

    def AddEquation(self, eq: str):
        check(lib().gopf_model_add_equation(self._h, _s(eq)))
        self.Equations.append(eq.replace(" ", ""))

    def AddSource(self, eqNo: int, s: "Source"):
        """Model.AddSource (pf/model.go:151-154).  The time function runs on the host, once per
        right-hand-side evaluation."""
        pos = np.ascontiguousarray(s.Pos, dtype=np.float64)
        check(lib().gopf_model_add_source(self._h, int(eqNo), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                          int(pos.shape[0]), s._cb, None))
        self._sources.append(s)

    def RegisterFunction(self, name: str, F):
        """F: expression string, or WhiteNoise(...).Generate."""
        if isinstance(F, tuple) and F[0] == "white_noise":
            check(lib().gopf_model_register_white_noise(self._h, _s(name), ctypes.c_double(F[1].Strength),
                                                        ctypes.c_uint64(F[1].seed)))
        elif isinstance(F, str):
            check(lib().gopf_model_register_function(self._h, _s(name), _s(F)))
        else:
            raise GopfError("RegisterFunction: arbitrary closures cannot run on the device; pass an expression "
                            "string (see include/gopf_cuda.h) or WhiteNoise(...).Generate")

    def SetKSpaceNoise(self, on: bool = True):
        """Draw the spectrum of plain explicit WhiteNoise terms at the k-point instead of transforming a
        real-space noise field every step (include/gopf_cuda.h: gopf_model_set_kspace_noise)."""
        check(lib().gopf_model_set_kspace_noise(self._h, 1 if on else 0))

    def FunctionSource(self, name: str, kernel: bool = False) -> str:
        """C source the registered function `name` compiles to (kernel=True: the CUDA unit for NVRTC)."""
        need = ctypes.c_int64(0)
        check(lib().gopf_model_function_source(self._h, _s(name), 1 if kernel else 0, None, ctypes.c_int64(0),
                                               ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        check(lib().gopf_model_function_source(self._h, _s(name), 1 if kernel else 0, buf, need, None))
        return buf.value.decode("utf-8")

    def FunctionPassSource(self, name: str, line_length: int) -> str:
        """CUDA unit of the forward pass with the registered function compiled into its load."""
        need = ctypes.c_int64(0)
        check(lib().gopf_model_function_pass_source(self._h, _s(name), int(line_length), None, ctypes.c_int64(0), ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        check(lib().gopf_model_function_pass_source(self._h, _s(name), int(line_length), buf, need, None))
        return buf.value.decode("utf-8")

    def FunctionPassCompile(self, name: str, line_length: int):
        """NVRTC-compile it for sm_100a (needs no GPU); returns (cubin size, mangled kernel name)."""
        n = ctypes.c_int64(0)
        low = ctypes.create_string_buffer(512)
        check(lib().gopf_model_function_pass_compile(self._h, _s(name), int(line_length), ctypes.byref(n), low, 512))
        return n.value, low.value.decode()

    def FunctionCompile(self, name: str) -> int:
        """NVRTC-compile the function's kernel for sm_100a (needs no GPU); returns the cubin size."""
        n = ctypes.c_int64(0)
        check(lib().gopf_model_function_compile(self._h, _s(name), ctypes.byref(n)))
        return n.value

    def KUpdateSource(self, dims, dt: float, tab_mask: int = 0, filter_addr: int = 0, filter_n: int = 0, lp_addr: int = 0) -> str:
        """CUDA unit the k-space update of this model is specialised to (inspection, compile checks, tests)."""
        need = ctypes.c_int64(0)
        args = (self._h, len(dims), int_array(dims), ctypes.c_double(dt), ctypes.c_uint(tab_mask), ctypes.c_uint64(filter_addr),
                int(filter_n), ctypes.c_uint64(lp_addr))
        check(lib().gopf_model_kupdate_source(*args, None, ctypes.c_int64(0), ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        check(lib().gopf_model_kupdate_source(*args, buf, need, None))
        return buf.value.decode("utf-8")

    def KUpdateCompile(self, dims, dt: float, tab_mask: int = 0, filter_addr: int = 0, filter_n: int = 0, lp_addr: int = 0) -> int:
        n = ctypes.c_int64(0)
        check(lib().gopf_model_kupdate_compile(self._h, len(dims), int_array(dims), ctypes.c_double(dt), ctypes.c_uint(tab_mask),
                                               ctypes.c_uint64(filter_addr), int(filter_n), ctypes.c_uint64(lp_addr), ctypes.byref(n)))
        return n.value

    def RegisterTableField(self, name: str, values: np.ndarray):
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.ndim != 2:
            raise GopfError("table field: values must be [n_steps][n_nodes]")
        check(lib().gopf_model_register_table_field(self._h, _s(name), values.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                    ctypes.c_int64(values.shape[0])))

    def RegisterDerivedField(self, d):
        if isinstance(d, tuple) and d[0] == "ideal_mixture_derived":
            term = d[1]
            v = f"re({term.Field})"
            c3, c4 = -term.IdealMix.C3 / 6.0, term.IdealMix.C4 / 12.0
            expr = f"3.0*({c3!r})*{v}*{v}+4.0*({c4!r})*{v}*{v}*{v}"  # pairCorrelationTerm.go:144-156
            check(lib().gopf_model_register_function(self._h, _s(f"ideal_mixture_{term.Field}_nonlin"), _s(expr)))
        else:
            raise GopfError("RegisterDerivedField: only catalog derived fields are device-expressible")

    @staticmethod
    def _wants_derived(dfields, tag):
        if dfields is None:
            return False
        items = dfields if isinstance(dfields, list) else [dfields]
        return any(isinstance(x, tuple) and x[0] == tag for x in items)

    def RegisterImplicitTerm(self, name: str, t, dFields=None):
        t._register(self, name, "implicit")
        self.ImplicitTerms[name] = t

    def RegisterExplicitTerm(self, name: str, t, dFields=None):
        if isinstance(t, ConservativeNoise):
            t._register(self, name, "explicit", self._wants_derived(dFields, "cons_noise_derived"))
        else:
            t._register(self, name, "explicit")
        self.ExplicitTerms[name] = t

    def RegisterMixedTerm(self, name: str, t, dFields=None):
        t._register(self, name, "mixed", self._wants_derived(dFields, "ideal_mixture_derived"))
        self.MixedTerms[name] = t

    def Init(self):
        check(lib().gopf_model_init(self._h))

    @property
    def DerivedFieldNames(self) -> List[str]:
        n = ctypes.c_int(0)
        check(lib().gopf_model_num_derived_fields(self._h, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            buf = ctypes.create_string_buffer(256)
            check(lib().gopf_model_derived_field_name(self._h, i, buf, 256))
            out.append(buf.value.decode())
        return out

    def AllFieldNames(self) -> List[str]:
        return [f.Name for f in self.Fields] + self.DerivedFieldNames

    @property
    def RHS(self):
        out = []
        for i in range(len(self.Equations)):
            a, b = ctypes.c_int(0), ctypes.c_int(0)
            check(lib().gopf_model_num_terms(self._h, i, ctypes.byref(a), ctypes.byref(b)))
            out.append(_TermCounts(a.value, b.value))
        return out

    def EqNumber(self, fieldName: str) -> int:
        e = ctypes.c_int(0)
        check(lib().gopf_model_eq_number(self._h, _s(fieldName), ctypes.byref(e)))
        return e.value

    def NumNodes(self) -> int:
        if not self.Fields:
            raise GopfError("Model: No fields added")
        return self.Fields[0].Data.shape[0]

    def close(self):
        for s in list(self._solvers):
            s.close()
        if self._h:
            lib().gopf_model_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def NewModel() -> Model:
    return Model()


# ---- steppers / solver ------------------------------------------------------------------
class _Stepper:
    """pf.TimeStepper view (pf/solver.go:15-19) of the solver's device stepper."""

    def __init__(self, solver: "Solver", name: str):
        self._solver, self.name = solver, name
        self.Dt = solver.Dt

    def SetFilter(self, filt):
        self._solver._set_filter(filt)

    def GetTime(self) -> float:
        t = ctypes.c_double(0.0)
        check(lib().gopf_solver_get_time(self._solver._h, ctypes.byref(t)))
        return t.value

    def Step(self, m=None):
        self._solver.Propagate(1)


class NewtonKrylov:
    """Settings of the non-linear solve inside ImplicitEuler (nonlin.NewtonKrylov in the reference,
    pf/implicitEuler.go:221-229)."""

    def __init__(self, Maxiter=50, StepSize=1e-3, Tol=1e-7, Stencil=6, Restart=30, InnerTol=1e-4, MaxRestarts=4):
        self.Maxiter, self.StepSize, self.Tol, self.Stencil = Maxiter, StepSize, Tol, Stencil
        self.Restart, self.InnerTol, self.MaxRestarts = Restart, InnerTol, MaxRestarts


class ImplicitEuler:
    """pf.ImplicitEuler (pf/implicitEuler.go:20-229).  Assign to ``solver.Stepper`` as in the
    reference (pf/implicitEuler_test.go:193-198); ``FT`` is accepted for signature parity and unused
    (the device solver owns its transform)."""

    def __init__(self, Dt: float, FT=None, NonlinSolver: Optional[NewtonKrylov] = None):
        self.Dt, self.FT, self.NonlinSolver = Dt, FT, NonlinSolver
        self._solver = None

    def _attach(self, solver: "Solver"):
        if abs(self.Dt - solver.Dt) > 0.0:
            raise GopfError("ImplicitEuler.Dt must equal the solver's dt on the device path")
        self._solver = solver
        check(lib().gopf_solver_set_stepper(solver._h, b"implicit_euler"))
        o = self.NonlinSolver or NewtonKrylov()
        check(lib().gopf_solver_set_newton_krylov(solver._h, int(o.Maxiter), ctypes.c_double(o.StepSize), ctypes.c_double(o.Tol),
                                                  int(o.Stencil), int(o.Restart), ctypes.c_double(o.InnerTol), int(o.MaxRestarts)))

    def SetFilter(self, filt):
        self._solver._set_filter(filt)

    def GetTime(self) -> float:
        t = ctypes.c_double(0.0)
        check(lib().gopf_solver_get_time(self._solver._h, ctypes.byref(t)))
        return t.value

    def Step(self, m=None):
        self._solver.Propagate(1)

    @property
    def Converged(self) -> bool:
        c, n = ctypes.c_int(0), ctypes.c_int64(0)
        check(lib().gopf_solver_newton_krylov_status(self._solver._h, ctypes.byref(c), ctypes.byref(n)))
        return bool(c.value)

    @property
    def ResidualEvaluations(self) -> int:
        c, n = ctypes.c_int(0), ctypes.c_int64(0)
        check(lib().gopf_solver_newton_krylov_status(self._solver._h, ctypes.byref(c), ctypes.byref(n)))
        return n.value


class SDDTimeConstants:
    """pf.SDDTimeConstants (pf/sdd.go:14-23)."""

    def __init__(self, Orientation: float = 1.0, DimerLength: float = 1.0):
        self.Orientation, self.DimerLength = Orientation, DimerLength


class SDDMonitor:
    """pf.SDDMonitor (pf/sdd.go:25-53), refreshed from the device after every Propagate."""

    def __init__(self):
        self.MaxForce = self.ForcePowerSpectrum = self.MaxTorque = self.FieldNorm = self.FieldNormChange = 0.0


class SDD:
    """pf.SDD (pf/sdd.go:86-443), the shrinking-dimer saddle-point stepper.  Used like the reference:
    ``sdd = NewSDD(domainSize, model); sdd.Init(...); sdd.Dt = dt; solver.Stepper = sdd``.  The struct's
    exported fields are plain attributes here; they are pushed to the device before every Propagate
    (the reference reads them through the pointer at every step) and CurrentStep / Monitor / the
    orientation are read back after it."""

    _KEYS = ("Alpha", "Dt", "MinDimerLength", "InitDimerLength")

    def __init__(self, domainSize, model: "Model"):
        self.TimeConstants = SDDTimeConstants(1.0, 1.0)
        self.Alpha = 0.5
        self.Dt = 0.0
        self.CurrentStep = 0
        self.MinDimerLength = 0.0
        self.Monitor = SDDMonitor()
        self.InitDimerLength = 0.0
        self._orientation = np.zeros(model.NumNodes() * len(model.Fields), dtype=np.float64)
        self._initialized = False
        self._orientation_dirty = False
        self._solver: Optional["Solver"] = None

    # -- reference API (host arithmetic of :359-427 is plain float math and stays on the host)
    def Init(self, init, final):
        self.SetInitialOrientation(np.concatenate([(b.Data - a.Data).real for a, b in zip(init, final)]))

    def SetInitialOrientation(self, orient):
        orient = np.asarray(orient, dtype=np.float64)
        if orient.shape[0] != self._orientation.shape[0]:
            raise GopfError("Inconsistent length of the passed orientaiton vector")
        self.InitDimerLength = math.sqrt(float(np.dot(orient, orient)))
        self._orientation = orient / self.InitDimerLength
        self._initialized = True
        self._orientation_dirty = True

    def GetTime(self) -> float:
        return float(self.CurrentStep) * self.Dt

    def DimerLength(self, t: float) -> float:
        l = self.InitDimerLength * math.exp(-t / self.TimeConstants.DimerLength)
        return self.MinDimerLength if l < self.MinDimerLength else l

    def RequiredDimerLengthTime(self, l: float) -> float:
        return self.TimeConstants.DimerLength * math.log(self.InitDimerLength / l)

    def SetFilter(self, filt):
        raise GopfError("SDD: Does not support modal filters")

    def Step(self, m=None):
        self._solver.Propagate(1)

    @property
    def orientation(self) -> np.ndarray:
        return self._orientation

    # -- device plumbing
    def _attach(self, solver: "Solver"):
        self._solver = solver
        check(lib().gopf_solver_set_stepper(solver._h, b"sdd"))
        self._orientation_dirty = self._initialized

    def _push(self):
        h = self._solver._h
        if self._orientation_dirty:
            o = np.ascontiguousarray(self._orientation)
            check(lib().gopf_solver_sdd_set_orientation(h, o.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_int64(o.shape[0])))
            self._orientation_dirty = False
        for k in self._KEYS:
            check(lib().gopf_solver_sdd_set(h, _s(k), ctypes.c_double(getattr(self, k))))
        check(lib().gopf_solver_sdd_set(h, b"TimeConstants.Orientation", ctypes.c_double(self.TimeConstants.Orientation)))
        check(lib().gopf_solver_sdd_set(h, b"TimeConstants.DimerLength", ctypes.c_double(self.TimeConstants.DimerLength)))
        check(lib().gopf_solver_sdd_set(h, b"CurrentStep", ctypes.c_double(self.CurrentStep)))

    def _pull(self):
        h = self._solver._h
        v = ctypes.c_double(0.0)

        def get(key):
            check(lib().gopf_solver_sdd_get(h, _s(key), ctypes.byref(v)))
            return v.value

        self.CurrentStep = int(get("CurrentStep"))
        for k in ("MaxForce", "ForcePowerSpectrum", "MaxTorque", "FieldNorm", "FieldNormChange"):
            setattr(self.Monitor, k, get("Monitor." + k))
        if self._initialized:
            out = np.empty_like(self._orientation)
            check(lib().gopf_solver_sdd_get_orientation(h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
            self._orientation = out


def NewSDD(domainSize, model: "Model") -> SDD:
    return SDD(domainSize, model)


class Solver:
    """pf.Solver (pf/solver.go:29-134) over ``gopf_solver``."""

    @property
    def Stepper(self):
        return self._stepper

    @Stepper.setter
    def Stepper(self, st):
        if isinstance(st, (ImplicitEuler, SDD)):
            st._attach(self)
        self._stepper = st

    def __init__(self, m: Model, domainSize, dt: float, device: int = -1):
        self.Model, self.Dt = m, dt
        self.Callbacks: List[Callable] = []
        self.Monitors: list = []
        self.StartEpoch = 0
        self._h = ctypes.c_void_p()
        check(lib().gopf_solver_create(m._h, len(domainSize), int_array(domainSize), ctypes.c_double(dt), device,
                                       ctypes.byref(self._h)))
        m._solvers.append(self)
        self.Stepper = _Stepper(self, "euler")

    # -- reference API
    def AddCallback(self, cb):
        self.Callbacks.append(cb)

    def AddMonitor(self, mon):
        self.Monitors.append(mon)

    def SetStepper(self, name: str):
        check(lib().gopf_solver_set_stepper(self._h, _s(name)))
        self.Stepper = _Stepper(self, name)

    def Propagate(self, nsteps: int):
        """Solver.Propagate on the host Field.Data arrays (upload, steps, download)."""
        if isinstance(self._stepper, SDD):
            self._stepper._push()
        check(lib().gopf_solver_propagate(self._h, int(nsteps)))
        self._reraise_source_errors()
        if isinstance(self._stepper, SDD):
            self._stepper._pull()

    def Solve(self, nepochs: int, nsteps: int):
        for i in range(nepochs):
            self.Propagate(nsteps)
            for cb in self.Callbacks:
                cb(self, i + self.StartEpoch)
            for mon in self.Monitors:
                mon.Add(self.Model.Bricks)

    def _reraise_source_errors(self):
        for src in self.Model._sources:
            if src._error is not None:
                err, src._error = src._error, None
                raise GopfError(f"Source time function raised {err!r}") from err

    # -- device-resident control
    def Upload(self):
        check(lib().gopf_solver_upload(self._h))

    def StepDevice(self, nsteps: int):
        if isinstance(self._stepper, SDD):
            self._stepper._push()
        check(lib().gopf_solver_step(self._h, int(nsteps)))
        self._reraise_source_errors()
        if isinstance(self._stepper, SDD):
            self._stepper._pull()

    def Download(self):
        check(lib().gopf_solver_download(self._h))

    def DownloadReal(self, field_index: int, big_endian: bool = False) -> np.ndarray:
        """Real part of one field from the device-resident state (float64; '>f8' when big_endian)."""
        n = self.Model.Fields[field_index].Data.shape[0]
        out = np.empty(n, dtype=np.float64)
        check(lib().gopf_solver_download_real(self._h, int(field_index), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                              1 if big_endian else 0))
        return out.view(">f8") if big_endian else out

    def DownloadUint8(self, field_index: int):
        """(uint8 array, min, max): RealPartAsUint8 of one field from the device-resident state with
        pfutil.MinReal / MaxReal (pf/util.go:108-117, pf/fileIO.go:31-34), 1 byte per cell over PCIe."""
        n = self.Model.Fields[field_index].Data.shape[0]
        out = np.empty(n, dtype=np.uint8)
        mn, mx = ctypes.c_double(0.0), ctypes.c_double(0.0)
        check(lib().gopf_solver_download_uint8(self._h, int(field_index), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                               ctypes.byref(mn), ctypes.byref(mx)))
        return out, mn.value, mx.value

    def SolveOnDevice(self, nepochs: int, nsteps: int):
        """Solver.Solve with the state resident on the device between epochs: one upload, callbacks
        after every epoch (they read what they need through DownloadReal / Download; host Field.Data
        is NOT refreshed for them), one download at the end."""
        self.Upload()
        for i in range(nepochs):
            self.StepDevice(nsteps)
            for cb in self.Callbacks:
                cb(self, i + self.StartEpoch)
        self.Download()

    def Synchronize(self):
        check(lib().gopf_solver_synchronize(self._h))

    def SetStream(self, stream: int):
        check(lib().gopf_solver_set_stream(self._h, ctypes.c_void_p(stream)))

    def _set_filter(self, filt):
        if filt is None:
            check(lib().gopf_solver_set_filter(self._h, None, 0))
            return
        data = np.ascontiguousarray(filt.Data, dtype=np.float64)
        check(lib().gopf_solver_set_filter(self._h, data.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), data.shape[0]))

    def BlockedLayout(self):
        """(s, active): blocked k-space layout of the fused path on large 3-D grids (s = 0: row-major)."""
        s, a = ctypes.c_int(0), ctypes.c_int(0)
        check(lib().gopf_solver_blocked_layout(self._h, ctypes.byref(s), ctypes.byref(a)))
        return s.value, bool(a.value)

    def FusedForm(self):
        """(form, derived_form) of the fused single-field kernels: form 0 interpreter / 1 polynomial / 2 tabulated;
        derived_form 0 interpreter / 1 integer power / 2 real polynomial (include/gopf_cuda.h)."""
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        check(lib().gopf_solver_fused_form(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    @property
    def IsFused(self) -> bool:
        f = ctypes.c_int(0)
        check(lib().gopf_solver_is_fused(self._h, ctypes.byref(f)))
        return bool(f.value)

    def ForceGeneric(self, on: bool = True):
        check(lib().gopf_solver_force_generic(self._h, 1 if on else 0))

    def SetJit(self, on: bool = True):
        """Compile registered functions with NVRTC into straight-line kernels at first use
        (default: the GOPF_JIT environment variable)."""
        check(lib().gopf_solver_set_jit(self._h, 1 if on else 0))

    def SetJitInPass(self, on: bool = True):
        """With SetJit: registered functions compiled into the load of their first forward pass."""
        check(lib().gopf_solver_set_jit_inpass(self._h, 1 if on else 0))

    def JitKernels(self) -> int:
        n = ctypes.c_int(0)
        check(lib().gopf_solver_jit_kernels(self._h, ctypes.byref(n)))
        return n.value

    def JitLog(self) -> str:
        buf = ctypes.create_string_buffer(4096)
        check(lib().gopf_solver_jit_log(self._h, buf, 4096))
        return buf.value.decode("utf-8", "replace")

    def KernelLaunches(self, reset: bool = False) -> int:
        n = ctypes.c_int64(0)
        check(lib().gopf_solver_kernel_launches(self._h, ctypes.byref(n), 1 if reset else 0))
        return n.value

    def GetSpectrum(self, index: int) -> np.ndarray:
        out = np.empty(self.Model.NumNodes(), dtype=np.complex128)
        check(lib().gopf_solver_get_spectrum(self._h, index, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return out

    def LPMultiplier(self, slot: int = 0) -> float:
        v = ctypes.c_double(0.0)
        check(lib().gopf_solver_lp_multiplier(self._h, slot, ctypes.byref(v)))
        return v.value

    def ProfileBegin(self):
        check(lib().gopf_solver_profile_begin(self._h))

    def ProfileEnd(self):
        n = ctypes.c_int(0)
        check(lib().gopf_solver_profile_end(self._h, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            name = ctypes.create_string_buffer(64)
            ms, launches, nbytes = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_double(0)
            check(lib().gopf_solver_profile_get(self._h, i, name, 64, ctypes.byref(ms), ctypes.byref(launches),
                                                ctypes.byref(nbytes)))
            out.append({"kernel": name.value.decode(), "total_ms": ms.value, "launches": launches.value,
                        "bytes_per_launch": nbytes.value})
        return out

    def close(self):
        if self._h:
            lib().gopf_solver_destroy(self._h)
            self._h = ctypes.c_void_p()
            if self in self.Model._solvers:
                self.Model._solvers.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Float64IO:
    """pf.Float64IO (pf/fileIO.go:47-62): big-endian float64 of the real part of every field, one
    file per field and epoch, named <prefix>_<field>_<epoch>.bin.  ``from_device`` = True reads the
    device-resident state (for Solver.SolveOnDevice callbacks); False writes host Field.Data."""

    def __init__(self, prefix: str, from_device: bool = False):
        self.Prefix, self.from_device = prefix, from_device

    def SaveFields(self, s: "Solver", epoch: int):
        for i, f in enumerate(s.Model.Fields):
            fname = f"{self.Prefix}_{f.Name}_{epoch}.bin"
            if self.from_device:
                s.DownloadReal(i, big_endian=True).tofile(fname)  # bytes are already big-endian
            else:
                np.ascontiguousarray(f.Data.real).astype(">f8").tofile(fname)


def RealPartAsUint8(data: np.ndarray, mn: float, mx: float) -> np.ndarray:
    """pf.RealPartAsUint8 (pf/util.go:108-117) on a host array."""
    if abs(mx - mn) < 1e-10:
        mx = mn + 1.0
    return ((255.0 * (np.asarray(data).real - mn)) / (mx - mn)).astype(np.uint8)


class Uint8IO:
    """pf.Uint8IO (pf/fileIO.go:15-44): every field's real part scaled to 0..255, one file per field and
    epoch, named <prefix>_<field>_<epoch>.bin.  ``from_device`` = True quantises on the device
    (Solver.SolveOnDevice callbacks); False uses host Field.Data."""

    def __init__(self, prefix: str, from_device: bool = False):
        self.Prefix, self.from_device = prefix, from_device

    def SaveFields(self, s: "Solver", epoch: int):
        for i, f in enumerate(s.Model.Fields):
            fname = f"{self.Prefix}_{f.Name}_{epoch}.bin"
            if self.from_device:
                s.DownloadUint8(i)[0].tofile(fname)
            else:
                re = f.Data.real
                RealPartAsUint8(f.Data, float(re.min()), float(re.max())).tofile(fname)


def NewUint8IO(prefix: str, from_device: bool = False) -> Uint8IO:
    return Uint8IO(prefix, from_device)


def NewFloat64IO(prefix: str, from_device: bool = False) -> Float64IO:
    return Float64IO(prefix, from_device)


def LoadFloat64(fname: str) -> np.ndarray:
    """pf.LoadFloat64 (pf/fileIO.go:66-83)."""
    return np.fromfile(fname, dtype=">f8").astype(np.float64)


def NewSolver(m: Model, domainSize, dt: float, device: int = -1) -> Solver:
    """pf.NewSolver (pf/solver.go:40-62)."""
    return Solver(m, domainSize, dt, device)
