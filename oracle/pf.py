"""Oracle restatement of the reference's ``pf`` package pieces on the hot path:
equation parser, Model, Euler, RK4, Solver.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations: /root/reference.

Arrays are 1-D numpy complex128 vectors that are always mutated IN PLACE so the
aliasing the reference relies on (``m.Fields[i].Data`` and ``m.Bricks[name]``
share one backing array, pf/model.go:142-143) is preserved.
"""
from __future__ import annotations

import math
import re
from typing import Callable, Dict, List, Optional

import numpy as np

from . import pfutil

# ==========================================================================
# Go regexp helpers
# ==========================================================================


def go_find_all(pattern: str, s: str) -> List[str]:
    """Go ``regexp.FindAllString(s, -1)`` semantics (regexp/regexp.go allMatches):
    successive leftmost matches; an EMPTY match that abuts the preceding match is
    dropped.  (Python's ``re.findall`` keeps those, which would hand empty factor
    names to DerivedFieldCalcFromDesc.)"""
    rx = re.compile(pattern)
    out: List[str] = []
    pos_, prev_end, end = 0, -1, len(s)
    while pos_ <= end:
        m = rx.search(s, pos_)
        if m is None:
            break
        accept = True
        if m.end() == pos_:  # empty match at the search position
            if m.start() == prev_end:
                accept = False
            pos_ = pos_ + 1
        else:
            pos_ = m.end()
        prev_end = m.end()
        if accept:
            out.append(m.group(0))
    return out


def go_find_string(pattern: str, s: str) -> str:
    m = re.compile(pattern).search(s)
    return m.group(0) if m else ""


# ==========================================================================
# pf/util.go string helpers
# ==========================================================================
_POWER_RE = re.compile(r"\^(-?\d+\.?\d*)")


def get_power(pattern: str) -> float:
    """pf/util.go:67-80 GetPower."""
    m = _POWER_RE.search(pattern)
    if m is None:
        return 1.0
    return float(m.group(1))


def sort_factors(expr: str) -> str:
    """pf/util.go:296-300 SortFactors (Go sort.Strings = bytewise order)."""
    parts = expr.split("*")
    parts.sort(key=lambda x: x.encode())
    return "*".join(parts)


def get_field_name(term: str, field_names: List[str]) -> str:
    """pf/util.go:82-104 GetFieldName."""
    field = ""
    for f in field_names:
        if f in term:
            without = term.replace(f, "")
            ok = True
            for f1 in field_names:
                if f1 in without:
                    ok = False
                    break
            if ok and len(f) > len(field):
                field = f
    return field


class SubStringDelimiter:
    """pf/util.go:134-139."""

    def __init__(self, SubString: str, PreceedingDelimiter: str = ""):
        self.SubString = SubString
        self.PreceedingDelimiter = PreceedingDelimiter

    def __repr__(self):
        return f"SubStringDelimiter({self.SubString!r}, {self.PreceedingDelimiter!r})"


def _get_first_delimiter(value: str, delimiters: List[str]) -> str:
    for d in delimiters:
        if value[0:1] == d:
            return d
    return ""


def split_on_many(value: str, delimiters: List[str]) -> List[SubStringDelimiter]:
    """pf/util.go:152-199 SplitOnMany (breadth-first queue, first listed delimiter
    that occurs is split on first)."""
    substrings: List[SubStringDelimiter] = []
    queue = [SubStringDelimiter(value, _get_first_delimiter(value, delimiters))]
    all_delims = "".join(delimiters)
    while queue:
        cur = queue.pop(0)
        if not any(ch in all_delims for ch in cur.SubString):
            substrings.append(cur)
            continue
        delim = delimiters[0]
        for d in delimiters:
            if d in cur.SubString:
                delim = d
                break
        splits = [s for s in cur.SubString.split(delim) if s != ""]
        queue.append(SubStringDelimiter(splits[0], cur.PreceedingDelimiter))
        for s in splits[1:]:
            queue.append(SubStringDelimiter(s, delim))
    return substrings


def is_bilinear(term: str, field: str, field_names: List[str]) -> bool:
    """pf/rhsBuilder.go:69-106 isBilinear.  Field names are used as raw regular
    expressions exactly as the reference does."""
    if len(go_find_all(field, term)) != 1:
        return False
    for f in field_names:
        if f == field:
            continue
        if len(go_find_all(f, term)) > 0:
            return False
    res = go_find_string(field + r"*[^/\*]*", term)
    m = _POWER_RE.search(res)
    if m is None:
        return True
    try:
        power = float(m.group(1))
    except ValueError:
        return True
    return abs(power - 1.0) < 1e-10


def get_non_linear_field_expressions(pattern: str, field: str, field_names: List[str]) -> str:
    """pf/util.go:17-36 GetNonLinearFieldExpressions."""
    expr = ""
    for fn in field_names:
        if fn == field and is_bilinear(pattern, field, field_names):
            continue
        res = go_find_string(fn + r"[^\*]*", pattern)
        if res != "":
            expr += res + "*"
    if len(expr) > 1:
        return expr[:-1]
    return expr


def field_name_from_leibniz(leibniz: str) -> str:
    """pf/rhsBuilder.go:58-66."""
    if len(leibniz) <= 3:
        raise ValueError("rhsbuilder: Length of the Leibniz formatted string has to be at least 3")
    if leibniz[0:1] != "d" or leibniz[-3:] != "/dt":
        raise ValueError("rhsbuilder: Passed string is not a leibniz formatted string")
    return leibniz[1:-3]


def known_prefixes() -> List[str]:
    """pf/rhsBuilder.go:244-252."""
    return ["-", "LAP^4", "LAP^2", "LAP", "*"]


def get_known_prefixes(s: str) -> List[str]:
    """pf/rhsBuilder.go:259-272."""
    pref: List[str] = []
    prefixes = known_prefixes()
    while prefixes:
        p = prefixes.pop(0)
        if s.startswith(p):
            pref.append(p)
            prefixes = known_prefixes()
            s = s[len(p):]
    return pref


def remove_known_prefixes(s: str) -> str:
    """pf/rhsBuilder.go:254-257, 274-287 (recursiveRemove)."""
    prefixes = known_prefixes()
    while prefixes:
        p = prefixes.pop(0)
        if s.startswith(p):
            s = s[len(p):]
            prefixes = known_prefixes()
    return s


def panic_on_prefix_in_name(name: str):
    """pf/model.go:479-484 (log.Fatalf in the reference -> exception here)."""
    if len(get_known_prefixes(name)) > 0:
        raise ValueError(f"The words {known_prefixes()} are reserved. Do not include them in your variable names")


# ==========================================================================
# Frequency wrapper
# ==========================================================================
class Frequency:
    """``Frequency func(i int) []float64`` (pf/diffOp.go:9-10) plus a cached
    (N, dim) table so term loops vectorise."""

    def __init__(self, fn: Callable[[int], List[float]], table_fn: Optional[Callable[[int], np.ndarray]] = None):
        self.fn = fn
        self._table_fn = table_fn
        self._cache: Dict[int, np.ndarray] = {}

    def __call__(self, i: int):
        return self.fn(i)

    def table(self, n: int) -> np.ndarray:
        t = self._cache.get(n)
        if t is None:
            if self._table_fn is not None:
                t = self._table_fn(n)
            else:
                t = np.array([self.fn(i) for i in range(n)], dtype=np.float64).reshape(n, -1)
            self._cache[n] = t
        return t


def as_frequency(f) -> Frequency:
    if isinstance(f, Frequency):
        return f
    owner = getattr(f, "__self__", None)
    if owner is not None and hasattr(owner, "freq_table"):
        def tab(n, owner=owner):
            t = owner.freq_table()
            return t if n == t.shape[0] else np.array([owner.Freq(i) for i in range(n)], dtype=np.float64)
        return Frequency(f, tab)
    return Frequency(f)


# ==========================================================================
# pf/diffOp.go
# ==========================================================================
class LaplacianN:
    """pf/diffOp.go:19-30: ft[i] *= (-(2 pi |f(i)|)^2)^Power."""

    def __init__(self, Power: int):
        self.Power = int(Power)

    def multiplier(self, freq, n: int) -> np.ndarray:
        f = as_frequency(freq).table(n)
        norm = np.sqrt(np.sum(f * f, axis=1))  # gonum floats.Norm(f, 2)
        base = -np.power(2.0 * math.pi * norm, 2.0)
        return np.power(base, float(self.Power))

    def Eval(self, freq, ft: np.ndarray) -> np.ndarray:
        ft *= self.multiplier(freq, ft.shape[0])
        return ft


# ==========================================================================
# pf/model.go bricks
# ==========================================================================
class Field:
    """pf/model.go:16-57."""

    def __init__(self, name: str, N: int, data: Optional[np.ndarray] = None):
        panic_on_prefix_in_name(name)
        if data is None:
            self.Data = np.zeros(N, dtype=np.complex128)
        else:
            data = np.asarray(data)
            if data.dtype != np.complex128:
                data = data.astype(np.complex128)
            if data.shape[0] != N:
                raise ValueError("model: Inconsistent length of data")
            self.Data = data
        self.Name = name

    def Get(self, i):
        return self.Data[i]

    def Copy(self) -> "Field":
        f = Field(self.Name, self.Data.shape[0], None)
        f.Data[:] = self.Data
        return f


def NewField(name, N, data=None) -> Field:
    return Field(name, N, data)


class DerivedField:
    """pf/model.go:62-78."""

    def __init__(self, Data: np.ndarray, Name: str, Calc: Callable[[np.ndarray], None]):
        self.Data = Data
        self.Name = Name
        self.Calc = Calc

    def Get(self, i):
        return self.Data[i]

    def Update(self):
        self.Calc(self.Data)


class Scalar:
    """pf/model.go:86-111."""

    def __init__(self, Name: str, Value: complex):
        self.Name = Name
        self.Value = complex(Value)

    def Get(self, i):
        if isinstance(i, np.ndarray):
            return np.full(i.shape, self.Value, dtype=np.complex128)
        return self.Value


def NewScalar(name: str, value: complex) -> Scalar:
    panic_on_prefix_in_name(name)
    return Scalar(name, value)


class RHS:
    """pf/rhsBuilder.go:19-22."""

    def __init__(self):
        self.Terms: List[Callable] = []
        self.Denum: List[Callable] = []


def derived_field_calc_from_desc(desc: str, fields: List[Field]):
    """pf/util.go:38-65 DerivedFieldCalcFromDesc."""
    field_map = {f.Name: f for f in fields}
    res = go_find_all(r"[^\*]*", desc)
    names = [go_find_string(r"^[^\^]*", r) for r in res]
    powers = [get_power(r) for r in res]

    def calc(data: np.ndarray):
        data[:] = 1.0
        for nm, p in zip(names, powers):
            data *= pfutil.go_cpow(field_map[nm].Data, p)

    return calc


# ==========================================================================
# Model
# ==========================================================================
class Model:
    """pf/model.go:119-484."""

    def __init__(self):
        self.Fields: List[Field] = []
        self.DerivedFields: List[DerivedField] = []
        self.Bricks: Dict[str, object] = {}
        self.ImplicitTerms: Dict[str, object] = {}
        self.ExplicitTerms: Dict[str, object] = {}
        self.MixedTerms: Dict[str, object] = {}
        self.Equations: List[str] = []
        self.RHS: List[RHS] = []
        self.AllSources: List[list] = []
        self.RHSModifiers: List[tuple] = []

    # model.go:141-149
    def AddField(self, f: Field):
        self.Fields.append(f)
        self.Bricks[f.Name] = f

    def AddScalar(self, s: Scalar):
        self.Bricks[s.Name] = s

    # model.go:151-154
    def AddSource(self, eq_no: int, s):
        self.AllSources[eq_no].append(s)

    # model.go:157-162
    def AddEquation(self, eq: str):
        eq = eq.replace(" ", "")
        self.Equations.append(eq)
        self.UpdateDerivedFields(eq)
        self.AllSources.append([])

    # model.go:165-195
    def UpdateDerivedFields(self, eq: str):
        rhs = eq.split("=")[1]
        splitted = [s.SubString for s in split_on_many(rhs, ["+", "-"])]
        field = field_name_from_leibniz(eq.split("=")[0])
        field_names = [f.Name for f in self.Fields]
        for s in splitted:
            if self.IsUserDefinedTerm(s):
                continue
            new_fields = sort_factors(get_non_linear_field_expressions(s, field, field_names))
            if new_fields != "" and not self.IsFieldName(new_fields):
                d = DerivedField(
                    np.zeros(self.Fields[0].Data.shape[0], dtype=np.complex128),
                    new_fields,
                    derived_field_calc_from_desc(new_fields, self.Fields),
                )
                self.DerivedFields.append(d)
                self.Bricks[new_fields] = d

    def AllFieldNames(self) -> List[str]:
        return [f.Name for f in self.Fields] + [f.Name for f in self.DerivedFields]

    def IsFieldName(self, name: str) -> bool:
        return any(f.Name == name for f in self.Fields) or any(f.Name == name for f in self.DerivedFields)

    def IsBrickName(self, name: str) -> bool:
        return name in self.Bricks

    # model.go:237-241
    def SyncDerivedFields(self):
        for f in self.DerivedFields:
            f.Update()

    # model.go:244-260
    def Init(self):
        self.RHS = []
        for eq in self.Equations:
            self.RHS.append(Build(eq, self))
        self.SyncDerivedFields()
        for k, v in self.ImplicitTerms.items():
            if not is_implicit(v, self.Bricks, self.NumNodes(), lambda i: [0.4, 0.4]):
                raise RuntimeError(
                    f"Model: Term {k} is not implicit (e.g. it varies when the fields are varied)")

    def NumNodes(self) -> int:
        if not self.Fields:
            raise RuntimeError("Model: No fields added")
        return self.Fields[0].Data.shape[0]

    # model.go:283-303
    def GetRHS(self, field_no: int, freq, t: float) -> np.ndarray:
        n = self.Fields[field_no].Data.shape[0]
        data = np.zeros(n, dtype=np.complex128)
        tmp = np.zeros(n, dtype=np.complex128)
        for f in self.RHS[field_no].Terms:
            f(freq, t, tmp)
            data += tmp
        for s in self.AllSources[field_no]:
            s.Eval(freq, t, tmp)
            data += tmp
        for eq_no, mod in self.RHSModifiers:
            if eq_no == field_no:
                mod(data)
        return data

    # model.go:306-314
    def GetDenum(self, field_no: int, freq, t: float) -> np.ndarray:
        n = self.Fields[field_no].Data.shape[0]
        data = np.zeros(n, dtype=np.complex128)
        tmp = np.zeros(n, dtype=np.complex128)
        for f in self.RHS[field_no].Denum:
            f(freq, t, tmp)
            data += tmp
        return data

    # model.go:322-376
    def _register_derived_fields(self, d_fields):
        if d_fields:
            for f in d_fields:
                if not self.IsFieldName(f.Name):
                    self.DerivedFields.append(f)
                    self.Bricks[f.Name] = f

    def RegisterImplicitTerm(self, name, t, d_fields=None):
        panic_on_prefix_in_name(name)
        self.ImplicitTerms[name] = t
        self._register_derived_fields(d_fields)

    def RegisterExplicitTerm(self, name, t, d_fields=None):
        panic_on_prefix_in_name(name)
        self.ExplicitTerms[name] = t
        self._register_derived_fields(d_fields)

    def RegisterMixedTerm(self, name, t, d_fields=None):
        panic_on_prefix_in_name(name)
        self.MixedTerms[name] = t
        self._register_derived_fields(d_fields)

    def IsImplicitTerm(self, desc):
        return desc in self.ImplicitTerms

    def IsExplicitTerm(self, desc):
        return desc in self.ExplicitTerms

    def IsMixedTerm(self, desc):
        return desc in self.MixedTerms

    def IsUserDefinedTerm(self, desc):
        return self.IsImplicitTerm(desc) or self.IsExplicitTerm(desc) or self.IsMixedTerm(desc)

    # model.go:400-418.  F is vectorised here: F(i_array, bricks) -> array.
    def RegisterFunction(self, name: str, F):
        panic_on_prefix_in_name(name)
        n = self.Fields[0].Data.shape[0]
        idx = np.arange(n)

        def calc(out: np.ndarray):
            out[:] = F(idx, self.Bricks)

        self.RegisterDerivedField(DerivedField(np.zeros(n, dtype=np.complex128), name, calc))

    def RegisterDerivedField(self, d: DerivedField):
        self.DerivedFields.append(d)
        self.Bricks[d.Name] = d

    # model.go:441-455
    def EqNumber(self, field_name: str) -> int:
        rx = re.compile(r"d(.*?)/dt")
        for i, eq in enumerate(self.Equations):
            m = rx.search(eq)
            if m and m.group(1) == field_name:
                return i
        raise RuntimeError(f"EqNumber: Could not find an equation for field {field_name}")

    # model.go:473-478
    def RegisterRHSModifier(self, eq_number: int, modifier):
        self.RHSModifiers.append((eq_number, modifier))


def NewModel() -> Model:
    return Model()


# ==========================================================================
# pf/rhsBuilder.go
# ==========================================================================
def valid_name(name: str, model: Model) -> bool:
    """pf/rhsBuilder.go:109-122."""
    stripped = name.replace(" ", "")
    if stripped in ("", "LAP"):
        return True
    if stripped[:3] == "LAP":
        stripped = stripped[3:]
    return model.IsBrickName(stripped) or model.IsFieldName(stripped)


def concrete_term(term_delim: SubStringDelimiter, m: Model):
    """pf/rhsBuilder.go:125-190 ConcreteTerm."""
    term = term_delim.SubString
    sign = -1.0 if term_delim.PreceedingDelimiter == "-" else 1.0

    res = go_find_all(r"[^\*]*", term)
    brick_names: List[str] = []
    powers: List[float] = []
    for r in res:
        name = go_find_string(r"^[^\^]*", r)
        if (not m.IsFieldName(name)) and m.IsBrickName(name):
            brick_names.append(name)
            powers.append(get_power(r))
        elif not valid_name(name, m):
            raise ValueError(f"rhsBuilder: Name {name} is not defined!")

    field_name = get_field_name(sort_factors(term), m.AllFieldNames())

    lap = None
    if "LAP" in term:
        lap = LaplacianN(int(get_power(go_find_string(r"LAP*[^a-zA-Z]*", term))))

    def fn(freq, t, field: np.ndarray):
        idx = np.arange(field.shape[0])
        field[:] = complex(sign, 0.0)
        for bn, p in zip(brick_names, powers):
            field *= pfutil.go_cpow(np.asarray(m.Bricks[bn].Get(idx)), p)
        if field_name != "":
            field *= m.Bricks[field_name].Get(idx)
        if lap is not None:
            lap.Eval(freq, field)

    return fn


def construct_func(term, prefixes: List[str]):
    """pf/rhsBuilder.go:199-242 constructFunc."""
    if len(prefixes) == 0:
        return term
    p = prefixes[0]
    if p == "-":
        def f(freq, t, field, term=term):
            term(freq, t, field)
            field *= -1.0
    elif p in ("LAP^4", "LAP^2", "LAP"):
        power = {"LAP^4": 4, "LAP^2": 2, "LAP": 1}[p]

        def f(freq, t, field, term=term, power=power):
            term(freq, t, field)
            LaplacianN(power).Eval(freq, field)
    else:
        # " ", "+" do nothing; anything else (incl. "*", "") logs 'Unrecognized prefix'
        f = term
    return construct_func(f, prefixes[1:])


def Build(eq: str, m: Model) -> RHS:
    """pf/rhsBuilder.go:26-54 Build."""
    sides = eq.split("=")
    if len(sides) != 2:
        raise ValueError("build: equality sign can only occur once")
    field = field_name_from_leibniz(sides[0])
    rhs = RHS()
    for t in split_on_many(sides[1], ["+", "-"]):
        name = remove_known_prefixes(t.SubString)
        prefixes = get_known_prefixes(t.SubString)
        prefixes.append(t.PreceedingDelimiter)
        if m.IsImplicitTerm(name):
            rhs.Denum.append(construct_func(m.ImplicitTerms[name].Construct(m.Bricks), prefixes))
        elif m.IsExplicitTerm(name):
            rhs.Terms.append(construct_func(m.ExplicitTerms[name].Construct(m.Bricks), prefixes))
        elif m.IsMixedTerm(name):
            rhs.Denum.append(construct_func(m.MixedTerms[name].ConstructLinear(m.Bricks), prefixes))
            rhs.Terms.append(construct_func(m.MixedTerms[name].ConstructNonLinear(m.Bricks), prefixes))
        elif is_bilinear(t.SubString, field, m.AllFieldNames()):
            t.SubString = t.SubString.replace(field, "")
            rhs.Denum.append(concrete_term(t, m))
        else:
            rhs.Terms.append(concrete_term(t, m))
    return rhs


# ==========================================================================
# pf/userDefinedTerm.go:79-112 isImplicit
# ==========================================================================
class _PerturbedBrick:
    def __init__(self, parent, perturbation):
        self.parent = parent
        self.perturbation = perturbation

    def Get(self, i):
        return self.parent.Get(i) + self.perturbation


def is_implicit(t, bricks, N: int, freq) -> bool:
    tmp = {k: _PerturbedBrick(b, complex(0.2, 0.0)) for k, b in bricks.items()}
    f1 = t.Construct(bricks)
    f2 = t.Construct(tmp)
    a1 = np.zeros(N, dtype=np.complex128)
    a2 = np.zeros(N, dtype=np.complex128)
    f1(freq, 0.0, a1)
    f2(freq, 0.0, a2)
    return pfutil.cmplx_equal_approx(a1, a2, 1e-10)


# ==========================================================================
# pf/util.go:120-132 modal filter
# ==========================================================================
def RealPartAsUint8(data: np.ndarray, mn: float, mx: float) -> np.ndarray:
    """pf/util.go:108-117: the real part scaled such that min -> 0 and max -> 255, truncated."""
    if abs(mx - mn) < 1e-10:
        mx = mn + 1.0
    return ((255.0 * (np.asarray(data).real - mn)) / (mx - mn)).astype(np.uint8)


def uint8_payload(data: np.ndarray) -> np.ndarray:
    """Uint8IO.SaveFields for one field (pf/fileIO.go:31-34): pfutil.MinReal / MaxReal
    (pfutil/sliceOperations.go:60-81), then RealPartAsUint8."""
    re_ = np.asarray(data).real
    return RealPartAsUint8(data, float(re_.min()), float(re_.max()))


def apply_modal_filter(filt, freq, data: np.ndarray):
    f = as_frequency(freq).table(data.shape[0])
    f_rad = np.sqrt(np.sum(f * f, axis=1))  # sqrt(pfutil.Dot(f, f))
    value = f_rad * 2.0 / math.pi
    if hasattr(filt, "eval_array"):
        data *= filt.eval_array(value)
    else:
        data *= np.array([filt.Eval(float(v)) for v in value])


# ==========================================================================
# pf/euler.go
# ==========================================================================
class Euler:
    """pf/euler.go:6-64 semi-implicit Euler."""

    def __init__(self, Dt: float, FT, Filter=None):
        self.Dt = Dt
        self.FT = FT
        self.Filter = Filter
        self.CurrentStep = 0

    def Step(self, m: Model):
        c_dt = complex(self.Dt, 0.0)
        m.SyncDerivedFields()
        for f in m.Fields:
            self.FT.FFT(f.Data)
        for f in m.DerivedFields:
            self.FT.FFT(f.Data)
        t = self.GetTime()
        freq = as_frequency(self.FT.Freq)
        for i in range(len(m.Fields)):
            rhs = m.GetRHS(i, freq, t)
            denum = m.GetDenum(i, freq, t)
            d = m.Fields[i].Data
            d[:] = (d + c_dt * rhs) / (complex(1.0, 0.0) - c_dt * denum)  # euler.go:33
            if self.Filter is not None:
                apply_modal_filter(self.Filter, freq, d)
        for f in m.Fields:
            self.FT.IFFT(f.Data)
            pfutil.div_real_scalar(f.Data, float(f.Data.shape[0]))
        self.CurrentStep += 1

    def GetTime(self) -> float:
        return float(self.CurrentStep) * self.Dt

    def Propagate(self, nsteps: int, m: Model):
        for _ in range(nsteps):
            self.Step(m)

    def SetFilter(self, filt):
        self.Filter = filt


# ==========================================================================
# pf/implicitEuler.go
# ==========================================================================
FD_STENCILS = {  # central-difference weights of d/dh at offsets +-1, +-2, +-3 (antisymmetric)
    2: (0.5,),
    4: (2.0 / 3.0, -1.0 / 12.0),
    6: (0.75, -0.15, 1.0 / 60.0),
}


class NewtonKrylov:
    """Jacobian-free Newton-Krylov solver for F(x) = 0.

    PARITY UNPINNED: the reference delegates this to github.com/davidkleiven/gononlin v0.2.2
    ``nonlin.NewtonKrylov`` with gonum.org/v1/exp ``linsolve.GMRES`` (pf/implicitEuler.go:190-201,
    221-229); neither module is under /root/reference.  What the reference fixes are the settings
    (Maxiter 50, StepSize 1e-3, Tol 1e-7, Stencil 6, un-restarted GMRES) and, through its tests,
    the result to 5e-3 / 1e-3 against analytic solutions (pf/implicitEuler_test.go:31,43,55,67,208).
    This class is the published algorithm those modules implement, stated once so that the CUDA
    stepper (gopf_b200/csrc/implicit_euler.cu) can mirror it choice for choice:

      Newton:  x <- x + s,  J(x) s = -F(x), until max|F(x)| < Tol or Maxiter iterations
      J v   ~  sum_k w_k (F(x + k eps v) - F(x - k eps v)) / eps,  central stencil of `Stencil`
               points, eps = StepSize * sqrt(len(x)) / |v|_2 (RMS perturbation = StepSize)
      GMRES:   modified Gram-Schmidt Arnoldi, Givens rotations, restart `Restart`, stop at
               |r| <= InnerTol * |F(x)| or after MaxRestarts cycles
    Two converged runs agree to the solver tolerance, not to 1e-10.
    """

    def __init__(self, Maxiter=50, StepSize=1e-3, Tol=1e-7, Stencil=6, Restart=30, InnerTol=1e-4, MaxRestarts=4):
        self.Maxiter, self.StepSize, self.Tol, self.Stencil = Maxiter, StepSize, Tol, Stencil
        self.Restart, self.InnerTol, self.MaxRestarts = Restart, InnerTol, MaxRestarts
        self.residual_evaluations = 0

    def jac_vec(self, F, x, v, fx_unused=None):
        nv = float(np.linalg.norm(v))
        if nv == 0.0:
            return np.zeros_like(v)
        eps = self.StepSize * math.sqrt(x.shape[0]) / nv
        out = np.zeros_like(x)
        for k, w in enumerate(FD_STENCILS[self.Stencil], start=1):
            out += w * (F(x + (k * eps) * v) - F(x - (k * eps) * v))
            self.residual_evaluations += 2
        return out / eps

    def gmres(self, F, x, b):
        """Solves J(x) s = b; returns s."""
        n = x.shape[0]
        s = np.zeros(n)
        bnorm = float(np.linalg.norm(b))
        if bnorm == 0.0:
            return s
        target = self.InnerTol * bnorm
        for _cycle in range(self.MaxRestarts):
            r = b - self.jac_vec(F, x, s) if np.any(s) else b.copy()
            beta = float(np.linalg.norm(r))
            if beta <= target:
                break
            m = self.Restart
            V = np.zeros((m + 1, n))
            H = np.zeros((m + 1, m))
            cs, sn, g = np.zeros(m), np.zeros(m), np.zeros(m + 1)
            V[0] = r / beta
            g[0] = beta
            k_used = 0
            for j in range(m):
                w = self.jac_vec(F, x, V[j])
                for i in range(j + 1):
                    H[i, j] = float(np.dot(w, V[i]))
                    w -= H[i, j] * V[i]
                H[j + 1, j] = float(np.linalg.norm(w))
                if H[j + 1, j] > 0.0:
                    V[j + 1] = w / H[j + 1, j]
                for i in range(j):
                    t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                    H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                    H[i, j] = t
                d = math.hypot(H[j, j], H[j + 1, j])
                cs[j], sn[j] = (H[j, j] / d, H[j + 1, j] / d) if d > 0.0 else (1.0, 0.0)
                H[j, j] = cs[j] * H[j, j] + sn[j] * H[j + 1, j]
                H[j + 1, j] = 0.0
                g[j + 1] = -sn[j] * g[j]
                g[j] = cs[j] * g[j]
                k_used = j + 1
                if abs(g[j + 1]) <= target or H[j, j] == 0.0:
                    break
            y = np.zeros(k_used)
            for i in range(k_used - 1, -1, -1):
                acc = g[i] - float(np.dot(H[i, i + 1:k_used], y[i + 1:]))
                y[i] = acc / H[i, i] if H[i, i] != 0.0 else 0.0
            s += V[:k_used].T @ y
            if abs(g[k_used]) <= target:
                break
        return s

    def Solve(self, F, x0):
        x = np.array(x0, dtype=np.float64)
        converged = False
        for _ in range(self.Maxiter):
            fx = F(x)
            self.residual_evaluations += 1
            if float(np.max(np.abs(fx))) < self.Tol:
                converged = True
                break
            x = x + self.gmres(F, x, -fx)
        return x, converged


class ImplicitEuler:
    """pf/implicitEuler.go:20-207: exponential-integrator implicit step solved by Newton-Krylov,
    one semi-implicit Euler step as the initial guess."""

    def __init__(self, Dt: float, FT, Filter=None, NonlinSolver: Optional["NewtonKrylov"] = None):
        self.Dt = Dt
        self.FT = FT
        self.Filter = Filter
        self.CurrentStep = 0
        self.NonlinSolver = NonlinSolver
        self.last_converged = True

    def GetTime(self) -> float:
        return self.Dt * float(self.CurrentStep)

    def SetFilter(self, filt):
        self.Filter = filt  # implicitEuler.go:209-213: only the predictor sees it

    def fft(self, m: Model):
        m.SyncDerivedFields()
        for f in m.Fields:
            self.FT.FFT(f.Data)
        for f in m.DerivedFields:
            self.FT.FFT(f.Data)

    def ifft(self, m: Model):
        for f in m.Fields:
            self.FT.IFFT(f.Data)
            pfutil.div_real_scalar(f.Data, float(f.Data.shape[0]))

    def nonlinearIntegral(self, denum, rhs, rhs_prev):
        """:151-162, vectorised."""
        c_dt = complex(self.Dt, 0.0)
        a = rhs_prev
        b = (rhs - rhs_prev) / c_dt
        f = np.exp(denum * c_dt)
        small = np.abs(denum) < 1e-5
        safe = np.where(small, 1.0, denum)
        general = a * (f - 1.0) / safe + b * (f - safe * c_dt - 1.0) / (safe * safe)
        return np.where(small, 0.5 * c_dt * (rhs + rhs_prev * f), general)

    def updateEquation(self, new_fields: np.ndarray, rhs_prev, orig_fields, m: Model) -> np.ndarray:
        """:68-95."""
        n = m.Fields[0].Data.shape[0]
        for i, f in enumerate(m.Fields):  # vec2fields
            f.Data[:] = new_fields[i * n:(i + 1) * n]
        self.fft(m)
        t = self.GetTime()
        freq = as_frequency(self.FT.Freq)
        c_dt = complex(self.Dt, 0.0)
        out = np.zeros(len(m.Fields) * n, dtype=np.float64)
        for i in range(len(m.Fields)):
            rhs = m.GetRHS(i, freq, t)
            denum = m.GetDenum(i, freq, t)
            factor = np.exp(denum * c_dt)
            integral = self.nonlinearIntegral(denum, rhs, rhs_prev[i * n:(i + 1) * n])
            update = orig_fields[i] * factor + integral
            res = np.ascontiguousarray(m.Fields[i].Data - update)
            self.FT.IFFT(res)
            pfutil.div_real_scalar(res, float(n))
            out[i * n:(i + 1) * n] = res.real
        return out

    def Step(self, m: Model):
        """:165-207."""
        n = m.Fields[0].Data.shape[0]
        t = self.GetTime()
        freq = as_frequency(self.FT.Freq)
        self.fft(m)
        orig = [f.Data.copy() for f in m.Fields]
        rhs_prev = np.concatenate([m.GetRHS(i, freq, t) for i in range(len(m.Fields))])
        self.ifft(m)
        explicit = Euler(self.Dt, self.FT, self.Filter)
        explicit.CurrentStep = self.CurrentStep
        explicit.Step(m)
        x0 = np.concatenate([f.Data.real for f in m.Fields])
        if self.NonlinSolver is None:
            self.NonlinSolver = NewtonKrylov()
        x, self.last_converged = self.NonlinSolver.Solve(lambda v: self.updateEquation(v, rhs_prev, orig, m), x0)
        for i, f in enumerate(m.Fields):
            f.Data[:] = x[i * n:(i + 1) * n]
        self.CurrentStep += 1


# ==========================================================================
# pf/rk4.go
# ==========================================================================
class RK4:
    """pf/rk4.go:8-145.  Note Step never advances CurrentStep (rk4.go:130-135:
    only RK4.Propagate does), so through Solver.Propagate t stays 0."""

    def __init__(self, Dt: float, FT, Filter=None):
        self.Dt = Dt
        self.FT = FT
        self.Filter = Filter
        self.CurrentStep = 0

    def Step(self, m: Model):
        m.SyncDerivedFields()
        c_dt = complex(self.Dt, 0.0)
        for f in m.Fields:
            self.FT.FFT(f.Data)
        initial = [f.Copy() for f in m.Fields]
        final = [f.Copy() for f in m.Fields]
        k_factor = [f.Copy() for f in m.Fields]

        self._first_correction(m, k_factor)
        self.PrepareNextCorrection(initial, final, k_factor, m, 1.0 / 6.0)
        self._correction(m, k_factor, 0.5)
        self.PrepareNextCorrection(initial, final, k_factor, m, 1.0 / 3.0)
        self._correction(m, k_factor, 0.5)
        self.PrepareNextCorrection(initial, final, k_factor, m, 1.0 / 3.0)
        self._correction(m, k_factor, 1.0)
        self.PrepareNextCorrection(initial, final, k_factor, m, 1.0 / 6.0)

        t = self.GetTime()
        freq = as_frequency(self.FT.Freq)
        for i in range(len(m.Fields)):
            denum = m.GetDenum(i, freq, t)
            final[i].Data /= (complex(1.0, 0.0) - c_dt * denum)
            m.Fields[i].Data[:] = final[i].Data
            if self.Filter is not None:
                apply_modal_filter(self.Filter, freq, m.Fields[i].Data)
        for f in m.Fields:
            self.FT.IFFT(f.Data)
            pfutil.div_real_scalar(f.Data, float(f.Data.shape[0]))

    def PrepareNextCorrection(self, initial, final, k_factor, m, factor):
        for i in range(len(final)):
            final[i].Data += complex(factor * self.Dt, 0.0) * k_factor[i].Data
            m.Fields[i].Data[:] = initial[i].Data

    def _first_correction(self, m, k_factor):
        for f in m.DerivedFields:
            self.FT.FFT(f.Data)
        t = self.GetTime()
        freq = as_frequency(self.FT.Freq)
        for i in range(len(m.Fields)):
            k_factor[i].Data = m.GetRHS(i, freq, t)

    def _correction(self, m, k_factor, factor):
        t = self.GetTime()
        freq = as_frequency(self.FT.Freq)
        for i, f in enumerate(m.Fields):
            denum = m.GetDenum(i, freq, t)
            f.Data += complex(factor * self.Dt, 0.0) * k_factor[i].Data
            f.Data /= (complex(1.0, 0.0) - complex(factor * self.Dt, 0.0) * denum)
            self.FT.IFFT(f.Data)
            pfutil.div_real_scalar(f.Data, float(f.Data.shape[0]))
        m.SyncDerivedFields()
        for f in m.Fields:
            self.FT.FFT(f.Data)
        for f in m.DerivedFields:
            self.FT.FFT(f.Data)
        for i in range(len(m.Fields)):
            k_factor[i].Data = m.GetRHS(i, freq, t)

    def Propagate(self, nsteps, m):
        for _ in range(nsteps):
            self.Step(m)
            self.CurrentStep += 1

    def SetFilter(self, filt):
        self.Filter = filt

    def GetTime(self):
        return float(self.CurrentStep) * self.Dt


# ==========================================================================
# pf/solver.go
# ==========================================================================
class Solver:
    """pf/solver.go:29-134."""

    def __init__(self, m: Model, domain_size, dt: float, workers: int = 1):
        m.Init()
        self.Model = m
        self.Dt = dt
        self.Callbacks: List[Callable] = []
        self.Monitors: list = []
        self.StartEpoch = 0
        self.FT = pfutil.NewFFTW(domain_size, workers)
        self.Stepper = Euler(dt, self.FT)
        N = pfutil.prod_int(domain_size)
        for f in m.Fields:
            if f.Data.shape[0] != N:
                raise RuntimeError("solver: Inconsistent domain size and number of grid points")

    def AddCallback(self, cb):
        self.Callbacks.append(cb)

    def Propagate(self, nsteps: int):
        for _ in range(nsteps):
            self.Stepper.Step(self.Model)
            t = self.Stepper.GetTime()
            for term in self.Model.ImplicitTerms.values():
                term.OnStepFinished(t, self.Model.Bricks)
            for term in self.Model.ExplicitTerms.values():
                term.OnStepFinished(t, self.Model.Bricks)
            for term in self.Model.MixedTerms.values():
                term.OnStepFinished(t, self.Model.Bricks)

    def SetStepper(self, name: str):
        if name == "euler":
            self.Stepper = Euler(self.Dt, self.FT)
        elif name == "rk4":
            self.Stepper = RK4(self.Dt, self.FT)
        else:
            raise ValueError("Unknown stepper scheme")

    def Solve(self, nepochs: int, nsteps: int):
        for i in range(nepochs):
            self.Propagate(nsteps)
            for cb in self.Callbacks:
                cb(self, i + self.StartEpoch)
            for mon in self.Monitors:
                mon.Add(self.Model.Bricks)


def NewSolver(m: Model, domain_size, dt: float, workers: int = 1) -> Solver:
    return Solver(m, domain_size, dt, workers)
