"""Pins oracle/pf.py (parser, Model, Euler, RK4, Solver) against the reference's
own known-answer tests.  Each test cites the Go test it restates.  No GPU needed.
"""
import math

import numpy as np
import pytest

from oracle import pf, pfutil


def c(x):
    return complex(x, 0.0)


# ------------------------------------------------------------------ diffOp
def test_laplacian_n():
    # pf/diffOp_test.go:10-57
    N = 4
    data = np.arange(N, dtype=np.float64).astype(np.complex128)
    freq = lambda i: [float(i)]
    e1 = np.array([-2.0 * math.pi * 2.0 * math.pi * j * j for j in range(N)]) * data
    e2 = np.array([math.pow(2.0 * math.pi, 4.0) * j ** 4 for j in range(N)]) * data
    for power, expect in ((1, e1), (2, e2)):
        tmp = data.copy()
        pf.LaplacianN(power).Eval(freq, tmp)
        assert pfutil.cmplx_equal_approx(expect, tmp, 1e-10)


# ------------------------------------------------------------------ rhsBuilder
def test_name_from_leibniz():
    # pf/rhsBuilder_test.go:10-41
    assert pf.field_name_from_leibniz("dc/dt") == "c"
    assert pf.field_name_from_leibniz("dkappa/dt") == "kappa"
    for bad in ("dc", "ac/dt", "dc/dq"):
        with pytest.raises(ValueError):
            pf.field_name_from_leibniz(bad)


def test_is_bilinear():
    # pf/rhsBuilder_test.go:43-92
    cases = [
        ("c", "2*c", True, ["c"]),
        ("conc", "conc^2", False, ["conc"]),
        ("c", "c*n*r", False, ["c", "n", "r"]),
        ("voltage", "voltage^1.62", False, ["voltage"]),
        ("voltage", "current*voltage^1.0", False, ["voltage", "current"]),
        ("current", "P*current^-2", False, ["current"]),
    ]
    for field, expr, expect, allf in cases:
        assert pf.is_bilinear(expr, field, allf) is expect, (field, expr)


def _electric_model():
    model = pf.NewModel()
    model.AddField(pf.NewField("current", 1, np.array([c(2.0)])))
    model.AddField(pf.NewField("voltage", 1, np.array([c(-4.0)])))
    model.AddField(pf.NewField("magnetic", 1, np.array([c(1.5)])))
    model.AddScalar(pf.NewScalar("resistance", c(2.0)))
    model.AddEquation("dcurrent/dt = voltage^2")
    model.AddEquation("dvoltage/dt = resistance*current^2")
    model.AddEquation("dmagnetic/dt = current*magnetic^3")
    model.SyncDerivedFields()
    return model


def test_concrete_term():
    # pf/rhsBuilder_test.go:94-156
    model = _electric_model()
    freq = lambda i: [1.0, 1.0]
    for expr, sign, expect in [
        ("resistance*current^2", "+", 8.0),
        ("voltage^2", "+", 16.0),
        ("resistance*current^2", "-", -8.0),
        ("voltage^2", "-", -16.0),
    ]:
        term = pf.concrete_term(pf.SubStringDelimiter(expr, sign), model)
        got = np.zeros(1, dtype=np.complex128)
        term(freq, 0.0, got)
        assert pfutil.cmplx_equal_approx(got, np.array([c(expect)]), 1e-10), expr


def test_panic_on_unknown_name():
    # pf/rhsBuilder_test.go:158-206
    model = pf.NewModel()
    model.AddField(pf.NewField("conc", 8))
    for expr, should in [("LAP conc", False), ("conc", False), ("m1*conc", True),
                         ("m1*LAP conc", True), ("LAP otherField", True)]:
        if should:
            with pytest.raises(ValueError):
                pf.concrete_term(pf.SubStringDelimiter(expr), model)
        else:
            pf.concrete_term(pf.SubStringDelimiter(expr), model)


def test_lap_user_defined():
    # pf/rhsBuilder_test.go:208-247
    N = 16
    model = pf.NewModel()
    field = pf.NewField("conc", N * N)
    field.Data[:] = 0.1 * np.arange(N * N)
    model.AddField(field)
    model.RegisterFunction("myfunc", lambda i, bricks: bricks["conc"].Get(i))
    model.AddEquation("dconc/dt = LAP myfunc")
    model.Init()
    terms = model.RHS[0].Terms
    assert len(terms) == 1
    ft = pfutil.NewFFTW([N, N])
    out = np.zeros(N * N, dtype=np.complex128)
    terms[0](ft.Freq, 0.0, out)
    for i in range(N * N):
        f = ft.Freq(i)
        expect = -4.0 * math.pi * math.pi * (f[0] * f[0] + f[1] * f[1]) * field.Data[i].real
        assert abs(expect - out[i].real) < 1e-10


def test_non_alphabetic_order():
    # pf/rhsBuilder_test.go:249-273
    model = pf.NewModel()
    cluster = pf.NewField("cluster", 8)
    solute = pf.NewField("solute", 8)
    cluster.Data[:] = 1.0
    solute.Data[:] = 2.0
    expect = cluster.Data * solute.Data
    model.AddField(cluster)
    model.AddField(solute)
    model.AddEquation("dcluster/dt = LAP cluster")
    model.AddEquation("dsolute/dt = solute*cluster")
    model.Init()
    rhs = model.GetRHS(1, lambda i: [1.0], 0.0)
    assert pfutil.cmplx_equal_approx(expect, rhs, 1e-10)


class SingleTerm:
    @staticmethod
    def _ones(freq, t, field):
        field[:] = 1.0

    def Construct(self, bricks):
        return self._ones

    ConstructLinear = Construct
    ConstructNonLinear = Construct

    def OnStepFinished(self, t, bricks):
        pass


def test_negative_sign_before_user_defined():
    # pf/rhsBuilder_test.go:297-361
    field = pf.NewField("conc", 8)
    expect = np.full(8, c(-1.0))
    freq = lambda i: [0.0, 0.0]
    ev = np.zeros(8, dtype=np.complex128)

    model = pf.NewModel()
    model.AddField(field)
    model.RegisterExplicitTerm("TERM", SingleTerm(), None)
    rhs = pf.Build("dconc/dt=-TERM", model)
    assert (len(rhs.Terms), len(rhs.Denum)) == (1, 0)
    rhs.Terms[0](freq, 0.0, ev)
    assert pfutil.cmplx_equal_approx(ev, expect, 1e-10)

    model = pf.NewModel()
    model.AddField(field)
    model.RegisterImplicitTerm("TERM", SingleTerm(), None)
    rhs = pf.Build("dconc/dt=-TERM", model)
    assert (len(rhs.Terms), len(rhs.Denum)) == (0, 1)
    rhs.Denum[0](freq, 0.0, ev)
    assert pfutil.cmplx_equal_approx(ev, expect, 1e-10)

    model = pf.NewModel()
    model.AddField(field)
    model.RegisterMixedTerm("TERM", SingleTerm(), None)
    rhs = pf.Build("dconc/dt=-TERM", model)
    assert (len(rhs.Terms), len(rhs.Denum)) == (1, 1)
    rhs.Denum[0](freq, 0.0, ev)
    assert pfutil.cmplx_equal_approx(ev, expect, 1e-10)
    rhs.Terms[0](freq, 0.0, ev)
    assert pfutil.cmplx_equal_approx(ev, expect, 1e-10)


PREFIX_CASES = [
    ("mystring", []),
    ("*mystring", ["*"]),
    ("LAPmystring", ["LAP"]),
    ("LAP*mystring", ["LAP", "*"]),
    ("*LAP*mystring", ["*", "LAP", "*"]),
    ("*LAPLAPmystring", ["*", "LAP", "LAP"]),
    ("*LAP^2mystring", ["*", "LAP^2"]),
]


def test_remove_and_get_known_prefixes():
    # pf/rhsBuilder_test.go:363-440
    for s, expect in PREFIX_CASES:
        assert pf.remove_known_prefixes(s) == "mystring"
        assert pf.get_known_prefixes(s) == expect


def test_constructor_with_prefix_handling():
    # pf/rhsBuilder_test.go:442-510
    def myfunc(freq, t, data):
        data[0] = c(1.0)

    freq = lambda i: [1.0, 1.0]
    two_pi = 2 * math.pi
    cases = [
        (["-"], -1.0),
        (["-", "-"], 1.0),
        (["-", "-", "-"], -1.0),
        (["LAP"], -2.0 * two_pi ** 2),
        (["LAP", "-"], 2.0 * two_pi ** 2),
        (["-", "LAP"], 2.0 * two_pi ** 2),
        (["LAP^2"], 4.0 * two_pi ** 4),
        (["LAP^4"], 16.0 * two_pi ** 8),
        (["LAP^4", "-", "-"], 16.0 * two_pi ** 8),
        (["-", "LAP^4", "-"], 16.0 * two_pi ** 8),
    ]
    for prefixes, expect in cases:
        data = np.zeros(2, dtype=np.complex128)
        pf.construct_func(myfunc, list(prefixes))(freq, 0.0, data)
        assert abs(data[0].real - expect) < 1e-6 and abs(data[0].imag) < 1e-6, prefixes
        assert data[1] == 0


# ------------------------------------------------------------------ util
def test_get_non_linear_field_exp():
    # pf/util_test.go:11-45
    names = ["conc1", "conc2", "eta1", "eta2"]
    for expr, field, expect in [
        ("conc1^2*eta1*factor", "conc1", "conc1^2*eta1"),
        ("conc1^2*eta1*factor", "eta1", "conc1^2*eta1"),
        ("conc2*conc1", "conc2", "conc1*conc2"),
        ("LAPconc2^2*eta2^3", "conc2", "conc2^2*eta2^3"),
    ]:
        assert pf.get_non_linear_field_expressions(expr, field, names) == expect


def test_derived_calc_from_desc():
    # pf/util_test.go:47-75
    fields = [
        pf.NewField("conc1", 2, np.array([c(1.0), c(2.0)])),
        pf.NewField("conc2", 2, np.array([c(3.0), c(4.0)])),
        pf.NewField("conc3", 2, np.array([c(5.0), c(6.0)])),
    ]
    for desc, expect in [("conc1^2*conc2", [3.0, 16.0]), ("conc3^2*conc2", [75.0, 144.0])]:
        arr = np.zeros(2, dtype=np.complex128)
        pf.derived_field_calc_from_desc(desc, fields)(arr)
        assert pfutil.cmplx_equal_approx(np.array(expect, dtype=np.complex128), arr, 1e-10)


def test_get_power():
    # pf/util_test.go:77-100
    assert pf.get_power("conc1^2") == 2.0
    assert pf.get_power("conc1") == 1.0
    assert pf.get_power("conc4^-4.5") == -4.5


def test_get_field_name():
    # pf/util_test.go:102-124
    names = ["conc1", "conc2", "conc3", "conc1^2*conc2", "conc3^3", "conc2^4*conc1^2", "conc1^2"]
    assert pf.get_field_name("conc1^2*conc2*otherstuff", names) == "conc1^2*conc2"
    assert pf.get_field_name("*randomstuff*conc1^2*otherstuff", names) == "conc1^2"


def test_apply_modal_filter():
    # pf/util_test.go:126-144
    class Dummy:
        def Eval(self, x):
            return 0.5

    data = np.arange(10, dtype=np.float64).astype(np.complex128)
    pf.apply_modal_filter(Dummy(), lambda i: [0.0], data)
    assert np.allclose(data.real, 0.5 * np.arange(10), atol=1e-10) and np.all(data.imag == 0)


def test_split_on_many():
    # pf/util_test.go:146-186
    for value, delims, expect in [
        ("a+b", ["+"], ["a", "b"]),
        ("a+b+cd", ["+"], ["a", "b", "cd"]),
        ("a+b-cd", ["+", "-"], ["a", "b", "cd"]),
        ("cdb-the+two", ["+", "-", "7"], ["cdb", "the", "two"]),
    ]:
        assert sorted(s.SubString for s in pf.split_on_many(value, delims)) == expect
    # delimiters are carried with the piece that follows them
    got = {s.SubString: s.PreceedingDelimiter for s in pf.split_on_many("a+b-cd", ["+", "-"])}
    assert got == {"a": "", "b": "+", "cd": "-"}


def test_sort_factors():
    # pf/util_test.go:295-322
    for expr, expect in [
        ("solute*conc*temperature", "conc*solute*temperature"),
        ("solute", "solute"),
        ("current^2*voltage", "current^2*voltage"),
        ("voltage*current^2", "current^2*voltage"),
    ]:
        assert pf.sort_factors(expr) == expect


def test_go_find_all_drops_abutting_empty_matches():
    assert pf.go_find_all(r"[^\*]*", "a*b") == ["a", "b"]
    assert pf.go_find_all(r"[^\*]*", "a**b") == ["a", "", "b"]
    assert pf.go_find_all(r"[^\*]*", "") == [""]


# ------------------------------------------------------------------ model
def test_term_diffusion():
    # pf/model_test.go:17-43
    m = pf.NewModel()
    conc = pf.NewField("conc", 2, np.array([c(1.0), c(2.0)]))
    m.AddField(conc)
    m.AddEquation("dconc/dt = LAP conc")
    m.Init()
    assert len(m.RHS[0].Terms) == 0 and len(m.RHS[0].Denum) == 1
    values = np.zeros(2, dtype=np.complex128)
    m.RHS[0].Denum[0](lambda i: [float(i), float(i)], 0.0, values)
    two_pi_sq = (2.0 * math.pi) ** 2
    assert pfutil.cmplx_equal_approx(np.array([0.0, -2.0 * two_pi_sq], dtype=np.complex128), values, 1e-10)


def test_reaction_diffusion():
    # pf/model_test.go:45-108
    m = pf.NewModel()
    m.AddField(pf.NewField("concA", 2, np.array([c(1.0), c(2.0)])))
    m.AddField(pf.NewField("concB", 2, np.array([c(3.0), c(5.0)])))
    m.AddField(pf.NewField("concC", 2, np.array([c(-1.0), c(1.0)])))
    m.AddScalar(pf.NewScalar("kf", c(2.0)))
    m.AddScalar(pf.NewScalar("kr", c(0.2)))
    m.AddEquation("dconcA/dt = LAP concA - kf*concA^2*concB^3 + kr*concC")
    m.AddEquation("dconcB/dt = LAP concB - kf*concA^2*concB^3 + kr*concC")
    m.AddEquation("dconcC/dt = LAP concC - kr*concC + kf*concA^2*concB^3")
    m.Init()
    assert sorted(m.AllFieldNames()) == sorted(["concA", "concB", "concC", "concA^2*concB^3"])
    assert len(m.RHS) == 3
    assert [(len(r.Terms), len(r.Denum)) for r in m.RHS] == [(2, 1), (2, 1), (1, 2)]


def test_user_defined_terms():
    # pf/model_test.go:110-193
    N = 64

    class LapDensitySquared:
        n_construct = 0

        def Construct(self, bricks):
            self.n_construct += 1

            def fn(freq, t, field):
                field[:] = bricks["density^2"].Get(np.arange(field.shape[0]))
                pf.LaplacianN(1).Eval(freq, field)

            return fn

        def OnStepFinished(self, t, bricks):
            pass

    model = pf.NewModel()
    field = pf.NewField("density", N * N)
    model.AddField(field)
    term = LapDensitySquared()
    d = pf.DerivedField(np.zeros(N * N, dtype=np.complex128), "density^2",
                        lambda out: out.__setitem__(slice(None), pfutil.go_cpow(field.Data, 2)))
    model.RegisterExplicitTerm("LP_DENSITY_SQUARED", term, [d])
    model.AddEquation("ddensity/dt = LP_DENSITY_SQUARED")
    model.Init()
    assert len(model.Fields) == 1 and model.Fields[0].Name == "density"
    assert len(model.DerivedFields) == 1 and model.DerivedFields[0].Name == "density^2"
    assert len(model.ExplicitTerms) == 1 and term.n_construct == 1
    assert len(model.RHS[0].Terms) == 1 and len(model.RHS[0].Denum) == 0


def test_function():
    # pf/model_test.go:195-219
    model = pf.NewModel()
    f = pf.NewField("myfield", 8)
    f.Data[:] = np.arange(8)
    model.AddField(f)
    model.RegisterFunction("myfunc", lambda i, bricks: bricks["myfield"].Get(i))
    model.SyncDerivedFields()
    assert np.allclose(model.DerivedFields[0].Data.real, np.arange(8), atol=1e-10)


def test_equation_number():
    # pf/model_test.go:268-307
    for eqns, field, expect in [
        (["dconc/dt = 0"], "conc", 0),
        (["dconcA/dt = 0", "dconcB/dt = 0"], "concB", 1),
        (["dconcA/dt = 0", "dconcB/dt = 0"], "concA", 0),
        (["dtemp/dt = 0", "dconc/dt = 0", "dvoltage/dt = 0"], "voltage", 2),
    ]:
        model = pf.NewModel()
        model.Equations = eqns
        assert model.EqNumber(field) == expect


def test_modifier():
    # pf/model_test.go:309-345
    field = pf.NewField("conc", 8)
    field2 = pf.NewField("conc2", 8)
    field2.Data[:] = 1.0
    model = pf.NewModel()
    model.AddField(field)
    model.AddField(field2)

    def mod(data):
        data *= 2.0

    model.RegisterRHSModifier(1, mod)
    model.AddEquation("dconc/dt = conc2")
    model.AddEquation("dconc2/dt = -conc2")
    model.Init()
    freq = lambda i: [3.0, 3.0]
    assert pfutil.cmplx_equal_approx(model.GetRHS(0, freq, 0.0), np.full(8, c(1.0)), 1e-10)
    assert pfutil.cmplx_equal_approx(model.GetRHS(1, freq, 0.0), np.full(8, c(-2.0)), 1e-10)


def test_cahn_hilliard_parse_worked_example():
    # SURVEY 3.2 worked parse of examples/cahnHilliard/main.go:33
    m = pf.NewModel()
    m.AddScalar(pf.NewScalar("gamma", c(2.0)))
    m.AddScalar(pf.NewScalar("m1", c(-1.0)))
    m.AddField(pf.NewField("conc", 16))
    m.AddEquation("dconc/dt = LAP conc^3 + m1*LAP conc + m1*gamma*LAP^2 conc")
    m.Init()
    assert [d.Name for d in m.DerivedFields] == ["conc^3"]
    assert (len(m.RHS[0].Terms), len(m.RHS[0].Denum)) == (1, 2)
    freq = lambda i: [0.1 * i, 0.0]
    den = m.GetDenum(0, freq, 0.0)
    k2 = (2 * math.pi * 0.1 * np.arange(16)) ** 2
    assert np.allclose(den.real, k2 - 2.0 * k2 * k2, rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------ steppers
def _decay_model(eq, N=8, c0=1.0):
    field = pf.NewField("field", N * N)
    field.Data[:] = c0
    model = pf.NewModel()
    model.AddField(field)
    model.AddScalar(pf.Scalar("rate", c(-1.0)))
    model.AddEquation(eq)
    model.Init()
    return model, field


def test_euler_exponential_decay():
    # pf/euler_test.go:10-49
    model, field = _decay_model("dfield/dt = rate*field")
    st = pf.Euler(0.001, pfutil.NewFFTW([8, 8]))
    st.Propagate(1000, model)
    assert np.all(np.abs(field.Data.real - math.exp(-1.0)) < 1e-3) and np.all(np.abs(field.Data.imag) < 1e-3)
    assert abs(st.GetTime() - 1.0) < 1e-10


def test_euler_square_decay():
    # pf/euler_test.go:51-85
    model, field = _decay_model("dfield/dt = rate*field^2")
    st = pf.Euler(0.001, pfutil.NewFFTW([8, 8]))
    st.Propagate(1000, model)
    assert np.all(np.abs(field.Data.real - 0.5) < 1e-3) and np.all(np.abs(field.Data.imag) < 1e-3)


def test_rk4_simple_model():
    # pf/rk4_test.go:14-55
    model, field = _decay_model("dfield/dt = rate*field^2")
    st = pf.RK4(0.1, pfutil.NewFFTW([8, 8]))
    st.Propagate(10, model)
    assert np.all(np.abs(field.Data.real - 0.5) < 1e-6) and np.all(np.abs(field.Data.imag) < 1e-6)
    assert abs(st.GetTime() - 1.0) < 1e-10


def test_rk4_with_implicit():
    # pf/rk4_test.go:62-100
    c0 = 0.5
    model, field = _decay_model("dfield/dt = field + rate*field^2", c0=c0)
    st = pf.RK4(0.01, pfutil.NewFFTW([8, 8]))
    st.Propagate(100, model)
    A = 1.0 / c0 - 1.0
    expect = math.exp(1.0) / (A + math.exp(1.0))
    assert np.all(np.abs(field.Data.real - expect) < 1e-3) and np.all(np.abs(field.Data.imag) < 1e-3)


def test_solver_diffusion():
    # pf/solver_test.go:9-34
    m = pf.NewModel()
    conc = pf.NewField("conc", 16 * 16)
    conc.Data[128] = 1.0
    m.AddField(conc)
    m.AddEquation("dconc/dt = LAP conc")
    solver = pf.NewSolver(m, [16, 16], 0.1)
    solver.Solve(10, 10)
    assert abs(conc.Data.real.sum() - 1.0) < 1e-4
    assert np.all(conc.Data.real < 1.0) and np.all(conc.Data.real >= 0.0)


def test_gauss_seidel_field_ordering():
    # SURVEY 3.1 quirk (pf/euler.go:27-39): equation i+1 sees field i's UPDATED spectrum.
    N = 8
    a = pf.NewField("aa", N * N)
    b = pf.NewField("bb", N * N)
    a.Data[:] = 1.0
    b.Data[:] = 0.0
    m = pf.NewModel()
    m.AddField(a)
    m.AddField(b)
    m.AddScalar(pf.NewScalar("rate", c(-1.0)))
    m.AddScalar(pf.NewScalar("one", c(1.0)))
    m.AddEquation("daa/dt = rate*aa")
    m.AddEquation("dbb/dt = one*aa")
    m.Init()
    dt = 0.5
    pf.Euler(dt, pfutil.NewFFTW([N, N])).Step(m)
    a_new = 1.0 / (1.0 + dt)
    assert np.allclose(a.Data.real, a_new, atol=1e-13)
    assert np.allclose(b.Data.real, dt * a_new, atol=1e-13)  # not dt * 1.0
