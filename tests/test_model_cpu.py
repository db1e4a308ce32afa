"""Host-side Model / parser behind the C ABI (no GPU needed): term classification and
derived-field naming must match the oracle's restatement of pf/rhsBuilder.go and the
reference's own parser tests."""
import numpy as np
import pytest

from gopf_b200 import GopfError, pf as gpf
from oracle import pf as opf


def both_models(fields, scalars, equations, n=4):
    g, o = gpf.NewModel(), opf.NewModel()
    for name in fields:
        g.AddField(gpf.NewField(name, n))
        o.AddField(opf.NewField(name, n))
    for name, val in scalars:
        g.AddScalar(gpf.NewScalar(name, val))
        o.AddScalar(opf.NewScalar(name, val))
    for eq in equations:
        g.AddEquation(eq)
        o.AddEquation(eq)
    g.Init()
    o.Init()
    return g, o


CASES = [
    # examples/cahnHilliard/main.go:33
    (["conc"], [("gamma", 2.0), ("m1", -1.0)], ["dconc/dt = LAP conc^3 + m1*LAP conc + m1*gamma*LAP^2 conc"]),
    # pf/model_test.go:45-108
    (["concA", "concB", "concC"], [("kf", 2.0), ("kr", 0.2)],
     ["dconcA/dt = LAP concA - kf*concA^2*concB^3 + kr*concC",
      "dconcB/dt = LAP concB - kf*concA^2*concB^3 + kr*concC",
      "dconcC/dt = LAP concC - kr*concC + kf*concA^2*concB^3"]),
    # pf/model_test.go:17-43
    (["conc"], [], ["dconc/dt = LAP conc"]),
    # pf/euler_test.go:10-85, pf/rk4_test.go:62-100
    (["field"], [("rate", -1.0)], ["dfield/dt = rate*field"]),
    (["field"], [("rate", -1.0)], ["dfield/dt = rate*field^2"]),
    (["field"], [("rate", -1.0)], ["dfield/dt = field + rate*field^2"]),
    # pf/rhsBuilder_test.go:249-273
    (["cluster", "solute"], [], ["dcluster/dt = LAP cluster", "dsolute/dt = solute*cluster"]),
    # pf/rhsBuilder_test.go:94-113
    (["current", "voltage", "magnetic"], [("resistance", 2.0)],
     ["dcurrent/dt = voltage^2", "dvoltage/dt = resistance*current^2", "dmagnetic/dt = current*magnetic^3"]),
    # prefixes
    (["conc"], [("kappa", 0.5)], ["dconc/dt = -LAP^2 conc^3 - kappa*LAP^4 conc + LAP conc^2"]),
]


@pytest.mark.parametrize("fields,scalars,equations", CASES)
def test_classification_matches_oracle(fields, scalars, equations):
    g, o = both_models(fields, scalars, equations)
    assert g.AllFieldNames() == o.AllFieldNames()
    grhs = g.RHS
    assert [(len(r.Terms), len(r.Denum)) for r in grhs] == [(len(r.Terms), len(r.Denum)) for r in o.RHS]
    for f in fields:
        assert g.EqNumber(f) == o.EqNumber(f)


def test_reaction_diffusion_counts():
    # pf/model_test.go:85-107 literal expectations
    g, _ = both_models(*CASES[1])
    assert sorted(g.AllFieldNames()) == sorted(["concA", "concB", "concC", "concA^2*concB^3"])
    assert [(len(r.Terms), len(r.Denum)) for r in g.RHS] == [(2, 1), (2, 1), (1, 2)]


def test_unknown_name_fails_like_the_reference():
    # pf/rhsBuilder_test.go:158-206: unknown scalar / field -> panic
    for eq in ("dconc/dt = m1*conc^2", "dconc/dt = m1*LAP conc^2", "dconc/dt = LAP otherField"):
        m = gpf.NewModel()
        m.AddField(gpf.NewField("conc", 8))
        m.AddEquation(eq)
        with pytest.raises(GopfError, match="is not defined"):
            m.Init()


def test_bad_equations_fail():
    m = gpf.NewModel()
    m.AddField(gpf.NewField("conc", 8))
    with pytest.raises(GopfError, match="equality sign"):
        m.AddEquation("dconc/dt = a = b")
    with pytest.raises(GopfError, match="leibniz"):
        m.AddEquation("conc = LAP conc")
    with pytest.raises(GopfError, match="reserved"):
        m.AddField(gpf.NewField("LAPfield", 8))


def test_user_term_classification():
    # pf/spectralViscosity_test.go:44-58, pf/rhsBuilder_test.go:297-361 (sign prefix keeps the class)
    m = gpf.NewModel()
    m.AddField(gpf.NewField("conc", 16))
    m.RegisterImplicitTerm("SPECTRAL_VISC", gpf.SpectralViscosity(1.0, 0.25, 2), None)
    m.AddEquation("dconc/dt = -SPECTRAL_VISC")
    m.Init()
    assert [(len(r.Terms), len(r.Denum)) for r in m.RHS] == [(0, 1)]

    m = gpf.NewModel()
    m.AddField(gpf.NewField("density", 16))
    term = gpf.IdealMixtureTerm(gpf.IdealMix(1.0, 1.0), "density", 1.0, False)
    m.RegisterMixedTerm("IDEAL_MIX", term, None)
    m.AddEquation("ddensity/dt = IDEAL_MIX")
    # pf/pairCorrelationTerm_test.go:262-272: missing derived field is an error
    with pytest.raises(GopfError, match="Missing derived field"):
        m.Init()
    m.RegisterDerivedField(term.DerivedField(16, m.Bricks))
    m.Init()
    assert [(len(r.Terms), len(r.Denum)) for r in m.RHS] == [(1, 1)]
    assert "ideal_mixture_density_nonlin" in m.DerivedFieldNames


def test_function_expression_errors():
    m = gpf.NewModel()
    m.AddField(gpf.NewField("conc", 8))
    with pytest.raises(GopfError, match="unknown name"):
        m.RegisterFunction("FN", "conc*missing")
    with pytest.raises(GopfError, match="closures"):
        m.RegisterFunction("FN", lambda i, bricks: 0.0)
    m.RegisterFunction("FN", "-(0.1*conc*(1-H(conc)) - 0.1*(1-conc)*H(conc))")
    assert m.DerivedFieldNames == ["FN"]


def test_vandeven_table_matches_oracle():
    from oracle import terms
    for order in (3, 5, 10):
        got = gpf.NewVandeven(order).Data
        exp = terms.NewVandeven(order).Data
        assert np.max(np.abs(got - exp)) < 1e-15


def test_hessian_with_model_term_counts():
    # pf/tensorialHessian_test.go:105-146 through the C ABI parser
    from gopf_b200 import pf as gpf
    N = 16
    m = gpf.NewModel()
    m.AddField(gpf.NewField("conc1", N * N))
    m.AddField(gpf.NewField("conc2", N * N))
    m.RegisterImplicitTerm("HESSIAN", gpf.TensorialHessian([1.0, 2.0, 2.0, 2.0]), None)
    m.AddEquation("dconc1/dt = HESSIAN")
    m.AddEquation("dconc2/dt = -conc2")
    m.Init()
    rhs = m.RHS
    assert len(rhs[0].Terms) == 0 and len(rhs[0].Denum) == 1


def test_registered_field_data_cannot_be_rebound():
    """The C model holds the host pointer of a registered field (pf.Field.Data, pf/model.go:16-22): rebinding the
    Python attribute would leave it dangling, so it is refused; writing into the array is the supported way."""
    import numpy as np
    import pytest
    from gopf_b200 import pf as gpf
    from gopf_b200._lib import GopfError
    f = gpf.NewField("c", 16, np.zeros(16, dtype=np.complex128))
    f.Data = np.ones(16, dtype=np.complex128)  # free to rebind before registration
    m = gpf.NewModel()
    m.AddField(f)
    with pytest.raises(GopfError):
        f.Data = np.zeros(16, dtype=np.complex128)
    f.Data[:] = 2.0
    assert m._host_buffers[0][0] is f.Data
