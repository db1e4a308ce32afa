"""Randomised cross-check of the C++ equation parser (gopf_b200/csrc/parser.cu, model.cu) against the
oracle's restatement of pf/rhsBuilder.go + pf/util.go: for seeded random equations built from the
reference's grammar (fields, powers, scalar prefactors, LAP / LAP^n operators, signs) both must agree
on the derived-field names and on how many terms go to the explicit side and to the denominator.
No GPU needed."""
import random

import pytest

from gopf_b200 import pf as gpf
from oracle import pf as opf

FIELDS = ["conc", "eta", "phi"]
SCALARS = [("m1", -1.0), ("gamma", 2.0), ("kappa", 0.25)]


def random_term(rng, own_field):
    parts = []
    for _ in range(rng.choice([0, 0, 1, 2])):
        s = rng.choice(SCALARS)[0]
        parts.append(s if rng.random() < 0.7 else f"{s}^{rng.choice([2, 3])}")
    lap = rng.choice(["", "", "LAP ", "LAP^2 ", "LAP^4 "])
    nf = rng.choice([1, 1, 1, 2])
    fields = []
    for k in range(nf):
        f = own_field if (k == 0 and rng.random() < 0.5) else rng.choice(FIELDS)
        p = rng.choice([1, 1, 2, 3])
        fields.append(f if p == 1 else f"{f}^{p}")
    body = "*".join(parts + [lap + fields[0]] + fields[1:])
    return body


def random_equation(rng, field):
    n = rng.choice([1, 2, 3, 4])
    eq = f"d{field}/dt = "
    for k in range(n):
        sign = rng.choice(["+", "-"])
        if k == 0:
            eq += "-" if (sign == "-" and rng.random() < 0.5) else ""
        else:
            eq += f" {sign} "
        eq += random_term(rng, field)
    return eq


@pytest.mark.parametrize("seed", range(60))
def test_random_equations_classify_like_the_oracle(seed):
    rng = random.Random(seed)
    nfields = rng.choice([1, 2, 3])
    fields = FIELDS[:nfields]
    eqs = [random_equation(rng, f) for f in fields]
    g, o = gpf.NewModel(), opf.NewModel()
    for name in FIELDS[:nfields]:
        g.AddField(gpf.NewField(name, 4))
        o.AddField(opf.NewField(name, 4))
    for name, val in SCALARS:
        g.AddScalar(gpf.NewScalar(name, val))
        o.AddScalar(opf.NewScalar(name, val))
    # terms may name fields that are not in this model: both sides must then fail, or both succeed
    errs = []
    for mod, m in ((gpf, g), (opf, o)):
        try:
            for eq in eqs:
                m.AddEquation(eq)
            m.Init()
            errs.append(None)
        except Exception as e:  # noqa: BLE001 -- the reference panics; either side raises its own type
            errs.append(e)
    assert (errs[0] is None) == (errs[1] is None), (eqs, errs)
    if errs[0] is not None:
        return
    assert g.AllFieldNames() == o.AllFieldNames(), eqs
    assert [(len(r.Terms), len(r.Denum)) for r in g.RHS] == [(len(r.Terms), len(r.Denum)) for r in o.RHS], eqs
