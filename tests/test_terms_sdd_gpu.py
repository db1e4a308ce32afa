"""GPU parity of the shrinking-dimer stepper (pf.SDD, pf/sdd.go) through the C
ABI against the oracle restatement, plus the reference's own SDD tests (pf/sdd_test.go) restated on
the device path.  Tolerance: 1e-10 on fields, orientation and monitor values."""
import math

import numpy as np
import pytest

from gopf_b200 import pf as gpf
from oracle import pf as opf
from oracle import sdd as osdd
from _sdd_shapes import box_blur_5x5, insert_circle_at_center

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def new_sdd(mod, dims, model):
    return gpf.NewSDD(dims, model) if mod is gpf else osdd.NewSDD(dims, model)


def compare(gs, os_, gfields, ofields, gsdd, osd):
    for gf, of in zip(gfields, ofields):
        assert rel_l2(gf.Data, of.Data) <= TOL
    assert rel_l2(gsdd.orientation, osd.orientation) <= TOL
    assert gsdd.CurrentStep == osd.CurrentStep
    for k in ("MaxForce", "ForcePowerSpectrum", "MaxTorque", "FieldNorm", "FieldNormChange"):
        a, b = getattr(gsdd.Monitor, k), getattr(osd.Monitor, k)
        assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (k, a, b)


def example(mod, dims, eqs, fields_init, scalars=()):
    model = mod.NewModel()
    fields = []
    for name, init in fields_init:
        f = mod.NewField(name, init.shape[0], init.astype(np.complex128).copy())
        model.AddField(f)
        fields.append(f)
    for name, v in scalars:
        model.AddScalar(mod.NewScalar(name, v))
    for eq in eqs:
        model.AddEquation(eq)
    return model, fields


@pytest.mark.parametrize("dims", [[16, 16], [32, 16], [16, 16, 16]], ids=lambda d: "x".join(map(str, d)))
def test_sdd_single_field_vs_oracle(dims):
    # the ExampleModel of pf/sdd_test.go:327-342 on several grids, random orientation
    n = int(np.prod(dims))
    rng = np.random.default_rng(11)
    init = np.where(np.arange(n) > 5, 0.1, 0.0) + 0.01 * rng.standard_normal(n)
    orient = rng.standard_normal(n)
    orient *= 2.0 / np.linalg.norm(orient)  # InitDimerLength = 2
    dt = 0.01
    out = []
    for mod in (gpf, opf):
        model, fields = example(mod, dims, ["dconc/dt = conc^3 - conc + LAP conc"], [("conc", init)])
        solver = mod.NewSolver(model, dims, dt)
        sdd = new_sdd(mod, dims, model)
        sdd.SetInitialOrientation(orient)
        sdd.Dt = dt
        sdd.MinDimerLength = 1e-3
        solver.Stepper = sdd
        solver.Solve(5, 8)
        out.append((solver, fields, sdd))
    (gs, gf, gsdd), (os_, of, osd) = out
    assert not gs.IsFused and gs.KernelLaunches() > 0
    compare(gs, os_, gf, of, gsdd, osd)
    assert abs(gs.Stepper.GetTime() - os_.Stepper.GetTime()) < 1e-12


def test_sdd_two_fields_vs_oracle():
    # two coupled fields: the weighted force of BOTH fields reads the first field's block
    # (pf/sdd.go:204-206 as written), the torque and the projection run over both blocks
    dims = [16, 16]
    n = 256
    rng = np.random.default_rng(5)
    a0 = 0.3 * rng.standard_normal(n)
    b0 = 0.3 * rng.standard_normal(n)
    orient = rng.standard_normal(2 * n)
    orient *= 1.5 / np.linalg.norm(orient)
    dt = 0.02
    eqs = ["dpsi/dt = kap*LAP psi - psi^3 + chi", "dchi/dt = kap*LAP chi - psi^2*chi"]
    out = []
    for mod in (gpf, opf):
        model, fields = example(mod, dims, eqs, [("psi", a0), ("chi", b0)], [("kap", 0.7)])
        solver = mod.NewSolver(model, dims, dt)
        sdd = new_sdd(mod, dims, model)
        sdd.SetInitialOrientation(orient)
        sdd.Dt = dt
        sdd.Alpha = 0.4
        sdd.TimeConstants.Orientation = 2.0
        sdd.TimeConstants.DimerLength = 0.5
        solver.Stepper = sdd
        solver.Solve(3, 7)
        out.append((solver, fields, sdd))
    (gs, gf, gsdd), (os_, of, osd) = out
    compare(gs, os_, gf, of, gsdd, osd)


def test_double_well_saddle():
    # pf/sdd_test.go:68-107 on the device
    N = 4
    out = []
    for mod in (gpf, opf):
        init = mod.NewField("concInit", N * N)
        final = mod.NewField("concFinal", N * N)
        field = mod.NewField("conc", N * N)
        init.Data[:] = -1.0
        final.Data[:] = 1.5
        field.Data[:] = 0.5 * (init.Data + final.Data)
        model = mod.NewModel()
        model.AddField(field)
        model.AddEquation("dconc/dt = conc - conc^3")
        dt = 0.1
        sdd = new_sdd(mod, [N, N], model)
        sdd.Init([init], [final])
        sdd.InitDimerLength = 0.1
        solver = mod.NewSolver(model, [N, N], dt)
        sdd.Dt = dt
        solver.Stepper = sdd
        final_time = sdd.RequiredDimerLengthTime(0.000001 * sdd.DimerLength(0.0))
        solver.Solve(1, int(final_time / dt) + 1)
        out.append(field)
    assert not np.any(np.isnan(out[0].Data))
    assert np.max(np.abs(out[0].Data)) < 1e-6
    assert np.max(np.abs(out[0].Data - out[1].Data)) < 1e-10


def nucleation(mod, N, nsteps):
    # pf/sdd_test.go:163-246; MINUS_CHEM_POT as a device expression / the same closure on the oracle
    gamma, rho = 0.5, 0.05
    init = -np.ones(N * N)
    final = -np.ones(N * N)
    start = -np.ones(N * N)
    insert_circle_at_center(final, N, 15)
    insert_circle_at_center(init, N, 10)
    final, init = box_blur_5x5(final, N), box_blur_5x5(init, N)
    insert_circle_at_center(start, N, 12)
    start = box_blur_5x5(start, N)
    field = mod.NewField("phi", N * N, start.astype(np.complex128))
    model = mod.NewModel()
    model.AddField(field)
    model.AddScalar(mod.NewScalar("gamma", gamma))
    if mod is gpf:
        model.RegisterFunction("MINUS_CHEM_POT", f"(1.0 - phi*phi)*(phi + {3.0 * rho / 4.0!r})")
    else:
        model.RegisterFunction("MINUS_CHEM_POT", lambda i, b: (1.0 - b["phi"].Get(i) ** 2) * (b["phi"].Get(i) + 3.0 * rho / 4.0))
    model.AddEquation("dphi/dt = MINUS_CHEM_POT + gamma*LAP phi")
    sdd = new_sdd(mod, [N, N], model)
    sdd.InitDimerLength = 1.0
    sdd.MinDimerLength = 5e-6
    dt = 0.7
    sdd.Dt = dt
    sdd.Init([mod.NewField("a", N * N, init.astype(np.complex128))], [mod.NewField("b", N * N, final.astype(np.complex128))])
    solver = mod.NewSolver(model, [N, N], dt)
    solver.Stepper = sdd
    solver.Solve(1, nsteps)
    return field, sdd, gamma, rho


def test_classical_nucleation_vs_oracle_and_critical_radius():
    N = 64
    gfield, gsdd, gamma, rho = nucleation(gpf, N, 100)
    ofield, osd, _, _ = nucleation(opf, N, 100)
    assert rel_l2(gfield.Data, ofield.Data) <= TOL
    # the torque is the difference of the forces at two images MinDimerLength = 5e-6 apart, divided
    # by that length (pf/sdd.go:277-283): rounding differences are amplified by ~1/l, so the
    # orientation agrees to ~1e-9 (measured 1.8e-9), not 1e-10, while the field does
    assert rel_l2(gsdd.orientation, osd.orientation) <= 1e-7
    # pf/sdd_test.go:248-271: radius of the critical droplet after 1000 time units
    gfield, gsdd, gamma, rho = nucleation(gpf, N, int(1000.0 / 0.7))
    assert not np.any(np.isnan(gfield.Data))
    rc = 2.0 * (math.sqrt(gamma / 2.0) * 2.0 / 3.0) / rho
    Rc = math.sqrt(float(np.sum(0.5 * (1.0 + gfield.Data.real))) / math.pi)
    assert abs(Rc - rc) < 0.3


def test_revert_orientation_vector():
    # pf/sdd_test.go:344-384
    N = 16
    init = np.where(np.arange(N * N) > 5, 0.1, 0.0)
    orient = np.where(np.arange(N * N) > 5, -1.0, 1.0)
    res = []
    for sign in (1.0, -1.0):
        model, fields = example(gpf, [N, N], ["dconc/dt = conc^3 - conc + LAP conc"], [("conc", init)])
        solver = gpf.NewSolver(model, [N, N], 0.01)
        sdd = gpf.NewSDD([N, N], model)
        sdd.SetInitialOrientation(sign * orient)
        sdd.Dt = 0.01
        solver.Stepper = sdd
        solver.Solve(100, 1)
        res.append(fields[0].Data.copy())
    assert np.max(np.abs(res[0] - res[1])) < 1e-10


def test_sdd_panics():
    # pf/sdd_test.go:386-410, pf/sdd.go:159-161, 431-433
    N = 16
    init = np.where(np.arange(N * N) > 5, 0.1, 0.0)
    model, fields = example(gpf, [N, N], ["dconc/dt = conc^3 - conc + LAP conc"], [("conc", init)])
    solver = gpf.NewSolver(model, [N, N], 0.01)
    sdd = gpf.NewSDD([N, N], model)
    solver.Stepper = sdd
    sdd.Dt = 0.3
    with pytest.raises(gpf.GopfError, match="initialized first"):
        solver.Solve(1, 1)
    sdd.SetInitialOrientation(np.ones(N * N))
    sdd.Dt = 0.0
    with pytest.raises(gpf.GopfError, match="Timestep not set"):
        solver.Solve(10, 1)
    sdd.Dt = 0.3
    solver.Solve(10, 1)
    assert sdd.CurrentStep == 10
    with pytest.raises(gpf.GopfError, match="modal filters"):
        sdd.SetFilter(gpf.NewVandeven(5))
    with pytest.raises(gpf.GopfError, match="Inconsistent length"):
        sdd.SetInitialOrientation(np.ones(3))
