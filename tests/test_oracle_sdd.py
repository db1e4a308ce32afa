"""Oracle restatement of the Shrinking-Dimer-Dynamics stepper (oracle/sdd.py <- pf/sdd.go) pinned
to every test of pf/sdd_test.go, including its numpy-generated golden vector for the Householder
reflection.  No GPU needed."""
import math

import numpy as np
import pytest

from oracle import pf, pfutil
from oracle.sdd import NewSDD, SDD, diagonalShermannMorrison

from _sdd_shapes import box_blur_5x5, insert_circle_at_center


def test_fourier_outer_product():
    # pf/sdd_test.go:14-57
    N = 8
    i = np.arange(N * N, dtype=np.float64)
    data = (i / 10.0).astype(np.complex128)
    vector = ((i * i - i) / 10.0).astype(np.complex128)
    result = vector.real * float(np.sum(data.real * vector.real))
    ft = pfutil.NewFFTW([N, N])
    ft.FFT(data)
    ft.FFT(vector)
    dot = np.sum(np.conj(vector) * data) / float(N * N)
    vector *= dot
    ft.IFFT(vector)
    vector /= float(N * N)
    assert np.max(np.abs(vector.real - result)) < 1e-6 and np.max(np.abs(vector.imag)) < 1e-6


def test_dimer_length_time():
    # pf/sdd_test.go:59-66.  The Go test builds SDD{InitDimerLength: 5.0}: TimeConstants is the zero
    # value, t = 0 * log(10) = 0 and exp(-0/0) is NaN, so its comparison passes vacuously.  The
    # meaningful statement (NewSDD's time constant 1.0) is checked here, and the NaN as well.
    from oracle.sdd import SDDTimeConstants
    dimer = SDD.__new__(SDD)
    dimer.InitDimerLength, dimer.MinDimerLength = 5.0, 0.0
    dimer.TimeConstants = SDDTimeConstants(1.0, 1.0)
    t = dimer.RequiredDimerLengthTime(0.5)
    assert abs(dimer.DimerLength(t) - 0.5) < 1e-10
    dimer.MinDimerLength = 0.7  # :366-368
    assert dimer.DimerLength(t) == 0.7


def test_double_well():
    # pf/sdd_test.go:68-107: saddle of the double well at 0
    N = 4
    init = pf.NewField("concInit", N * N)
    final = pf.NewField("concFinal", N * N)
    field = pf.NewField("conc", N * N)
    init.Data[:] = -1.0
    final.Data[:] = 1.5
    field.Data[:] = 0.5 * (init.Data + final.Data)
    model = pf.NewModel()
    model.AddField(field)
    model.AddEquation("dconc/dt = conc - conc^3")
    dt = 0.1
    sdd = NewSDD([N, N], model)
    sdd.Init([init], [final])
    sdd.InitDimerLength = 0.1
    solver = pf.NewSolver(model, [N, N], dt)
    sdd.Dt = dt
    solver.Stepper = sdd
    final_time = sdd.RequiredDimerLengthTime(0.000001 * sdd.DimerLength(0.0))
    solver.Solve(1, int(final_time / dt) + 1)
    assert not np.any(np.isnan(field.Data))
    assert np.max(np.abs(field.Data)) < 1e-6


def test_2d_surface():
    # pf/sdd_test.go:109-149: E(x, y) = (x^2 - 1)^2 + y^2 has a saddle at (0, 0); two one-node fields
    x = pf.NewField("xCrd", 1)
    y = pf.NewField("yCrd", 1)
    model = pf.NewModel()
    model.AddField(x)
    model.AddField(y)
    model.AddScalar(pf.NewScalar("FOUR", 4.0))
    model.AddScalar(pf.NewScalar("TWO", 2.0))
    model.AddEquation("dxCrd/dt = FOUR*xCrd - FOUR*xCrd^3")
    model.AddEquation("dyCrd/dt = -TWO*yCrd")
    dt = 0.001
    solver = pf.NewSolver(model, [1, 1], dt)
    stepper = NewSDD([1, 1], model)
    stepper.Dt = dt
    stepper.SetInitialOrientation([0.1, 0.5])
    x.Data[0] = 0.2
    y.Data[0] = 0.3
    solver.Stepper = stepper
    solver.Solve(1, int(stepper.RequiredDimerLengthTime(1e-5) / dt))
    assert abs(x.Data[0].real) < 1e-3 and abs(y.Data[0].real) < 1e-3
    assert abs(stepper.orientation[0] - 1.0) < 1e-3 and abs(stepper.orientation[1]) < 1e-3


def classical_nucleation_setup(mod, sdd_factory, N=64):
    """pf/sdd_test.go:163-246 up to solver.Solve; `mod` is oracle.pf or the device mirror."""
    gamma, rho = 0.5, 0.05
    init = -np.ones(N * N)
    final = -np.ones(N * N)
    start = -np.ones(N * N)
    insert_circle_at_center(final, N, 15)
    insert_circle_at_center(init, N, 10)
    final = box_blur_5x5(final, N)
    init = box_blur_5x5(init, N)
    insert_circle_at_center(start, N, 12)
    start = box_blur_5x5(start, N)
    field = mod.NewField("phi", N * N, start.astype(np.complex128))
    model = mod.NewModel()
    model.AddField(field)
    model.AddScalar(mod.NewScalar("gamma", gamma))
    return model, field, init, final, gamma, rho


def test_classical_nucleation():
    # pf/sdd_test.go:163-271: radius of the critical droplet within 0.3 of 2 sigma / rho
    N = 64
    model, field, init, final, gamma, rho = classical_nucleation_setup(pf, NewSDD, N)
    model.RegisterFunction("MINUS_CHEM_POT", lambda i, b: (1.0 - b["phi"].Get(i) ** 2) * (b["phi"].Get(i) + 3.0 * rho / 4.0))
    model.AddEquation("dphi/dt = MINUS_CHEM_POT + gamma*LAP phi")
    sdd = NewSDD([N, N], model)
    sdd.InitDimerLength = 1.0
    sdd.MinDimerLength = 5e-6
    dt = 0.7
    sdd.Dt = dt
    sdd.Init([pf.NewField("a", N * N, init.astype(np.complex128))], [pf.NewField("b", N * N, final.astype(np.complex128))])
    solver = pf.NewSolver(model, [N, N], dt)
    solver.Stepper = sdd
    solver.Solve(int(1000.0 / dt), 1)
    data = field.Data
    assert not np.any(np.isnan(data))
    rc = 2.0 * (math.sqrt(gamma / 2.0) * 2.0 / 3.0) / rho
    Rc = math.sqrt(float(np.sum(0.5 * (1.0 + data.real))) / math.pi)
    assert abs(Rc - rc) < 0.3
    assert sdd.Monitor.MaxTorque >= 0.0 and sdd.Monitor.FieldNorm > 0.0


def test_diagonal_shermann_morrison():
    # pf/sdd_test.go:273-325
    diag = np.array([0.4, -0.2, 1.4])
    u = np.array([-1.0, 2.0, 3.0])
    v = np.array([2.0, 2.3, 1.2])
    b = np.array([4.0, 6.0, 8.2])
    A = np.diag(diag) + np.outer(u, v)
    res = np.linalg.solve(A, b)
    dsm = diagonalShermannMorrison((1.0 / diag).astype(np.complex128), u.astype(np.complex128), v.astype(np.complex128))
    bc = b.astype(np.complex128)
    dsm.dot(bc)
    assert np.max(np.abs(bc.real - res)) < 1e-10


def example_model():
    # pf/sdd_test.go:327-342
    N = 16
    f1 = pf.NewField("conc", N * N)
    f1.Data[6:] = 0.1
    model = pf.NewModel()
    model.AddField(f1)
    model.AddEquation("dconc/dt = conc^3 - conc + LAP conc")
    return model, pf.NewSolver(model, [N, N], 0.01)


def test_revert_orientation_vector():
    # pf/sdd_test.go:344-384: the trajectory is invariant under v -> -v
    model, solver = example_model()
    N = 16
    orient = np.where(np.arange(N * N) > 5, -1.0, 1.0)
    stepper = NewSDD([N, N], model)
    stepper.SetInitialOrientation(orient)
    stepper.Dt = 0.01
    solver.Stepper = stepper
    orig = model.Fields[0].Data.copy()
    solver.Solve(100, 1)
    first = model.Fields[0].Data.copy()
    model.Fields[0].Data[:] = orig
    stepper = NewSDD([N, N], model)
    stepper.SetInitialOrientation(-orient)
    stepper.Dt = 0.01
    solver.Stepper = stepper
    solver.Solve(100, 1)
    assert np.max(np.abs(model.Fields[0].Data - first)) < 1e-10


def test_panic_on_zero_time_step():
    # pf/sdd_test.go:386-410
    model, solver = example_model()
    stepper = NewSDD([16, 16], model)
    stepper.SetInitialOrientation(np.ones(256))
    stepper.Dt = 0.0
    solver.Stepper = stepper
    with pytest.raises(RuntimeError, match="Timestep not set"):
        solver.Solve(10, 1)
    stepper.Dt = 0.3
    solver.Solve(10, 1)


def test_uninitialised_and_filter_panics():
    # pf/sdd.go:159-161, 431-433
    model, solver = example_model()
    stepper = NewSDD([16, 16], model)
    stepper.Dt = 0.1
    with pytest.raises(RuntimeError, match="initialized first"):
        stepper.Step(model)
    with pytest.raises(RuntimeError, match="modal filters"):
        stepper.SetFilter(None)
    with pytest.raises(RuntimeError, match="Inconsistent length"):
        stepper.SetInitialOrientation(np.ones(3))


def test_householder_golden():
    # pf/sdd_test.go:412-455: expected values generated with numpy by the reference's author
    N = 4
    model = pf.NewModel()
    model.AddField(pf.NewField("field", N * N))
    sdd = NewSDD([N, N], model)
    i = np.arange(N * N, dtype=np.float64)
    vec = i.astype(np.complex128)
    orient = (i * i - 3.0 * i)
    orient = (orient / math.sqrt(float(np.sum(orient * orient)))).astype(np.complex128)
    sdd.ft.FFT(vec)
    sdd.ft.FFT(orient)
    sdd.householder(vec, orient, 2.0, vec.shape[0])
    sdd.ft.IFFT(vec)
    vec /= float(N * N)
    expect = [0.0, 1.41446756, 2.41446756, 3.0, 3.17106489, 2.92766222, 2.26979199, 1.19745421,
              -0.28935113, -2.19062403, -4.50636448, -7.23657249, -10.38124806, -13.94039118, -17.91400186, -22.3020801]
    assert np.max(np.abs(vec.real - np.array(expect))) < 1e-6


def test_householder_denum():
    # pf/sdd_test.go:457-517: x_{n+1} = x_n + dt (I - sigma v v^T) x_{n+1} against a dense solve
    sigma, dt = 1.0, 0.8
    model, _ = example_model()
    N = 16
    sdd = NewSDD([N, N], model)
    sdd.Dt = dt
    i = np.arange(N * N, dtype=np.float64)
    orient = 0.1 * (i * i - 4.0 * i)
    x = 0.1 * i
    orient = orient / math.sqrt(float(np.sum(orient * orient)))
    A = (1.0 - dt) * np.eye(N * N) + dt * sigma * np.outer(orient, orient)
    res = np.linalg.solve(A, x)
    c_orient = orient.astype(np.complex128)
    c_x = x.astype(np.complex128)
    sdd.ft.FFT(c_orient)
    sdd.ft.FFT(c_x)
    dsm = sdd.householderDenum(np.ones(N * N, dtype=np.complex128), c_orient, sigma)
    dsm.dot(c_x)
    sdd.ft.IFFT(c_x)
    c_x /= float(N * N)
    assert np.max(np.abs(c_x.real - res)) < 1e-10
