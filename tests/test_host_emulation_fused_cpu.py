"""The fused single-field step -- k_fused_real and k_fused_kspace of gopf_b200/csrc/step_kernels.cuh, the
kernels behind the headline 3-D Cahn-Hilliard number -- run on the HOST and checked against the
oracle: tests/host_emul/emul_fused.cpp compiles the kernels' own source with g++, executes every CUDA thread of
a block as an OS thread and launches the kernels in the order Solver::euler_step_fused does, with the program
the C++ model compiles (gopf_model_fused_program_image).  Covered: the real-polynomial fast form with the
single-tile (LATE) and two-tile k-space kernels, tile widths 2 and 4, 2-D and 3-D, a general program (pfc: pair
correlation + ideal mixture through the rolled interpreters), and white noise drawn in k-space inside the fused
kernel against the same stream drawn by the general-path evaluators.  Test infrastructure only; small grids (a
block costs a few hundred thread creations).  tests/test_step_gpu.py holds the parity tests proper.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import scipy.fft

from gopf_b200 import pf as gpf
from gopf_b200 import synthetic
from gopf_b200._lib import check, lib
from oracle import pf as opf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def fused(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("emul_fused") / "emul_fused.so"
    subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-pthread", "-DGOPF_KNOISE", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I",
                    os.path.join(ROOT, "tests", "host_emul"), "-I", os.path.join(ROOT, "gopf_b200", "csrc"), "-o", str(so),
                    os.path.join(ROOT, "tests", "host_emul", "emul_fused.cpp")], check=True)
    return ctypes.CDLL(str(so))


def _image(fn, *args, tail=()):
    need = ctypes.c_int64(0)
    check(fn(*args, None, ctypes.c_int64(0), ctypes.byref(need), *tail))
    buf = ctypes.create_string_buffer(need.value)
    check(fn(*args, buf, need, None, *tail))
    return buf


def fused_propagate(dll, m, field, dims, dt, nsteps, step0=0, tx=4, late=True):
    """Solver::propagate on the fused path: upload (forward transform), nsteps fused steps, download."""
    rank, edge = len(dims), dims[0]
    assert all(d == edge for d in dims)
    d_index = ctypes.c_int(-1)
    need = ctypes.c_int64(0)
    check(lib().gopf_model_fused_program_image(m._h, rank, ctypes.c_double(dt), None, ctypes.c_int64(0), ctypes.byref(need), None))
    assert need.value == dll.emul_fused_sizeof_program()
    prog = ctypes.create_string_buffer(need.value)
    check(lib().gopf_model_fused_program_image(m._h, rank, ctypes.c_double(dt), prog, need, None, ctypes.byref(d_index)))
    derived = _image(lib().gopf_model_derived_image, m._h, d_index.value, tail=(None,))
    S = np.ascontiguousarray(scipy.fft.fftn(field.Data.reshape(dims)).reshape(-1))
    rc = dll.emul_fused_steps(rank, edge, prog, derived, S.ctypes.data_as(DP), nsteps, ctypes.c_ulonglong(step0), tx, 1 if late else 0)
    assert rc == 0, rc
    field.Data[:] = scipy.fft.ifftn(S.reshape(dims)).reshape(-1)


def ch_pair(dims, seed=0):
    n = int(np.prod(dims))
    init = synthetic.cahn_hilliard_initial(n, seed)
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


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("late", [True, False], ids=["single-tile", "two-tile"])
@pytest.mark.parametrize("tx", [2, 4])
def test_cahn_hilliard_2d_100_steps_on_the_fused_kernels(fused, tx, late):
    # cfg 1 scaled down (examples/cahnHilliard/main.go): 10 epochs x 10 steps, <= 1e-10 (BASELINE.json north_star)
    dims = [32, 32]
    (gm, gf), (om, of) = ch_pair(dims)
    osolver = opf.NewSolver(om, dims, synthetic.CAHN_HILLIARD_DT)
    for epoch in range(10):
        fused_propagate(fused, gm, gf, dims, synthetic.CAHN_HILLIARD_DT, 10, step0=10 * epoch, tx=tx, late=late)
        osolver.Propagate(10)
    assert rel_l2(gf.Data, of.Data) <= 1e-10


@pytest.mark.parametrize("late", [True, False], ids=["single-tile", "two-tile"])
def test_cahn_hilliard_3d_on_the_fused_kernels(fused, late):
    # cfg 2 scaled down: the step of the headline number, against the oracle and the committed golden trajectory
    dims = [16, 16, 16]
    (gm, gf), (om, of) = ch_pair(dims)
    osolver = opf.NewSolver(om, dims, synthetic.CAHN_HILLIARD_DT)
    fused_propagate(fused, gm, gf, dims, synthetic.CAHN_HILLIARD_DT, 10, late=late)
    osolver.Propagate(10)
    assert rel_l2(gf.Data, of.Data) <= 1e-10
    golden = np.load(os.path.join(ROOT, "tests", "golden", "ch_3d_16_euler.npz"))
    assert rel_l2(gf.Data, golden["after_10"]) <= 1e-10


def test_general_program_on_the_fused_kernels(fused):
    """cfg 5 without noise: pair correlation (implicit, exp per k) and the ideal-mixture polynomial through the
    rolled interpreters of both fused kernels."""
    import test_step_gpu as T

    dims = [32, 32]
    (gm, gf, _), (om, of, osolver) = _pfc_pair(T, dims)
    fused_propagate(fused, gm, gf, dims, 0.1, 20)
    osolver.Propagate(20)
    assert rel_l2(gf.Data, of.Data) <= 1e-10


def _pfc_pair(T, dims):
    # tests/test_step_gpu.py pfc_models with the device solver left out (no GPU here)
    import types
    shim = types.SimpleNamespace(**{k: getattr(gpf, k) for k in dir(gpf) if not k.startswith("__")})
    shim.NewSolver = lambda m, d, dt, device=-1: None
    saved = T.gpf
    T.gpf = shim
    try:
        return T.pfc_models(dims, True)
    finally:
        T.gpf = saved


def test_kspace_noise_inside_the_fused_kernel_is_the_general_paths_stream(fused):
    """dconc/dt = LAP conc^3 + m1*LAP conc + NOISE with Model.SetKSpaceNoise: one derived field in use, so the model
    stays on the fused kernels; k_fused_kspace builds its k-point from per-axis tables, update_cells from the node
    number -- both must draw the same Hermitian spectrum (same Philox stream per frequency pair)."""
    from test_host_emulation_cpu import EmulatedSolver, emul as _emul_fixture  # noqa: F401

    dims = [16, 16, 16]
    n = 16 ** 3
    res = []
    for path in ("fused", "general"):
        m = gpf.NewModel()
        f = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 3))
        m.AddField(f)
        m.AddScalar(gpf.NewScalar("m1", -1.0))
        m.RegisterFunction("NOISE", gpf.WhiteNoise(1e-3, seed=11).Generate)
        m.AddEquation("dconc/dt = LAP conc^3 + m1*LAP conc + NOISE")
        m.SetKSpaceNoise(True)
        if path == "fused":
            fused_propagate(fused, m, f, dims, 0.01, 5)
        else:
            so = os.path.join(os.path.dirname(fused._name), "emul.so")
            subprocess.run(["g++", "-O1", "-ffp-contract=off", "-w", "-DGOPF_KNOISE", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I",
                            os.path.join(ROOT, "gopf_b200", "csrc"), "-o", so, os.path.join(ROOT, "tests", "host_emul", "emul.cpp")],
                           check=True)
            dll = ctypes.CDLL(so)
            dll.emul_sizeof_program.restype = ctypes.c_int
            dll.emul_sizeof_derived.restype = ctypes.c_int
            EmulatedSolver(dll, m, dims, 0.01).Propagate(5)
        res.append(f.Data.copy())
    assert np.max(np.abs(res[0].imag)) < 1e-12
    noise_free = synthetic.cahn_hilliard_initial(n, 3)
    assert rel_l2(res[0], noise_free) > 1e-3  # the noise is there
    assert rel_l2(res[0], res[1]) <= 1e-12
