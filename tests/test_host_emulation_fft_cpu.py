"""The hand-written FFT pass kernels (gopf_b200/csrc/fft_engine.cuh, fft_kernels.cuh) run on the HOST:
tests/host_emul/emul_fft.cpp compiles the kernels' own source with g++ and executes every CUDA thread of a
block as an OS thread, `__syncthreads()` as a barrier, the dynamic shared memory as one buffer per block.
Checked against the DFT definition (numpy's pocketfft): contiguous and strided passes, every tile width, forward
and inverse, scale on store -- and the forward pass with a derived field evaluated in its load, both with the
library's interpreter and with the loader the run-time specialisation generates (GOPF_JIT_INPASS: the very
translation unit NVRTC receives, jit::derived_pass_source).  Test infrastructure only; sizes are small because a
block costs a few hundred thread creations.  The device tests (tests/test_fft_gpu.py) remain the parity tests
proper.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from gopf_b200 import pf as gpf
from gopf_b200 import workloads
from gopf_b200._lib import check, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP = ctypes.POINTER(ctypes.c_double)
MAX_FIELDS = 4


def _build(out, extra=()):
    # -Bsymbolic: libgopfcuda.so (loaded RTLD_GLOBAL by gopf_b200._lib) exports host stubs of the same kernels
    # under the same names; the emulation must call its own host-compiled bodies, not those
    subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-pthread", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I", os.path.join(ROOT, "tests", "host_emul"),
                    "-I", os.path.join(ROOT, "gopf_b200", "csrc"), *extra, "-o", str(out),
                    os.path.join(ROOT, "tests", "host_emul", "emul_fft.cpp")], check=True)
    return ctypes.CDLL(str(out))


@pytest.fixture(scope="module")
def fft(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    return _build(tmp_path_factory.mktemp("emul_fft") / "emul_fft.so")


def _pass(dll, x, axis, inverse, scale=1.0, tx=0, in_place=False):
    n0, n1, n2 = x.shape
    src = np.array(x, dtype=np.complex128, order="C")  # a copy: in-place runs must not touch the caller's array
    out = src if in_place else np.empty_like(src)
    rc = dll.emul_fft_pass(n0, n1, n2, axis, 1 if inverse else 0, ctypes.c_double(scale), src.ctypes.data_as(DP),
                           out.ctypes.data_as(DP), tx)
    assert rc == 0, rc
    return out


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256])
def test_contiguous_pass_is_the_dft(fft, n):
    # FFTWWrapper.FFT / IFFT along the fastest axis (pfutil/fftWrap.go:26-39): sign -1 forward, +1 inverse, unnormalised
    rng = np.random.default_rng(n)
    x = rng.standard_normal((1, 5, n)) + 1j * rng.standard_normal((1, 5, n))  # 5 lines: a ragged last block
    assert rel(_pass(fft, x, 2, False), np.fft.fft(x, axis=2)) < 4e-16 * np.log2(n) + 1e-16
    assert rel(_pass(fft, x, 2, True), np.fft.ifft(x, axis=2) * n) < 4e-16 * np.log2(n) + 1e-16
    # the last inverse pass carries the 1/N of sliceOperations.go:34-40 on its store
    assert rel(_pass(fft, x, 2, True, scale=1.0 / n), np.fft.ifft(x, axis=2)) < 4e-16 * np.log2(n) + 1e-16


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
def test_long_lines(fft, n):
    """The line lengths of cfg 3 (1024^3) and beyond: three radix stages, the instantiations the device tests only
    reach on the largest grids."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal((1, 2, n)) + 1j * rng.standard_normal((1, 2, n))
    assert rel(_pass(fft, x, 2, False), np.fft.fft(x, axis=2)) < 1e-14
    assert rel(_pass(fft, x, 2, True, scale=1.0 / n), np.fft.ifft(x, axis=2)) < 1e-14
    if n <= 2048:  # strided: one tile of n x 2 cells
        y = rng.standard_normal((n, 1, 2)) + 1j * rng.standard_normal((n, 1, 2))
        assert rel(_pass(fft, y, 0, False, tx=2), np.fft.fft(y, axis=0)) < 1e-14
        assert rel(_pass(fft, y, 0, True, tx=2), np.fft.ifft(y, axis=0) * n) < 1e-14


@pytest.mark.parametrize("tx", [2, 4, 8])
@pytest.mark.parametrize("shape,axis", [((1, 32, 16), 1), ((16, 2, 8), 0), ((8, 64, 8), 1), ((128, 1, 8), 0)],
                         ids=["mid32", "slow16", "mid64", "slow128"])
def test_strided_pass_is_the_dft(fft, shape, axis, tx):
    rng = np.random.default_rng(17)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    n = shape[axis]
    assert rel(_pass(fft, x, axis, False, tx=tx), np.fft.fft(x, axis=axis)) < 4e-16 * np.log2(n) + 1e-16
    assert rel(_pass(fft, x, axis, True, tx=tx, in_place=True), np.fft.ifft(x, axis=axis) * n) < 4e-16 * np.log2(n) + 1e-16


def test_three_passes_make_the_3d_transform_and_round_trip(fft):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((8, 16, 32)) + 1j * rng.standard_normal((8, 16, 32))
    y = x
    for axis in (2, 1, 0):  # FftPlan::exec_device
        y = _pass(fft, y, axis, False, tx=4)
    assert rel(y, np.fft.fftn(x)) < 1e-15
    for axis in (0, 1, 2):  # Solver::inverse_to_real: 1/N on the last pass
        y = _pass(fft, y, axis, True, tx=4, scale=1.0 / x.size if axis == 2 else 1.0)
    assert rel(y, x) < 1e-15


# ---- the forward pass with a derived field in its load ------------------------------------------------------------------------
def _derived_image(m, index):
    need = ctypes.c_int64(0)
    check(lib().gopf_model_derived_image(m._h, index, None, ctypes.c_int64(0), ctypes.byref(need), None))
    buf = ctypes.create_string_buffer(need.value)
    check(lib().gopf_model_derived_image(m._h, index, buf, need, None, None))
    return buf


def _pass_derived(dll, image, fields, shape):
    n0, n1, n2 = shape
    ptrs = (DP * MAX_FIELDS)(*[fields[i].ctypes.data_as(DP) if i < len(fields) else None for i in range(MAX_FIELDS)])
    out = np.empty(shape, dtype=np.complex128)
    rc = dll.emul_fft_pass_derived(n0, n1, n2, image, ptrs, ctypes.c_ulonglong(0), out.ctypes.data_as(DP))
    assert rc == 0, rc
    return out


FUNCTIONS = {
    "DERIV_PHASE_ORDER": (workloads.DERIV_PHASE_EXPR,
                          lambda c, p: -(-0.5 * 0.1 * c.real ** 2 * workloads._dH(p.real) + 0.5 * 0.1 * (1.0 - c.real) ** 2 * workloads._dH(p.real)
                                         + 0.1 * workloads._dLandau(p.real))),
    "CHEMICALPOT": (workloads.CHEMICALPOT_EXPR,
                    lambda c, p: -((0.1 * c.real * (1.0 - workloads._H(p.real)) - 0.1 * (1.0 - c.real) * workloads._H(p.real)) * 1.0)),
    "WITH_IMAG": ("re(conc)*im(phase) + exp(0.1*phase) - im(conc)", lambda c, p: c.real * p.imag + np.exp(0.1 * p.real) - c.imag),
    "ONE_FIELD": ("dH(phase)", lambda c, p: workloads._dH(p.real)),
}


def _function_model(shape):
    n = int(np.prod(shape))
    m = gpf.NewModel()
    m.AddField(gpf.NewField("conc", n, np.zeros(n, dtype=np.complex128)))
    m.AddField(gpf.NewField("phase", n, np.zeros(n, dtype=np.complex128)))
    for name, (expr, _) in FUNCTIONS.items():
        m.RegisterFunction(name, expr)
    return m


@pytest.mark.parametrize("n", [16, 64, 256])
def test_forward_pass_with_interpreted_function_in_its_load(fft, n):
    shape = (1, 6, n)
    rng = np.random.default_rng(5)
    conc = rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1e-2, 1e-2, shape)
    phase = rng.uniform(-0.5, 1.5, shape) + 1j * rng.uniform(-1e-2, 1e-2, shape)
    m = _function_model(shape)
    for index, (name, (_, ref)) in enumerate(FUNCTIONS.items()):
        got = _pass_derived(fft, _derived_image(m, index), [conc, phase], shape)
        assert rel(got, np.fft.fft(ref(conc, phase), axis=2)) < 1e-14, name


@pytest.mark.parametrize("name", list(FUNCTIONS))
def test_forward_pass_with_the_generated_loader_on_the_host(name, tmp_path):
    """GOPF_JIT_INPASS: jit::derived_pass_source's translation unit -- the text NVRTC compiles -- built for the host."""
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    n = 64
    shape = (2, 5, n)
    m = _function_model(shape)
    unit = tmp_path / "unit.cu"
    unit.write_text(m.FunctionPassSource(name, n))
    dll = _build(tmp_path / "emul_fft_jit.so", extra=[f'-DEMUL_JIT_UNIT="{unit}"'])
    assert dll.emul_fft_has_jit_loader() == 1
    rng = np.random.default_rng(9)
    conc = rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1e-2, 1e-2, shape)
    phase = rng.uniform(-0.5, 1.5, shape) + 1j * rng.uniform(-1e-2, 1e-2, shape)
    index = list(FUNCTIONS).index(name)
    got = _pass_derived(dll, _derived_image(m, index), [conc, phase], shape)
    assert rel(got, np.fft.fft(FUNCTIONS[name][1](conc, phase), axis=2)) < 1e-14
    # plain loads of the same unit are untouched by the hook
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    assert rel(_pass(dll, x, 2, False), np.fft.fft(x, axis=2)) < 1e-15


# ---- SquaredGradient loads (pf/squareGradientTerm.go:38-65) ----------------------------------------------------------
def _wrap_freq(n):
    f = np.arange(n) / float(n)
    f[f > 0.5] -= 1.0
    return f


@pytest.mark.parametrize("shape", [(1, 8, 16), (8, 8, 8), (16, 16, 16)], ids=["2d", "3d-8", "3d-16"])
def test_gradient_multiplier_in_the_pass_along_its_axis_is_the_per_node_form(fft, shape):
    """The inverse transform of i 2 pi f_c u^ (Nyquist zeroed): the multiplier applied per node in the FIRST pass
    (LK_GRADIENT, literal FFTWWrapper.Freq) against the multiplier from the axis table applied in the pass ALONG the
    component's axis (LK_GRADIENT_LINE), and both against numpy."""
    n0, n1, n2 = shape
    rank = 2 if n0 == 1 else 3
    dims = [n1, n2] if rank == 2 else [n0, n1, n2]
    axes = [a for a in (0, 1, 2) if shape[a] > 1]
    rng = np.random.default_rng(8)
    u = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    axis_of_component = {0: 1, 1: 2, 2: 0}  # FftPlan::axis_of_component
    for comp in range(rank):
        ax_c = axis_of_component[comp]
        f = _wrap_freq(shape[ax_c])
        f[np.abs(f - 0.5) < 1e-10] = 0.0
        bshape = [1, 1, 1]
        bshape[ax_c] = shape[ax_c]
        want = np.fft.ifftn(u * (2j * np.pi * f.reshape(bshape)), axes=axes) * np.prod([shape[a] for a in axes])
        results = []
        for line in (False, True):
            x = np.array(u, dtype=np.complex128, order="C")
            for i, ax in enumerate(axes):
                out = np.empty_like(x)
                tx = 0 if ax == 2 else 2
                if line and ax == ax_c:
                    tab = np.ascontiguousarray(_wrap_freq(shape[ax]))
                    rc = fft.emul_fft_pass_gradient(n0, n1, n2, ax, rank, 0, 0, 0, comp, tab.ctypes.data_as(DP), x.ctypes.data_as(DP),
                                                    out.ctypes.data_as(DP), tx)
                elif not line and i == 0:
                    d = dims + [1] * (3 - len(dims))
                    rc = fft.emul_fft_pass_gradient(n0, n1, n2, ax, rank, d[0], d[1], d[2], comp, None, x.ctypes.data_as(DP),
                                                    out.ctypes.data_as(DP), tx)
                else:
                    rc = fft.emul_fft_pass(n0, n1, n2, ax, 1, ctypes.c_double(1.0), x.ctypes.data_as(DP), out.ctypes.data_as(DP), tx)
                assert rc == 0, rc
                x = out
            results.append(x)
        assert rel(results[0], want) < 1e-13, comp
        assert rel(results[1], want) < 1e-13, comp
        assert rel(results[1], results[0]) < 1e-13, comp


@pytest.mark.parametrize("shape,dim", [((1, 6, 16), 2), ((4, 4, 16), 3)], ids=["2d", "3d"])
def test_sum_of_squares_in_the_forward_pass_load(fft, shape, dim):
    n0, n1, n2 = shape
    rng = np.random.default_rng(9)
    g = [np.ascontiguousarray(rng.normal(size=shape) + 1j * rng.normal(size=shape)) for _ in range(dim)]
    out = np.zeros(shape, dtype=np.complex128)
    rc = fft.emul_fft_pass_sum_squares(n0, n1, n2, 2, dim, g[0].ctypes.data_as(DP), g[1].ctypes.data_as(DP),
                                       g[2].ctypes.data_as(DP) if dim > 2 else None, out.ctypes.data_as(DP), 0)
    assert rc == 0, rc
    assert rel(out, np.fft.fft(sum(x * x for x in g), axis=2)) < 1e-14
