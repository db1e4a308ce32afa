"""Run-time specialisation (gopf_b200/csrc/jit.h), the parts that need no GPU:

* the C a registered function (Model.RegisterFunction, pf/model.go:400-412) is turned into is compiled
  with gcc and evaluated against numpy on the same inputs -- the reference's own closures for cfg 4
  (examples/strain_single_precipitate/main.go:18-64) and seeded random expressions;
* NVRTC turns the generated kernels and the specialised k-space update (pf/euler.go:27-39) of cfg 4 /
  cfg 5 into sm_100a images, and the images are fully folded (no stack frame, no program loads).
"""
import ctypes
import os
import re
import shutil
import subprocess
import types

import numpy as np
import pytest

from gopf_b200 import elasticity as gel
from gopf_b200 import pf as gpf
from gopf_b200 import workloads


def _model(n=4):
    m = gpf.NewModel()
    m.AddField(gpf.NewField("conc", n, np.zeros(n, dtype=np.complex128)))
    m.AddField(gpf.NewField("phase", n, np.zeros(n, dtype=np.complex128)))
    m.AddScalar(gpf.NewScalar("kappa", 0.37))
    return m


NUMPY_NS = {
    "re": lambda z: z.real, "im": lambda z: z.imag,
    "H": workloads._H, "dH": workloads._dH, "dLandau": workloads._dLandau,
    "Landau": lambda x: x * x - 2.0 * x * x * x + x * x * x * x,
    "exp": np.exp, "log": np.log, "sin": np.sin, "cos": np.cos, "tanh": np.tanh, "sqrt": np.sqrt, "abs": np.abs,
    "negpart": lambda x: np.minimum(x, 0.0), "pi": np.pi, "kappa": 0.37,
}


def numpy_eval(expr, conc, phase):
    """The expression language is Python's arithmetic with '^' for '**'; bare field names are real parts."""
    src = expr.replace("^", "**")
    src = re.sub(r"\b(re|im)\((conc|phase)\)", r"\1(_\2)", src)
    ns = dict(NUMPY_NS, conc=conc.real, phase=phase.real, _conc=conc, _phase=phase)
    return eval(src, {"__builtins__": {}}, ns) + np.zeros(conc.shape)


def random_expression(rng, depth=0):
    atoms = ["conc", "phase", "re(conc)", "im(phase)", "im(conc)", "kappa", "pi", "0.25", "1.5", "2.0", "3e-1"]
    if depth >= 3 or rng.random() < 0.25:
        return atoms[rng.integers(len(atoms))]
    a, b = random_expression(rng, depth + 1), random_expression(rng, depth + 1)
    kind = rng.integers(12)
    if kind < 4:
        return f"({a} {'+-*'[rng.integers(3)]} {b})"
    if kind == 4:
        return f"({a} / (abs({b}) + 0.5))"
    if kind == 5:
        return f"(-{a})"
    if kind == 6:
        return f"({a})^{rng.integers(0, 5)}"
    if kind == 7:
        return f"(abs({a}) + 0.5)^(0.3*{b})"
    if kind == 8:
        return f"{['H', 'dH', 'Landau', 'dLandau'][rng.integers(4)]}({a})"
    if kind == 9:
        return f"{['sin', 'cos', 'tanh', 'abs', 'negpart'][rng.integers(5)]}({a})"
    if kind == 10:
        return f"exp(0.1*tanh({a}))"
    return f"{['log', 'sqrt'][rng.integers(2)]}(abs({a}) + 0.1)"


FIXED = [
    workloads.CHEMICALPOT_EXPR,
    workloads.DERIV_PHASE_EXPR,
    workloads.SMEARING_EXPR,
    gpf.NegativeValuePenalty(1500.0, 3, "conc").Evaluate,
    f"(1.0 - conc*conc)*(conc + {3.0 * 0.05 / 4.0!r})",  # pf/sdd_test.go:163-246 chemical potential
    "3.0*(-0.16666666666666666)*re(conc)*re(conc)+4.0*(0.08333333333333333)*re(conc)*re(conc)*re(conc)",
    "0.0 - 0.0", "-0.0*conc", "1e308*10.0", "2.0^10 + conc^0",
]


@pytest.fixture(scope="module")
def compiled_functions(tmp_path_factory):
    """Every test expression as generated C, built into one shared object by gcc."""
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    rng = np.random.default_rng(20261017)
    exprs = list(FIXED)
    while len(exprs) < len(FIXED) + 60:
        e = random_expression(rng)
        try:  # the device program is bounded (GOPF_MAX_RPN ops, GOPF_RPN_STACK operands)
            m = _model()
            m.RegisterFunction("F", e)
        except gpf.GopfError:
            continue
        exprs.append(e)
    d = tmp_path_factory.mktemp("jit")
    unit = "#include <math.h>\n"
    for i, e in enumerate(exprs):
        m = _model()
        m.RegisterFunction("F", e)
        src = m.FunctionSource("F")
        assert "gopf_expr(" in src
        unit += f"#define gopf_expr gopf_expr_{i}\n#define gopf_ipow gopf_ipow_{i}\n{src}\n#undef gopf_expr\n#undef gopf_ipow\n"
        unit += (f"void eval_{i}(const double* c, const double* p, double* out, long n) {{\n"
                 f"    for (long k = 0; k < n; ++k) out[k] = gopf_expr_{i}(c[2*k], c[2*k+1], p[2*k], p[2*k+1], 0.0, 0.0, 0.0, 0.0);\n}}\n")
    (d / "f.c").write_text(unit)
    so = d / "f.so"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(d / "f.c"), "-lm"], check=True)
    return exprs, ctypes.CDLL(str(so))


def test_generated_c_of_registered_functions_matches_numpy(compiled_functions):
    exprs, so = compiled_functions
    rng = np.random.default_rng(7)
    n = 4096
    conc = rng.uniform(-1.5, 1.5, n) + 1j * rng.uniform(-1e-3, 1e-3, n)
    phase = rng.uniform(-0.5, 1.5, n) + 1j * rng.uniform(-1e-3, 1e-3, n)
    conc[:4] = [0.0, 1.0, -1.0, 0.5]
    phase[:4] = [0.0, 1.0, 0.5, -0.0]
    dp = ctypes.POINTER(ctypes.c_double)
    worst = 0.0
    for i, e in enumerate(exprs):
        out = np.empty(n)
        getattr(so, f"eval_{i}")(conc.ctypes.data_as(dp), phase.ctypes.data_as(dp), out.ctypes.data_as(dp), ctypes.c_long(n))
        with np.errstate(all="ignore"):
            want = numpy_eval(e, conc, phase)
        assert np.array_equal(np.isnan(out), np.isnan(want)) and np.array_equal(np.isinf(out), np.isinf(want)), e
        ok = np.isfinite(want)
        scale = max(1.0, float(np.max(np.abs(want[ok])))) if ok.any() else 1.0
        err = float(np.max(np.abs(out[ok] - want[ok]))) / scale if ok.any() else 0.0
        worst = max(worst, err)
        assert err < 1e-12, f"{e}: {err}"
    assert len(exprs) >= 60 and worst < 1e-12


def test_function_source_is_refused_for_other_derived_fields():
    m = _model()
    m.RegisterFunction("NOISE", gpf.WhiteNoise(1e-3, seed=1).Generate)
    with pytest.raises(gpf.GopfError, match="not a registered function"):
        m.FunctionSource("NOISE")
    with pytest.raises(gpf.GopfError, match="no derived field"):
        m.FunctionSource("nope")


def _resource_usage(cubin, function=None):
    out = subprocess.run(["cuobjdump", "-res-usage", cubin], check=True, capture_output=True, text=True).stdout
    pattern = r"REG:(\d+) STACK:(\d+)" if function is None else rf"Function {function}:\s+REG:(\d+) STACK:(\d+)"
    m = re.search(pattern, out)
    return int(m.group(1)), int(m.group(2))


def _sass_counts(cubin, function):
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", function, cubin], check=True, capture_output=True, text=True).stdout
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", sass, flags=re.M)
    return ops


needs_nvrtc = pytest.mark.skipif(not os.path.exists("/usr/local/cuda/lib64/libnvrtc.so.12"), reason="NVRTC not installed")


@needs_nvrtc
def test_registered_function_kernels_compile_for_sm100a(tmp_path, monkeypatch):
    monkeypatch.setenv("GOPF_JIT_DUMP", str(tmp_path))
    m = _model()
    m.RegisterFunction("CHEMICALPOT", workloads.CHEMICALPOT_EXPR)
    m.RegisterFunction("DERIV_PHASE_ORDER", workloads.DERIV_PHASE_EXPR)
    m.RegisterFunction("SMEARING_DERIV", workloads.SMEARING_EXPR)
    for name in ("CHEMICALPOT", "DERIV_PHASE_ORDER", "SMEARING_DERIV"):
        assert "gopf_jit_derived" in m.FunctionSource(name, kernel=True)
        assert m.FunctionCompile(name) > 1000
    if shutil.which("cuobjdump"):
        for cubin in sorted(tmp_path.glob("*.cubin")):
            regs, stack = _resource_usage(str(cubin))
            assert stack == 0 and regs <= 32, (cubin.name, regs, stack)  # 2048 resident threads per SM


@needs_nvrtc
@pytest.mark.parametrize("line_length", [64, 256, 1024])
def test_forward_pass_with_the_function_in_its_load_compiles(line_length, tmp_path, monkeypatch):
    """GOPF_JIT_INPASS: the library's own k_pass_contig<N>, recompiled by NVRTC from the embedded headers with a
    generated loader.  The image must keep the library kernel's frame (no spills of the register-resident line)."""
    monkeypatch.setenv("GOPF_JIT_DUMP", str(tmp_path))
    m = _model()
    m.RegisterFunction("DERIV_PHASE_ORDER", workloads.DERIV_PHASE_EXPR)
    m.RegisterFunction("WITH_IMAG", "re(conc)*im(phase) + exp(0.1*phase)")
    src = m.FunctionPassSource("DERIV_PHASE_ORDER", line_length)
    assert '#include "fft_kernels.cuh"' in src and "gopf_jit_load_line" in src and f"k_pass_contig<{line_length}>" in src
    assert "p1[2 * at(m0 + j)]" in src  # real parts only when the function reads nothing else
    assert "c1[j] = io.R.r[1][at(m0 + j)]" in m.FunctionPassSource("WITH_IMAG", line_length)
    for name in ("DERIV_PHASE_ORDER", "WITH_IMAG"):
        size, lowered = m.FunctionPassCompile(name, line_length)
        assert size > 10000 and lowered.startswith("_ZN4gopf13k_pass_contigILi%dEEE" % line_length)
    if shutil.which("cuobjdump"):
        lib_kernel = subprocess.run(["cuobjdump", "-res-usage", os.path.join(os.path.dirname(gpf.__file__), "lib", "libgopfcuda.so")],
                                    check=True, capture_output=True, text=True).stdout
        ref = re.search(rf"k_pass_contigILi{line_length}EEE\S*:\s+REG:(\d+) STACK:(\d+)", lib_kernel)
        for cubin in sorted(tmp_path.glob("*.cubin")):
            regs, stack = _resource_usage(str(cubin))
            # (when torch is in the process its bundled NVRTC 12.8 answers the dlopen; it frames 16 B more)
            assert stack <= int(ref.group(2)) + 32 and regs <= 128, (cubin.name, regs, stack, ref.groups())


class _NoSolver:
    class Stepper:
        @staticmethod
        def SetFilter(f):
            pass


def _host_only(module):
    """The pf surface without NewSolver, so that the workload builders run without a GPU."""
    shim = types.SimpleNamespace(**{k: getattr(module, k) for k in dir(module) if not k.startswith("__")})
    shim.NewSolver = lambda m, dims, dt: _NoSolver()
    return shim


@needs_nvrtc
@pytest.mark.parametrize("kind,tab_mask,with_filter", [("precipitate", 0, False), ("pfc", 1, True), ("pfc", 0, True)])
def test_kspace_update_specialises_to_a_folded_image(kind, tab_mask, with_filter, tmp_path, monkeypatch):
    monkeypatch.setenv("GOPF_JIT_DUMP", str(tmp_path))
    shim = _host_only(gpf)
    dims = [8, 8, 8]
    if kind == "precipitate":
        m, dt = workloads.build_precipitate(shim, shim, gel, dims, expressions=True)[0], workloads.PRECIPITATE_DT
    else:
        m, dt = workloads.build_pfc(shim, shim, dims, noise="device")[0], workloads.PFC_DT
    stand_in = dict(filter_addr=0x7F0000000000 if with_filter else 0, filter_n=1000, lp_addr=0x7F0000100000)
    src = m.KUpdateSource(dims, dt, tab_mask, **stand_in)
    assert '#include "kupdate.cuh"' in src and "jit_prog_words" in src and "{3, 8, 8, 8}" in src
    assert ("GOPF_FILTER(P) ((const double*)0x0ull)" in src) == (not with_filter)
    assert m.KUpdateCompile(dims, dt, tab_mask, **stand_in) > 1000
    if not shutil.which("cuobjdump"):
        return
    cubin = str(sorted(tmp_path.glob("*.cubin"))[-1])
    for fn in ("gopf_jit_rk4_rhs", "gopf_jit_rk4_point"):  # the RK4 passes of the same program
        regs, stack = _resource_usage(cubin, fn)
        rk_ops = _sass_counts(cubin, fn)
        assert stack == 0 and len(rk_ops) > 50 and not any(o.startswith(("LDL", "STL")) for o in rk_ops), (fn, regs, stack)
    regs, stack = _resource_usage(cubin, "gopf_jit_kupdate")
    ops = _sass_counts(cubin, "gopf_jit_kupdate")
    assert stack == 0 and regs <= 64 and len(ops) > 100, (regs, stack, len(ops))
    assert not any(o.startswith(("LDL", "STL")) for o in ops)
    # the program image is folded away: the only global loads left are the 16-byte spectrum / table cells
    loads = [o for o in ops if o.startswith("LDG")]
    assert loads and all(o.startswith("LDG.E.128") or o.startswith("LDG.E.64") for o in loads), sorted(set(loads))
    assert len(loads) <= 12, len(loads)


def test_specialisation_refuses_a_null_multiplier_address():
    shim = _host_only(gpf)
    m = workloads.build_precipitate(shim, shim, gel, [8, 8, 8], expressions=True)[0]
    with pytest.raises(gpf.GopfError, match="without a multiplier address"):
        m.KUpdateSource([8, 8, 8], workloads.PRECIPITATE_DT, 0, lp_addr=0)
