"""Mirror of the reference's ``pfutil`` hot-path surface over the C ABI.

``FFTWWrapper`` keeps the reference's exported names (pfutil/fftWrap.go:8-95):
``NewFFTW(n)``, ``Dimensions``, ``FFT``, ``IFFT``, ``Freq``, ``ConjugateNode``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import check, int_array, lib


def ProdInt(a) -> int:
    res = 1
    for v in a:
        res *= int(v)
    return res


def NodeIdx(domain_size, idx) -> int:
    """pfutil.NodeIdx (pfutil/indexPositionConversion.go:14-22)."""
    if len(domain_size) != len(idx):
        raise ValueError("util: Domain size and idx has to be of length 2 or 3")
    out = ctypes.c_int64(0)
    check(lib().gopf_node_idx(len(domain_size), int_array(domain_size), int_array(idx), ctypes.byref(out)))
    return out.value


def Pos(domain_size, node_num: int):
    """pfutil.Pos (pfutil/indexPositionConversion.go:37-44)."""
    out = (ctypes.c_int * 3)()
    check(lib().gopf_pos(len(domain_size), int_array(domain_size), ctypes.c_int64(node_num), out))
    return [out[i] for i in range(len(domain_size))]


def _c128_ptr(a: np.ndarray):
    if a.dtype != np.complex128 or not a.flags.c_contiguous:
        raise TypeError("expected a C-contiguous complex128 array ([]complex128)")
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class FFTWWrapper:
    """pfutil.FFTWWrapper over ``gopf_fft_plan`` (transform-level API)."""

    def __init__(self, n, device: int = -1):
        self.Dimensions = [int(v) for v in n]
        self._h = ctypes.c_void_p()
        check(lib().gopf_fft_plan_create(len(self.Dimensions), int_array(self.Dimensions), device,
                                         ctypes.byref(self._h)))

    def FFT(self, data: np.ndarray) -> np.ndarray:
        check(lib().gopf_fft_exec(self._h, _c128_ptr(data), -1))
        return data

    def IFFT(self, data: np.ndarray) -> np.ndarray:
        check(lib().gopf_fft_exec(self._h, _c128_ptr(data), 1))
        return data

    def exec_device(self, dev_ptr: int, sign: int, stream: int = 0):
        check(lib().gopf_fft_exec_device(self._h, ctypes.c_void_p(dev_ptr), sign, ctypes.c_void_p(stream)))

    def Freq(self, i: int):
        out = (ctypes.c_double * 3)()
        check(lib().gopf_freq(len(self.Dimensions), int_array(self.Dimensions), ctypes.c_int64(i), out))
        return [out[k] for k in range(len(self.Dimensions))]

    def freq_device(self, nodes) -> np.ndarray:
        nodes = np.ascontiguousarray(nodes, dtype=np.int64)
        out = np.empty((nodes.shape[0], len(self.Dimensions)), dtype=np.float64)
        check(lib().gopf_fft_freq_device(self._h, nodes.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                         ctypes.c_int64(nodes.shape[0]),
                                         out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        return out

    # ---- gradient-based catalog terms at operator level (include/gopf_cuda.h) ----------------------------------
    def GradientCalculate(self, indata: np.ndarray, out: np.ndarray, comp: int, keep_nyquist: bool = False) -> np.ndarray:
        """GradientCalculator{FT: self, Comp, KeepNyquist}.Calculate(indata, out) (pf/gradientCalculator.go:19-31)."""
        check(lib().gopf_gradient_calculate(self._h, _c128_ptr(indata), _c128_ptr(out), int(comp), 1 if keep_nyquist else 0))
        return out

    def AdvectionConstruct(self, field: np.ndarray, velocity, out: np.ndarray, transformed: bool = False) -> np.ndarray:
        """-sum_d v_d GRAD_d(field): Advection.PrepareModel's derived field negated by Construct (pf/advection.go:50-96)."""
        vs = [np.ascontiguousarray(v, dtype=np.complex128) for v in velocity]
        arr = (ctypes.POINTER(ctypes.c_double) * len(vs))(*[_c128_ptr(v) for v in vs])
        check(lib().gopf_advection_construct(self._h, _c128_ptr(field), arr, len(vs), _c128_ptr(out), 1 if transformed else 0))
        return out

    def DivGradConstruct(self, field: np.ndarray, func_values: np.ndarray, out: np.ndarray) -> np.ndarray:
        """DivGrad.Construct over PrepareModel's derived fields (pf/gradientCalculator.go:72-108)."""
        check(lib().gopf_div_grad_construct(self._h, _c128_ptr(field), _c128_ptr(func_values), _c128_ptr(out)))
        return out

    def WeightedLaplacianConstruct(self, field_hat: np.ndarray, prefactor_hat: np.ndarray, out: np.ndarray) -> np.ndarray:
        """WeightedLaplacian.Construct (pf/gradientCalculator.go:131-172); both inputs are spectra."""
        check(lib().gopf_weighted_laplacian_construct(self._h, _c128_ptr(field_hat), _c128_ptr(prefactor_hat), _c128_ptr(out)))
        return out

    def ConjugateNode(self, i: int) -> int:
        out = ctypes.c_int64(0)
        check(lib().gopf_conjugate_node(len(self.Dimensions), int_array(self.Dimensions), ctypes.c_int64(i),
                                        ctypes.byref(out)))
        return out.value

    def close(self):
        if self._h:
            lib().gopf_fft_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def NewFFTW(n, device: int = -1) -> FFTWWrapper:
    """pfutil.NewFFTW (pfutil/fftWrap.go:16-23)."""
    return FFTWWrapper(n, device)


def TmaLaunchCount(reset: bool = False) -> int:
    """Launches that took the copy-engine-fed long-line kernels since the last reset (diagnostics)."""
    n = ctypes.c_int64(0)
    check(lib().gopf_tma_launch_count(1 if reset else 0, ctypes.byref(n)))
    return n.value
