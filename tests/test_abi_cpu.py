"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header
declares, and its host-side (integer / k-table) helpers agree with the oracle
bit for bit.  No compute entry point is called (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from gopf_b200 import _lib, pfutil as gpfutil
from oracle import pfutil as opfutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names += re.findall(r"\b(gopf_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 12
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.gopf_abi_version() == 1


def test_index_helpers_match_oracle_and_reference_kats():
    # pfutil/grid_test.go:8-36, pfutil/indexPositionConversion_test.go:5-25
    assert gpfutil.NodeIdx([3, 4], [1, 2]) == 6
    assert gpfutil.NodeIdx([3, 4, 2], [2, 3, 1]) == 23
    for dom in ([5, 7], [11, 12, 13], [3, 4, 2]):
        for node in range(opfutil.prod_int(dom)):
            p = gpfutil.Pos(dom, node)
            assert p == opfutil.pos(dom, node)
            assert gpfutil.NodeIdx(dom, p) == node == opfutil.node_idx(dom, p)


def _host_freq(dims, i):
    out = (ctypes.c_double * 3)()
    _lib.check(_lib.lib().gopf_freq(len(dims), _lib.int_array(dims), ctypes.c_int64(i), out))
    return [out[k] for k in range(len(dims))]


def _host_conj(dims, i):
    out = ctypes.c_int64(0)
    _lib.check(_lib.lib().gopf_conjugate_node(len(dims), _lib.int_array(dims), ctypes.c_int64(i), ctypes.byref(out)))
    return out.value


@pytest.mark.parametrize("dims", [[8, 16], [9, 9], [8, 8, 8], [9, 9, 9], [4, 6, 5], [128, 128]])
def test_host_freq_and_conjugate_bit_exact(dims):
    # pfutil/fftwWrap_test.go:22-29, 59-90
    ft = opfutil.NewFFTW(dims)
    n = opfutil.prod_int(dims)
    step = max(1, n // 2000)
    for i in list(range(0, n, step)) + [n - 1]:
        assert _host_freq(dims, i) == ft.Freq(i)
        assert _host_conj(dims, i) == ft.ConjugateNode(i)


def test_errors_are_reported_not_swallowed():
    out = ctypes.c_int64(0)
    st = _lib.lib().gopf_node_idx(4, _lib.int_array([2, 2, 2, 2]), _lib.int_array([0, 0, 0, 0]), ctypes.byref(out))
    assert st != 0
    assert b"length 2 or 3" in _lib.lib().gopf_last_error()
