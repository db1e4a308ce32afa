"""gopf_b200: B200-native (sm_100a) implementation of gopf's spectral
time-stepping hot path behind the reference's own interface names.

The product is ``lib/libgopfcuda.so`` (C ABI: include/gopf_cuda.h).  This Python
package is the ctypes binding that tests and bench.py drive; it mirrors the Go
names (``pfutil.NewFFTW``, ``pf.NewModel`` ...) so parity tests read like the
reference's own tests.
"""
from ._lib import GopfError, LIB_PATH  # noqa: F401
