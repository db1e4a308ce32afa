"""Oracle restatement of the reference's ``pfutil`` pieces that sit on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations: /root/reference.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as _sfft


def prod_int(a) -> int:
    """pfutil/sliceOperations.go:43-49 ProdInt."""
    res = 1
    for v in a:
        res *= int(v)
    return res


# --------------------------------------------------------------------------
# index <-> position   (pfutil/indexPositionConversion.go:4-44)
# --------------------------------------------------------------------------
def node_idx(domain_size, idx) -> int:
    """NodeIdx: 2-D ``r*n1 + c``; 3-D ``d*n0*n1 + r*n1 + c`` with idx = [r, c, d]."""
    if len(domain_size) == 2 and len(idx) == 2:
        return idx[0] * domain_size[1] + idx[1]
    if len(domain_size) == 3 and len(idx) == 3:
        return idx[2] * domain_size[0] * domain_size[1] + idx[0] * domain_size[1] + idx[1]
    raise ValueError("util: Domain size and idx has to be of length 2 or 3")


def pos(domain_size, node_num: int):
    """Pos: inverse of NodeIdx (Go integer division truncates; inputs non-negative)."""
    if len(domain_size) == 2:
        return [node_num // domain_size[1], node_num % domain_size[1]]
    if len(domain_size) == 3:
        col = node_num % domain_size[1]
        row = (node_num // domain_size[1]) % domain_size[0]
        depth = node_num // (domain_size[0] * domain_size[1])
        return [row, col, depth]
    raise ValueError("util: Domain size has to be either 2 or 3")


# --------------------------------------------------------------------------
# FFTWWrapper   (pfutil/fftWrap.go:8-95)
# --------------------------------------------------------------------------
class FFTWWrapper:
    """c2c DFT over a row-major array of shape ``n`` (last axis fastest).

    FFT: sign -1, unnormalised.  IFFT: sign +1, unnormalised (the caller divides
    by N, pfutil/sliceOperations.go:34-40).  Both act in place on the caller's
    array and return it (fftWrap.go:26-39).
    """

    def __init__(self, n, workers: int = 1):
        self.Dimensions = [int(v) for v in n]
        self.N = prod_int(self.Dimensions)
        self.workers = workers

    # fftWrap.go:26-31
    def FFT(self, data: np.ndarray) -> np.ndarray:
        v = data.reshape(self.Dimensions)
        v[...] = _sfft.fftn(v, workers=self.workers)
        return data

    # fftWrap.go:34-39
    def IFFT(self, data: np.ndarray) -> np.ndarray:
        v = data.reshape(self.Dimensions)
        v[...] = _sfft.ifftn(v, norm="forward", workers=self.workers)
        return data

    # fftWrap.go:42-54
    def col(self, i):
        return i % self.Dimensions[1]

    def row(self, i):
        return (i // self.Dimensions[1]) % self.Dimensions[0]

    def depth(self, i):
        return i // (self.Dimensions[0] * self.Dimensions[1])

    # fftWrap.go:57-74 -- scalar form
    def Freq(self, i: int):
        d = self.Dimensions
        res = [0.0] * len(d)
        res[1] = float(self.col(i)) / float(d[1])
        res[0] = float(self.row(i)) / float(d[0])
        if len(res) > 2:
            res[2] = float(self.depth(i)) / float(d[2])
        for j in range(len(res)):
            if res[j] > 0.5:
                res[j] -= 1.0
        return res

    def freq_table(self) -> np.ndarray:
        """Freq(i) for every node as an (N, dim) float64 array, bit-identical to
        the scalar formula (one IEEE divide, then an exact subtract)."""
        d = self.Dimensions
        i = np.arange(self.N, dtype=np.int64)
        out = np.empty((self.N, len(d)), dtype=np.float64)
        out[:, 1] = (i % d[1]).astype(np.float64) / float(d[1])
        out[:, 0] = ((i // d[1]) % d[0]).astype(np.float64) / float(d[0])
        if len(d) > 2:
            out[:, 2] = (i // (d[0] * d[1])).astype(np.float64) / float(d[2])
        out[out > 0.5] -= 1.0
        return out

    # fftWrap.go:78-95
    def ConjugateNode(self, i: int) -> int:
        d = self.Dimensions
        c = self.col(i)
        r = self.row(i)
        conj_c = (d[1] - c) % d[1]
        conj_r = (d[0] - r) % d[0]
        conj_d = 0
        if len(d) == 3:
            conj_d = (d[2] - self.depth(i)) % d[2]
        return conj_d * d[0] * d[1] + conj_r * d[1] + conj_c


def NewFFTW(n, workers: int = 1) -> FFTWWrapper:
    """fftWrap.go:16-23."""
    return FFTWWrapper(n, workers)


# --------------------------------------------------------------------------
# slice helpers used inside the hot loops (pfutil/sliceOperations.go:20-90)
# --------------------------------------------------------------------------
def cmplx_equal_approx(a, b, tol) -> bool:
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    return bool(np.all(np.abs(a.real - b.real) <= tol) and np.all(np.abs(a.imag - b.imag) <= tol))


def div_real_scalar(data: np.ndarray, factor: float) -> np.ndarray:
    data /= complex(factor, 0.0)
    return data


def go_cpow(x: np.ndarray, p: float) -> np.ndarray:
    """Go ``cmplx.Pow(x, complex(p, 0))`` (math/cmplx/pow.go): polar form
    ``|x|^p * (cos(p*phase), sin(p*phase))``; x == 0 -> 0 (p > 0), 1 (p == 0),
    Inf (p < 0).  Leaves the same O(1e-16) imaginary residue on negative reals
    the reference carries ('hard parts')."""
    x = np.asarray(x, dtype=np.complex128)
    if p == 0.0:
        return np.ones_like(x)
    modulus = np.abs(x)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.power(modulus, p)
        theta = p * np.angle(x)
        out = r * (np.cos(theta) + 1j * np.sin(theta))
    zero = x == 0
    if np.any(zero):
        out = np.where(zero, complex(np.inf, 0.0) if p < 0 else 0.0, out)
    return out


# --------------------------------------------------------------------------
# synthetic inputs -- SplitMix64, shared definition with gopf_b200.synthetic
# (kept as an independent restatement so the oracle does not import the product)
# --------------------------------------------------------------------------
_MASK = (1 << 64) - 1


def splitmix64_uniform(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """u[j] in [0,1): SplitMix64 output of counter (seed + (offset+j+1)*golden),
    top 53 bits / 2^53.  Counter based so any slab can be generated alone."""
    with np.errstate(over="ignore"):
        j = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
        z = (np.uint64(seed & _MASK) + j * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
