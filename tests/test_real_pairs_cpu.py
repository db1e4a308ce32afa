"""The arithmetic of k_fused_real_pair_tma (gopf_b200/csrc/tma_kernels.cuh) restated in numpy, step for step, against
the plain per-line computation it replaces (k_fused_real: inverse transform, /N, g, forward transform; the last
inverse and first forward pass of pf/euler.go:16-47 around the real-space nonlinearity).  The kernel itself runs in
tests/test_tma_gpu.py; this pins the identities it relies on where no GPU is needed."""
import numpy as np
import pytest


def _dft(x):  # unnormalised forward transform, sign -1 (pfutil/fftWrap.go:26-31)
    return np.fft.fft(x)


def _swap(z):
    return z.imag + 1j * z.real


def paired_lines(A, B, g):
    """Two Hermitian lines A, B (transforms of real lines) -> transforms of g(a), g(b), the way the kernel does it."""
    N = A.shape[0]
    # v = cswap(A + iB); one forward DFT; swap back, /N: z = a + i b
    v = (A.imag + B.real) + 1j * (A.real - B.imag)
    v = _dft(v)
    a, b = v.imag / N, v.real / N
    # g on each real line, one forward DFT of g(a) + i g(b)
    Z = _dft(g(a) + 1j * g(b))
    q = Z[(N - np.arange(N)) % N]                      # partner Z'[N - k]
    GA = 0.5 * (Z.real + q.real) + 0.5j * (Z.imag - q.imag)
    GB = 0.5 * (Z.imag + q.imag) + 0.5j * (q.real - Z.real)
    return GA, GB


@pytest.mark.parametrize("N", [16, 512, 1024])
@pytest.mark.parametrize("power", [2, 3, 5])
def test_two_real_lines_through_one_complex_transform(N, power):
    rng = np.random.default_rng(N + power)
    a, b = rng.uniform(-1, 1, N), rng.uniform(-1, 1, N)
    A, B = _dft(a.astype(np.complex128)), _dft(b.astype(np.complex128))
    g = lambda x: x ** power
    GA, GB = paired_lines(A, B, g)
    # what k_fused_real does per line: swap(DFT(swap(A))) / N = a, then DFT(g(a))
    for line, got in ((A, GA), (B, GB)):
        real_line = _swap(_dft(_swap(line))) / N
        assert np.max(np.abs(real_line.imag)) < 1e-13
        want = _dft(real_line ** power)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13


def test_an_imaginary_residue_leaks_into_the_partner_at_its_own_size():
    """Why the paired kernel is only chosen for fields without an imaginary part: the residue eps of line A ends up in
    line B at O(eps) -- harmless at rounding level (1e-17), wrong for a genuinely complex field."""
    N = 64
    rng = np.random.default_rng(1)
    a, b = rng.uniform(-1, 1, N), rng.uniform(-1, 1, N)
    eps = 1e-9
    A = _dft((a + 1j * eps * rng.uniform(-1, 1, N)).astype(np.complex128))
    B = _dft(b.astype(np.complex128))
    g = lambda x: x ** 3
    _, GB = paired_lines(A, B, g)
    want = _dft((b ** 3).astype(np.complex128))
    err = np.linalg.norm(GB - want) / np.linalg.norm(want)
    assert 1e-11 < err < 1e-7
