"""Pins oracle/pfutil.py against the reference's own known-answer tests.

Each test cites the Go test it restates (/root/reference).  No GPU needed.
"""
import math

import numpy as np

from oracle import pfutil


def naive_dftn(x: np.ndarray, sign: int) -> np.ndarray:
    """Independent O(n^2)-per-axis DFT straight from the definition (FFTW manual,
    'What FFTW really computes': Y_k = sum_j X_j exp(sign 2 pi i j k / n))."""
    out = x.astype(np.complex128)
    for ax, n in enumerate(x.shape):
        j = np.arange(n)
        w = np.exp(sign * 2j * np.pi * np.outer(j, j) / n)
        out = np.moveaxis(np.tensordot(w, out, axes=([1], [ax])), 0, ax)
    return out


def test_index_kat():
    # pfutil/grid_test.go:8-36 (Grid.Index == NodeIdx, pfutil/grid.go:37)
    assert pfutil.node_idx([3, 4], [1, 2]) == 6
    assert pfutil.node_idx([3, 4, 2], [2, 3, 1]) == 12 + 2 * 4 + 3
    assert pfutil.pos([3, 4], 6) == [1, 2]
    assert pfutil.pos([3, 4, 2], 23) == [2, 3, 1]


def test_node2pos_round_trip():
    # pfutil/indexPositionConversion_test.go:5-25
    for node, dom in [(16, [5, 7]), (124, [11, 12, 13])]:
        assert pfutil.node_idx(dom, pfutil.pos(dom, node)) == node
    # exhaustive for the same shapes
    for dom in ([5, 7], [11, 12, 13]):
        for node in range(pfutil.prod_int(dom)):
            assert pfutil.node_idx(dom, pfutil.pos(dom, node)) == node


def test_freq_matches_fftfreq_convention():
    # pfutil/fftwWrap_test.go:22-29: Freq == gosfft Freq on 8x16 (tol 1e-8).  gosfft's
    # Freq is the standard fftfreq convention except Nyquist stays +0.5
    # (fftWrap.go:69 strict '>').
    nx, ny = 8, 16
    ft = pfutil.NewFFTW([nx, ny])
    fx = np.fft.fftfreq(nx)
    fy = np.fft.fftfreq(ny)
    for i in range(nx * ny):
        f = ft.Freq(i)
        r, c = divmod(i, ny)
        ex, ey = abs(fx[r]) if r == nx // 2 else fx[r], abs(fy[c]) if c == ny // 2 else fy[c]
        assert abs(f[0] - ex) < 1e-8 and abs(f[1] - ey) < 1e-8
    assert ft.Freq((nx // 2) * ny)[0] == 0.5  # Nyquist is +0.5, not -0.5


def test_freq_table_bit_exact_with_scalar():
    for dims in ([8, 16], [9, 9], [8, 8, 8], [9, 9, 9], [4, 6, 5]):
        ft = pfutil.NewFFTW(dims)
        tab = ft.freq_table()
        for i in range(ft.N):
            assert list(tab[i]) == ft.Freq(i)  # exact equality


def test_conjugate_node():
    # pfutil/fftwWrap_test.go:59-90
    tol = 1e-10
    for dims in ([8, 8], [9, 9], [8, 8, 8], [9, 9, 9]):
        ft = pfutil.NewFFTW(dims)
        for j in range(pfutil.prod_int(dims)):
            f1 = ft.Freq(j)
            f2 = ft.Freq(ft.ConjugateNode(j))
            for k in range(len(f1)):
                assert not (abs(f1[k] + f2[k]) > tol and abs(f1[k]) < 0.5 - tol)


def test_fft_ramp_8x16_against_definition():
    # pfutil/fftwWrap_test.go:11-57: FFTW == gosfft on data[i] = i, 8x16, forward then
    # inverse, tol 1e-6.  Both are the plain DFT; the independent check here is the
    # O(n^2) definition.
    nx, ny = 8, 16
    data = np.arange(nx * ny, dtype=np.float64).astype(np.complex128)
    ft = pfutil.NewFFTW([nx, ny])
    ref = naive_dftn(data.reshape(nx, ny), -1).reshape(-1)
    out = ft.FFT(data.copy())
    assert np.max(np.abs(out - ref)) < 1e-9
    assert abs(out[0] - data.sum()) < 1e-9
    back = ft.IFFT(out.copy())
    ref_back = naive_dftn(ref.reshape(nx, ny), +1).reshape(-1)
    assert np.max(np.abs(back - ref_back)) < 1e-8
    assert np.max(np.abs(back / (nx * ny) - data)) < 1e-10  # unnormalised inverse


def test_fft_3d_against_definition_and_in_place():
    rng = np.random.default_rng(1)
    dims = [4, 8, 16]
    data = rng.standard_normal(512) + 1j * rng.standard_normal(512)
    ft = pfutil.NewFFTW(dims)
    ref = naive_dftn(data.reshape(dims), -1).reshape(-1)
    buf = data.copy()
    ret = ft.FFT(buf)
    assert ret is buf  # in place, returns the caller's slice (fftWrap.go:26-31)
    assert np.max(np.abs(buf - ref)) < 1e-10


def test_go_cpow_matches_go_semantics():
    # math/cmplx.Pow polar form: (-2)^3 carries a ~1e-15 imaginary residue
    v = pfutil.go_cpow(np.array([-2.0 + 0j, 2.0 + 0j, 0j]), 3.0)
    assert abs(v[0].real + 8.0) < 1e-14 and 0 < abs(v[0].imag) < 1e-14
    assert v[1] == 8.0 and v[2] == 0.0
    assert pfutil.go_cpow(np.array([0j]), 0.0)[0] == 1.0


def test_splitmix_reference_vector():
    # SplitMix64 (Steele/Lea/Flood 2014) first outputs for seed 0: published test vector
    # 0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F
    u = pfutil.splitmix64_uniform(0, 3)
    exp = [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    for a, b in zip(u, exp):
        assert a == (b >> 11) / 9007199254740992.0
    # counter based: a slab generated alone equals the slice of the whole
    assert np.array_equal(pfutil.splitmix64_uniform(7, 10, offset=5), pfutil.splitmix64_uniform(7, 15)[5:])
