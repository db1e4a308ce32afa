"""The copy-engine-fed (TMA, warp-specialised) long-line kernels of tma_kernels.cuh against the
register-resident kernels they replace and against the oracle.

They run the same Stockham engine on the same cells, so the results must be BITWISE those of the
register-resident kernels (GOPF_TMA=0); the oracle comparison is the usual 1e-10 / 1e-13.  Reference
semantics: pfutil/fftWrap.go:26-39 (transform), pf/euler.go:16-47 (step)."""
import os

import numpy as np
import pytest

from gopf_b200 import pf as gpf
from gopf_b200 import pfutil as gpfutil
from gopf_b200 import synthetic
from oracle import pfutil as opfutil

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tma_env():
    keys = ("GOPF_TMA", "GOPF_TMA_MIN_N", "GOPF_TMA_L2", "GOPF_TMA_PASS", "GOPF_TMA_REAL", "GOPF_TMA_KSPACE", "GOPF_REAL_PAIRS")
    saved = {k: os.environ.get(k) for k in keys}
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _fft(dims, x, sign):
    y = x.copy()
    ft = gpfutil.NewFFTW(dims)
    (ft.FFT if sign < 0 else ft.IFFT)(y)
    ft.close()
    return y


@pytest.mark.parametrize("dims,min_n", [([1024, 4, 8], 1024), ([4, 1024, 8], 1024), ([2, 1024, 1024], 1024), ([1024, 1024], 1024),
                                        ([512, 8, 8], 512), ([8, 512, 16], 512), ([512, 512], 512)],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else f"min{v}")
def test_strided_pass_bitwise_and_vs_oracle(tma_env, dims, min_n):
    n = int(np.prod(dims))
    rng = np.random.default_rng(7)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    tma_env["GOPF_TMA_MIN_N"] = str(min_n)
    for sign in (-1, 1):
        tma_env["GOPF_TMA"] = "1"
        gpfutil.TmaLaunchCount(reset=True)
        a = _fft(dims, x, sign)
        assert gpfutil.TmaLaunchCount() > 0, "the copy-engine kernel did not run"
        tma_env["GOPF_TMA"] = "0"
        gpfutil.TmaLaunchCount(reset=True)
        b = _fft(dims, x, sign)
        assert gpfutil.TmaLaunchCount() == 0
        assert np.array_equal(a, b)
        ref = x.copy()
        oft = opfutil.NewFFTW(dims)
        (oft.FFT if sign < 0 else oft.IFFT)(ref)
        assert np.linalg.norm(a - ref) / np.linalg.norm(ref) < 1e-13


def _ch(dims, steps):
    n = int(np.prod(dims))
    m = gpf.NewModel()
    f = gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 0))
    m.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    s = gpf.NewSolver(m, dims, synthetic.CAHN_HILLIARD_DT)
    assert s.IsFused
    s.Upload()
    s.StepDevice(steps)
    s.Download()
    out = f.Data.copy()
    s.close()
    return out


_FUSED_CASES = [(which, dims, min_n, steps)
                for which in ("all", "pass", "real", "kspace")
                for dims, min_n, steps in (([1024, 1024], 1024, 12), ([512, 512], 512, 12), ([512, 512, 512], 512, 3))
                if not (which == "pass" and len(dims) == 2)]  # 2-D has no plain middle pass


@pytest.mark.parametrize("which,dims,min_n,steps", _FUSED_CASES,
                         ids=[f"{w}-{'x'.join(map(str, d))}-{n}-{k}" for w, d, n, k in _FUSED_CASES])
def test_fused_step_bitwise(tma_env, dims, min_n, steps, which):
    """Cahn-Hilliard through the fused kernels with the copy-engine variants switched on one at a time (the real
    field's two-lines-per-transform variant off: it is not bitwise, see test_real_pairs_*)."""
    tma_env["GOPF_REAL_PAIRS"] = "0"
    tma_env["GOPF_TMA_MIN_N"] = str(min_n)
    tma_env["GOPF_TMA"] = "0"
    ref = _ch(dims, steps)
    tma_env["GOPF_TMA"] = "1"
    for k in ("pass", "real", "kspace"):
        tma_env["GOPF_TMA_" + k.upper()] = "1" if which in ("all", k) else "0"
    gpfutil.TmaLaunchCount(reset=True)
    got = _ch(dims, steps)
    assert gpfutil.TmaLaunchCount() > 0
    assert np.array_equal(got, ref)


def _virtual_sharded(n, world, steps):
    """The slab-sharded phases of `world` ranks driven in lock-step on ONE GPU (the all-to-all is a block
    shuffle between the ranks' buffers): exercises the split row maps (pack / unpack folded into the passes
    on either side of the exchange, csrc/dist_solver.h) without a multi-GPU box."""
    import torch
    from gopf_b200 import dist as gdist
    dev = torch.device("cuda", 0)
    cells = n ** 3 // world
    stream = torch.cuda.Stream(device=dev)
    ranks = []
    for r in range(world):
        m = gpf.NewModel()
        f = gpf.NewField("conc", cells, synthetic.cahn_hilliard_initial(cells, 0, offset=r * cells))
        m.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
        m.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
        m.AddField(f)
        m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
        ph = gdist.CudaPhases(m, n, world, r, synthetic.CAHN_HILLIARD_DT, 0)
        ph.set_stream(stream.cuda_stream)
        bufs = [torch.empty(cells, dtype=torch.complex128, device=dev) for _ in range(3)]
        ranks.append((m, f, ph, bufs))

    def a2a(dst_i, src_i):
        blk = cells // world
        for p in range(world):
            for q in range(world):
                ranks[q][3][dst_i][p * blk:(p + 1) * blk].copy_(ranks[p][3][src_i][q * blk:(q + 1) * blk])

    S, A, B = 0, 1, 2
    with torch.cuda.stream(stream):
        for m, f, ph, b in ranks:
            b[A].copy_(torch.from_numpy(f.Data))
            ph.forward_local(b[A], b[B])
        a2a(S, B)
        for m, f, ph, b in ranks:
            ph.forward_finish(b[S])
        for _ in range(steps):
            for m, f, ph, b in ranks:
                ph.inverse_start(b[S], b[A])
            a2a(B, A)
            for m, f, ph, b in ranks:
                ph.inverse_mid(b[B], b[A])
                ph.real_step(b[A])
                ph.forward_mid(b[A], b[B])
            a2a(A, B)
            for m, f, ph, b in ranks:
                ph.kspace_step(b[A], b[S])
                ph.advance()
        for m, f, ph, b in ranks:
            ph.inverse_start(b[S], b[A])
        a2a(B, A)
        out = []
        for m, f, ph, b in ranks:
            ph.inverse_mid(b[B], b[A])
            ph.inverse_finish(b[A], b[A])
            out.append(b[A].cpu().numpy())
    stream.synchronize()
    for m, f, ph, b in ranks:
        ph.close()
    return np.concatenate(out)


@pytest.mark.parametrize("world", [2, 4])
def test_split_row_maps_virtual_ranks_bitwise_and_vs_single_gpu(tma_env, world):
    n, steps = 512, 3
    tma_env["GOPF_TMA_MIN_N"] = "512"
    tma_env["GOPF_REAL_PAIRS"] = "0"
    tma_env["GOPF_TMA"] = "0"
    ref = _virtual_sharded(n, world, steps)
    tma_env["GOPF_TMA"] = "1"
    gpfutil.TmaLaunchCount(reset=True)
    got = _virtual_sharded(n, world, steps)
    assert gpfutil.TmaLaunchCount() > 0
    assert np.array_equal(got, ref)
    single = _ch([n, n, n], steps)
    assert np.linalg.norm(got - single) / np.linalg.norm(single) <= 1e-13
    # the sharded phases with the paired real-space kernel (the slabs are real): equal to rounding
    tma_env["GOPF_REAL_PAIRS"] = "1"
    paired = _virtual_sharded(n, world, steps)
    assert not np.array_equal(paired, got), "the paired kernel did not run in the sharded phases"
    assert np.linalg.norm(paired - got) / np.linalg.norm(got) <= 1e-13


# ---- real fields: two lines per complex transform in the real-space kernel (k_fused_real_pair_tma) -------------------
def _ch_init(dims, steps, init):
    n = int(np.prod(dims))
    m = gpf.NewModel()
    f = gpf.NewField("conc", n, init.copy())
    m.AddScalar(gpf.NewScalar("gamma", synthetic.CAHN_HILLIARD_GAMMA))
    m.AddScalar(gpf.NewScalar("m1", synthetic.CAHN_HILLIARD_M1))
    m.AddField(f)
    m.AddEquation(synthetic.CAHN_HILLIARD_EQUATION)
    s = gpf.NewSolver(m, dims, synthetic.CAHN_HILLIARD_DT)
    assert s.IsFused
    s.Upload()
    s.StepDevice(steps)
    s.Download()
    out = f.Data.copy()
    s.close()
    return out


@pytest.mark.parametrize("dims,min_n,steps", [([1024, 1024], 1024, 20), ([512, 512], 512, 20), ([512, 512, 512], 512, 3),
                                              ([256, 1024], 1024, 8)],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else str(v))
def test_real_pairs_equal_the_complex_carrying_kernel_to_rounding(tma_env, dims, min_n, steps):
    """A real field through the paired real-space kernel against the same run with every line carried as complex data
    (GOPF_REAL_PAIRS=0): the imaginary residue of a line (1e-17 relative) lands in its partner instead of being carried
    along, nothing else differs.  pf/euler.go:16-47; tolerance of the P-GPU vs 1-GPU comparisons (1e-13)."""
    n = int(np.prod(dims))
    init = synthetic.cahn_hilliard_initial(n, 0)
    assert not init.imag.any()
    tma_env["GOPF_TMA"] = "1"
    tma_env["GOPF_TMA_MIN_N"] = str(min_n)
    tma_env["GOPF_REAL_PAIRS"] = "0"
    ref = _ch_init(dims, steps, init)
    tma_env["GOPF_REAL_PAIRS"] = "1"
    got = _ch_init(dims, steps, init)
    assert not np.array_equal(got, ref), "the paired kernel did not run"
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-13
    assert np.max(np.abs(got.imag)) <= 1e-13 * np.max(np.abs(got.real))


def test_real_pairs_are_not_used_for_a_complex_field(tma_env):
    """A field with an imaginary part keeps the complex-carrying kernel: bitwise the GOPF_REAL_PAIRS=0 result."""
    dims = [1024, 1024]
    n = int(np.prod(dims))
    init = synthetic.cahn_hilliard_initial(n, 0)
    init = init + 1j * 0.01 * np.roll(init.real, 17)
    tma_env["GOPF_TMA"] = "1"
    tma_env["GOPF_TMA_MIN_N"] = "1024"
    tma_env["GOPF_REAL_PAIRS"] = "0"
    ref = _ch_init(dims, 6, init)
    tma_env["GOPF_REAL_PAIRS"] = "1"
    got = _ch_init(dims, 6, init)
    assert np.array_equal(got, ref)


def _pfc(dims, steps):
    from gopf_b200 import workloads
    m, f, solver = workloads.build_pfc(gpf, gpf, dims, noise="device", kspace_noise=gpf.HasKSpaceNoise())
    assert solver.IsFused and solver.FusedForm()[0] == 2
    solver.Upload()
    solver.StepDevice(steps)
    solver.Download()
    out = f.Data.copy()
    solver.close()
    return out


@pytest.mark.parametrize("dims", [[512, 512], [1024, 1024]], ids=lambda d: "x".join(map(str, d)))
def test_real_pairs_with_the_tabulated_form_and_kspace_noise(tma_env, dims):
    """cfg 5's model (pair correlation + ideal mixture + Vandeven filter + white noise drawn in k-space) keeps a real
    field real -- real tabulated factor, Hermitian noise spectrum -- so its real-space kernel pairs lines too:
    equal to the complex-carrying run to rounding."""
    tma_env["GOPF_TMA_MIN_N"] = str(dims[-1])
    tma_env["GOPF_REAL_PAIRS"] = "0"
    ref = _pfc(dims, 10)
    tma_env["GOPF_REAL_PAIRS"] = "1"
    gpfutil.TmaLaunchCount(reset=True)
    got = _pfc(dims, 10)
    assert gpfutil.TmaLaunchCount() > 0, "the paired kernel did not run"
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-13
