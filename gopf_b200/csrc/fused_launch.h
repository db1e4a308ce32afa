// Host entry points of the fused single-field kernels (step_kernels.cuh).  The
// per-length instantiations live in fused_inst_*.cu so they compile in parallel.
#pragma once
#include "step_kernels.cuh"

namespace gopf {

bool fused_length_supported(int n);

// Slowest-axis kernel: finish forward of the derived field in W, Euler update of S,
// first inverse pass of the new S back into W.
cudaError_t launch_fused_kspace(const PassGeom& g, int tx_want, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P,
                                const FreqTabs& ft, const cplx* tw, cudaStream_t s);

// Contiguous-axis kernel.  mode 0: inverse, /N, derived function, forward (W in place;
// real_out optional).  mode 1: inverse, /N, store to real_out only.
cudaError_t launch_fused_real(const PassGeom& g, int mode, cplx* W, cplx* real_out, const DevDerived& D, double inv_n,
                              unsigned long long step, const cplx* tw, cudaStream_t s);

// ---- templates instantiated by fused_inst_*.cu ------------------------------------------
template <int N, int TX, bool LATE>
cudaError_t fused_kspace_n_tx(const PassGeom& g, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P,
                              const FreqTabs& ft, const cplx* tw, cudaStream_t s) {
    constexpr int T = PlanFor<N>::T;
    const bool tab = P.fast == 2;
    // exchange tile, plus (early prefetch only) the spectrum tile
    const size_t smem = (size_t)N * TX * sizeof(cplx) * (LATE ? 1 : 2);
    const bool split = g.in.split_log < 31 || g.in.a_split_log < 31 || g.out.split_log < 31 || g.out.a_split_log < 31 ||
                       g.axis == GOPF_AXIS0_BY_PLANE;
    if ((split || tab) && g.peer.n > 0) return cudaErrorNotSupported;
    if (split && tab) return cudaErrorNotSupported;
    if (tab && !LATE) return cudaErrorNotSupported;
    auto kern = split ? k_fused_kspace<N, TX, false, LATE, GOPF_KMODE_SPLIT>
                : tab ? k_fused_kspace<N, TX, false, true, GOPF_KMODE_TAB>
                      : (g.peer.n > 0 ? k_fused_kspace<N, TX, true, LATE, GOPF_KMODE_PLAIN>
                                      : k_fused_kspace<N, TX, false, LATE, GOPF_KMODE_PLAIN>);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long tiles = g.A * (g.bcount / TX);
    kern<<<(unsigned)tiles, T * TX, smem, s>>>(g, W, Wout, S, P, ft, tw);
    return cudaGetLastError();
}

int env_int(const char* name, int fallback);  // fused_launch.cu

template <int N, bool LATE>
cudaError_t fused_kspace_n_late(const PassGeom& g, int tx, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P,
                                const FreqTabs& ft, const cplx* tw, cudaStream_t s) {
    constexpr size_t tile_bytes = (size_t)N * sizeof(cplx) * (LATE ? 1 : 2);
    switch (tx) {
        case 2: return fused_kspace_n_tx<N, 2, LATE>(g, W, Wout, S, P, ft, tw, s);
        case 4:
            if constexpr (tile_bytes * 4 <= 200 * 1024) return fused_kspace_n_tx<N, 4, LATE>(g, W, Wout, S, P, ft, tw, s);
            break;
        case 8:
            if constexpr (tile_bytes * 8 <= 200 * 1024) return fused_kspace_n_tx<N, 8, LATE>(g, W, Wout, S, P, ft, tw, s);
            break;
        case 16:
            if constexpr (tile_bytes * 16 <= 200 * 1024 && PlanFor<N>::T * 16 <= 1024)
                return fused_kspace_n_tx<N, 16, LATE>(g, W, Wout, S, P, ft, tw, s);
            break;
        default: break;
    }
    return cudaErrorInvalidConfiguration;
}

// Tile width and spectrum-staging variant.  Fast-form programs take the single-tile (LATE)
// kernel with the widest tile that keeps two CTAs per SM (64 KB), and never less than 8 cells:
// row segments of 128 B or more are what HBM sectors and NVLink packets want (peer stores:
// 128-B segments 717 GB/s, 64-B 438 GB/s, 32-B 219 GB/s; scripts/peer_store_probe.cu).
// Measured on B200, fused k-space kernel, early two-tile variant -> LATE (scripts/tune_kspace.py):
//   256^3: 4082 -> 5258 GB/s (TX 16);  512^3: 2963 -> 4268 GB/s (TX 8);  1024^3: 2780 -> 3787 GB/s (TX 8).
// General programs keep the two-tile kernel (the interpreter stages cells in the exchange tile).
// GOPF_KSPACE_LATE / GOPF_KSPACE_TX override the choice (tuning).
template <int N>
cudaError_t fused_kspace_n(const PassGeom& g, int tx_want, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P,
                           const FreqTabs& ft, const cplx* tw, cudaStream_t s) {
    constexpr size_t line_bytes = (size_t)N * sizeof(cplx);
    constexpr int T = PlanFor<N>::T;
    bool late = P.fast != 0;
    const int env_late = env_int("GOPF_KSPACE_LATE", -1);
    if (env_late >= 0) late = P.fast && (env_late != 0 || P.fast == 2);
    int tx;
    if (late) {
        tx = 16;
        while (tx > 8 && line_bytes * tx > 64 * 1024) tx >>= 1;
        while (tx > 2 && (line_bytes * tx > 128 * 1024 || T * tx > 1024 || (g.bw % tx) != 0)) tx >>= 1;
    } else {
        tx = pick_tx(N, g.bcount, tx_want);
        while (tx > 2 && (line_bytes * tx * 2 > 128 * 1024 || (g.bw % tx) != 0)) tx >>= 1;
    }
    const int env_tx = env_int("GOPF_KSPACE_TX", 0);
    if (env_tx > 0 && (g.bw % env_tx) == 0) tx = env_tx;
    return late ? fused_kspace_n_late<N, true>(g, tx, W, Wout, S, P, ft, tw, s)
                : fused_kspace_n_late<N, false>(g, tx, W, Wout, S, P, ft, tw, s);
}

template <int N, int MODE>
cudaError_t fused_real_n_mode(const PassGeom& g, cplx* W, cplx* real_out, const DevDerived& D, double inv_n,
                              unsigned long long step, const cplx* tw, cudaStream_t s) {
    constexpr int T = ContigCfg<N>::T, LINES = ContigCfg<N>::LINES;
    const size_t smem = (size_t)LayoutPadded<N>::elems(N, LINES) * sizeof(cplx);
    auto kern = k_fused_real<N, MODE>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long blocks = (g.A + LINES - 1) / LINES;
    kern<<<(unsigned)blocks, T * LINES, smem, s>>>(g, W, real_out, D, inv_n, step, tw);
    return cudaGetLastError();
}

template <int N>
cudaError_t fused_real_n(const PassGeom& g, int mode, cplx* W, cplx* real_out, const DevDerived& D, double inv_n,
                         unsigned long long step, const cplx* tw, cudaStream_t s) {
    return mode == 0 ? fused_real_n_mode<N, 0>(g, W, real_out, D, inv_n, step, tw, s)
                     : fused_real_n_mode<N, 1>(g, W, real_out, D, inv_n, step, tw, s);
}

}  // namespace gopf
