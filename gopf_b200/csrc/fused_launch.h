// Host entry points of the fused single-field kernels (step_kernels.cuh).  The
// per-length instantiations live in fused_inst_*.cu so they compile in parallel.
#pragma once
#include "step_kernels.cuh"

namespace gopf {

bool fused_length_supported(int n);

// Slowest-axis kernel: finish forward of the derived field in W, Euler update of S,
// first inverse pass of the new S back into W.
cudaError_t launch_fused_kspace(const PassGeom& g, int tx_want, cplx* W, cplx* S, const DevKProgram& P,
                                const FreqTabs& ft, const cplx* tw, cudaStream_t s);

// Contiguous-axis kernel.  mode 0: inverse, /N, derived function, forward (W in place;
// real_out optional).  mode 1: inverse, /N, store to real_out only.
cudaError_t launch_fused_real(const PassGeom& g, int mode, cplx* W, cplx* real_out, const DevDerived& D, double inv_n,
                              unsigned long long step, const cplx* tw, cudaStream_t s);

// ---- templates instantiated by fused_inst_*.cu ------------------------------------------
template <int N, int TX>
cudaError_t fused_kspace_n_tx(const PassGeom& g, cplx* W, cplx* S, const DevKProgram& P, const FreqTabs& ft,
                              const cplx* tw, cudaStream_t s) {
    constexpr int T = PlanFor<N>::T;
    // exchange buffer (multi-stage lengths only) + the prefetched spectrum tile
    const size_t smem = (size_t)N * TX * sizeof(cplx) * 2;
    auto kern = k_fused_kspace<N, TX>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long tiles = g.A * (g.B / TX);
    kern<<<(unsigned)tiles, T * TX, smem, s>>>(g, W, S, P, ft, tw);
    return cudaGetLastError();
}

template <int N>
cudaError_t fused_kspace_n(const PassGeom& g, int tx_want, cplx* W, cplx* S, const DevKProgram& P, const FreqTabs& ft,
                           const cplx* tw, cudaStream_t s) {
    constexpr size_t line_bytes = (size_t)N * sizeof(cplx);
    int tx = pick_tx(N, g.B, tx_want);
    while (tx > 2 && line_bytes * tx * 2 > 128 * 1024) tx >>= 1;  // two tiles per CTA
    switch (tx) {
        case 2: return fused_kspace_n_tx<N, 2>(g, W, S, P, ft, tw, s);
        case 4:
            if constexpr (line_bytes * 4 * 2 <= 200 * 1024) return fused_kspace_n_tx<N, 4>(g, W, S, P, ft, tw, s);
            break;
        case 8:
            if constexpr (line_bytes * 8 * 2 <= 200 * 1024) return fused_kspace_n_tx<N, 8>(g, W, S, P, ft, tw, s);
            break;
        case 16:
            if constexpr (line_bytes * 16 * 2 <= 200 * 1024 && PlanFor<N>::T * 16 <= 1024)
                return fused_kspace_n_tx<N, 16>(g, W, S, P, ft, tw, s);
            break;
        default: break;
    }
    return cudaErrorInvalidConfiguration;
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
