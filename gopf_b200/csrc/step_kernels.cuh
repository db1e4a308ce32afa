// Step-level kernels: the pointwise k-space update, derived-field loaders and the
// fused pass kernels of the single-field fast path.
//
// Fused Euler step for one field c with one nonlinear (derived) field g(c)
// (e.g. Cahn-Hilliard, /root/reference/examples/cahnHilliard/main.go:33), k-space
// state S = c^ persistent on the device, work array W:
//
//   k_fused_kspace (slowest axis) : W = partial forward of g(c)
//        finish forward -> g^(k); S <- (S + dt*rhs)/(1 - dt*den) * filter   [euler.go:28-38]
//        first inverse pass of the NEXT step on the new S -> W              [euler.go:42-45]
//        traffic: read W, read S, write S, write W = 64 B / cell
//   middle axis inverse (3-D only), in place on W                          = 32 B / cell
//   k_fused_real (contiguous axis): last inverse pass, /N -> c(x) in registers,
//        g(c) (model.go:237-241, util.go:57-64), first forward pass -> W   = 32 B / cell
//   middle axis forward (3-D only), in place on W                          = 32 B / cell
//
// 160 B per cell-update in 3-D (96 B in 2-D) against the 192 B (128 B) contract
// model of SURVEY.md 8d, and 288 B for the reference's three separate transforms.
#pragma once
#include "fft_kernels.cuh"
#include "step_program.h"

namespace gopf {

struct SpectraPtrs {
    cplx* s[GOPF_MAX_SPECTRA];  // fields first (k-space state, updated in place), then derived / work spectra
};

struct FreqTabs {
    const double* f0;  // per normalised FFTW axis 0, 1, 2
    const double* f1;
    const double* f2;
    int rank;
};

// Reference Freq components [row, col, depth] from FFTW coordinates (fftWrap.go:42-74;
// consistent layouts only: any 2-D shape, cubic 3-D shapes).
__device__ __forceinline__ KPoint kpoint_at(const FreqTabs& ft, int i0, int i1, int i2) {
    return make_kpoint(ft.f1[i1], ft.f2[i2], ft.rank > 2 ? ft.f0[i0] : 0.0);
}

// ---- fused real-space kernel (contiguous axis) ------------------------------------
// MODE 0: inverse + /N + g + forward (steady state)
// MODE 1: inverse + /N, store real field only (download / generic path helper)
template <int N, int MODE>
__global__ void __launch_bounds__(ContigCfg<N>::T* ContigCfg<N>::LINES)
    k_fused_real(PassGeom g, cplx* __restrict__ W, cplx* __restrict__ real_out, const __grid_constant__ DevDerived D,
                 double inv_n, unsigned long long step, const cplx* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char gopf_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(gopf_smem_raw);
    constexpr int E = PlanFor<N>::E, T = PlanFor<N>::T, LINES = ContigCfg<N>::LINES;
    typedef typename std::conditional<(ContigCfg<N>::WARP_SYNC != 0), SyncWarp, SyncCta>::type Sync;
    const int tid = threadIdx.x;
    const int p = tid % T, l = tid / T;
    long long line = (long long)blockIdx.x * LINES + l;
    const bool live = line < g.A;
    if (!live) line = g.A - 1;
    const size_t base = (size_t)line * N;
    cplx v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = cswap(W[base + p + T * m]);
    line_fft<N, LayoutPadded<N>, Sync>(v, p, l, sm, tw);
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = mk(v[m].y * inv_n, v[m].x * inv_n);  // swap back, /N
    if (MODE == 1) {
        if (live) {
#pragma unroll
            for (int m = 0; m < E; ++m) real_out[base + p + T * m] = v[m];
        }
        return;
    }
    if (real_out != nullptr && live) {
#pragma unroll
        for (int m = 0; m < E; ++m) real_out[base + p + T * m] = v[m];
    }
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const cplx c = v[m];
        v[m] = eval_derived(D, [&](int) -> cplx { return c; }, step, base + p + T * m);
    }
    line_fft<N, LayoutPadded<N>, Sync>(v, p, l, sm, tw);
    if (live) {
#pragma unroll
        for (int m = 0; m < E; ++m) W[base + p + T * m] = v[m];
    }
}

// ---- fused k-space kernel (slowest active axis, strided) ---------------------------
// DO_FWD: W holds the partial forward transform of the derived field; finish it and
//         apply the Euler update to S.  Without DO_FWD, S is used as is.
// DO_INV: start the next inverse transform from the (new) S and leave it in W.
template <int N, int TX, bool DO_FWD, bool DO_INV>
__global__ void __launch_bounds__(PlanFor<N>::T* TX)
    k_fused_kspace(PassGeom g, cplx* __restrict__ W, cplx* __restrict__ S, const __grid_constant__ DevKProgram P,
                   FreqTabs ft, const cplx* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char gopf_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(gopf_smem_raw);
    constexpr int E = PlanFor<N>::E, T = PlanFor<N>::T;
    const int tid = threadIdx.x;
    const int l = tid % TX, t = tid / TX;
    const long long tilesB = g.B / TX;
    const long long tile = blockIdx.x;
    const long long a = tile / tilesB;
    const long long b = (tile - a * tilesB) * TX + l;
    const size_t base = (size_t)a * N * g.B + b;
    // fixed FFTW coordinates of this thread's line
    int c0 = 0, c1 = 0, c2 = 0;
    if (g.axis == 0) { c1 = (int)(b / g.n2); c2 = (int)(b % g.n2); }
    else { c0 = (int)a; c2 = (int)b; }  // axis 1
    cplx v[E];
    if (DO_FWD) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = W[base + (size_t)(t + T * m) * g.B];
        line_fft<N, LayoutInterleaved<TX>, SyncCta>(v, t, l, sm, tw);
    }
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int j = t + T * m;
        const size_t idx = base + (size_t)j * g.B;
        cplx cur = S[idx];
        if (DO_FWD) {
            const KPoint kp = (g.axis == 0) ? kpoint_at(ft, j, c1, c2) : kpoint_at(ft, c0, j, c2);
            const cplx nl = v[m];
            cur = euler_update(P, 0, kp, cur, [&](int bidx) -> cplx { return bidx == 0 ? cur : nl; });
            S[idx] = cur;
        }
        v[m] = cswap(cur);
    }
    if (DO_INV) {
        line_fft<N, LayoutInterleaved<TX>, SyncCta>(v, t, l, sm, tw);
#pragma unroll
        for (int m = 0; m < E; ++m) W[base + (size_t)(t + T * m) * g.B] = cswap(v[m]);
    }
}

}  // namespace gopf
