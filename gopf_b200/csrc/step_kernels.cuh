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
__global__ void __launch_bounds__(ContigCfg<N>::T* ContigCfg<N>::LINES, GOPF_MINB(ContigCfg<N>::T* ContigCfg<N>::LINES))
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
    for (int m = 0; m < E; ++m) v[m] = eval_derived_single(D, v[m], step, base + p + T * m);
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
__global__ void __launch_bounds__(PlanFor<N>::T* TX, GOPF_MINB(PlanFor<N>::T* TX))
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
    // Reference Freq components [row, col, depth] = FFTW axes [1, 2, 0] (fftWrap.go:42-74).
    // Two of them are fixed along this thread's line, the third runs with j.
    double fa, fb;          // the two fixed components
    const double* fline;    // table of the running component
    if (g.axis == 0) {
        fa = ft.f1[(int)(b / g.n2)];
        fb = ft.f2[(int)(b % g.n2)];
        fline = ft.f0;
    } else {  // axis 1
        fa = ft.f2[(int)b];
        fb = ft.rank > 2 ? ft.f0[(int)a] : 0.0;
        fline = ft.f1;
    }
    const size_t strideB = (size_t)g.B;
    // The spectrum tile is needed only after the forward transform: start it towards shared
    // memory now (cp.async, no registers held) so its HBM latency hides behind the W loads
    // and the first FFT.  Each thread later reads back exactly the cells it copied, so
    // cp.async.wait_group is the only synchronisation needed.
    cplx* sS = sm + (PlanFor<N>::NS > 1 ? N * TX : 0);
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sS + LayoutInterleaved<TX>::at(t + T * m, l));
        const cplx* src = S + base + (size_t)(t + T * m) * strideB;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    cplx v[E];
    if (DO_FWD) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = W[base + (size_t)(t + T * m) * strideB];
        line_fft<N, LayoutInterleaved<TX>, SyncCta>(v, t, l, sm, tw);
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (DO_FWD) {
        if (P.fast) {
            FastUpdate fu;
            fu.init(P);
            const double s2 = fa * fa + fb * fb;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int j = t + T * m;
                const double fl = fline[j];
                const cplx cur = fu.apply(fma(fl, fl, s2), sS[LayoutInterleaved<TX>::at(j, l)], v[m]);
                S[base + (size_t)j * strideB] = cur;
                v[m] = cswap(cur);
            }
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int j = t + T * m;
                const double fl = fline[j];
                const cplx old = sS[LayoutInterleaved<TX>::at(j, l)];
                const cplx cur = (g.axis == 0) ? euler_update_single_slow(P, fa, fb, fl, old, v[m])  // row, col, depth
                                               : euler_update_single_slow(P, fl, fa, fb, old, v[m]);
                S[base + (size_t)j * strideB] = cur;
                v[m] = cswap(cur);
            }
        }
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cswap(sS[LayoutInterleaved<TX>::at(t + T * m, l)]);
    }
    if (DO_INV) {
        line_fft<N, LayoutInterleaved<TX>, SyncCta>(v, t, l, sm, tw);
#pragma unroll
        for (int m = 0; m < E; ++m) W[base + (size_t)(t + T * m) * strideB] = cswap(v[m]);
    }
}

}  // namespace gopf
