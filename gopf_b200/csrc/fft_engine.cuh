// Register/shared-memory Stockham FFT engine for one line of N complex128 points.
//
// Replaces the FFTW plan execution behind FFTWWrapper.FFT/IFFT
// (/root/reference/pfutil/fftWrap.go:19-20,28,36).  Forward sign -1, unnormalised;
// the inverse is obtained by swapping re/im on the way in and out
// (IDFT(x) = swap(DFT(swap(x)))), so one set of butterflies and twiddles serves both.
//
// Decomposition N = R0*R1*R2 (each radix <= 16), decimation in frequency,
// autosort (natural order in, natural order out):
//   stage s (radix R, L = product of earlier radices, T = N/E threads per line):
//     butterfly b = t + T*i  (i < E/R) takes positions b + n*N/R, n < R
//     -> DFT_R -> times W_N^{k * (b - b mod L)} -> position (b mod L) + L*k + L*R*(b div L)
//   thread t always holds positions {t + T*m : m < E} in register slot m, both before
//   the first stage and after the last, for every radix.  That invariant is what lets
//   an inverse pass, a pointwise real-space function and the following forward pass
//   along the same axis run back to back in registers (fused kernels in step_kernels.cuh).
#pragma once
#include "cplx.cuh"

namespace gopf {

#define GOPF_SQRT1_2 0.70710678118654752440
#define GOPF_C1_16 0.92387953251128675613  // cos(pi/8)
#define GOPF_S1_16 0.38268343236508977173  // sin(pi/8)

// __launch_bounds__ second argument: aim for 512 resident threads per SM (<= 128 registers
// per thread), the occupancy at which 16 independent 128-bit loads per thread cover HBM latency.
#define GOPF_MINB(threads) ((threads) >= 512 ? 1 : ((512 / (threads)) > 8 ? 8 : (512 / (threads))))

__device__ __forceinline__ cplx cmulc(cplx a, double c, double s) {  // a * (c + i s)
    return mk(fma(a.x, c, -a.y * s), fma(a.x, s, a.y * c));
}

// ---- forward DFT of R points held in registers, natural order in and out -------------
template <int R>
struct Dft;

template <>
struct Dft<1> {
    static __device__ __forceinline__ void run(cplx (&a)[1]) {}
};
template <>
struct Dft<2> {
    static __device__ __forceinline__ void run(cplx (&a)[2]) {
        cplx t = a[0];
        a[0] = t + a[1];
        a[1] = t - a[1];
    }
};
__device__ __forceinline__ void dft4(cplx& a0, cplx& a1, cplx& a2, cplx& a3) {
    cplx t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = mul_mi(a1 - a3);
    a0 = t0 + t2;
    a2 = t0 - t2;
    a1 = t1 + t3;
    a3 = t1 - t3;
}
template <>
struct Dft<4> {
    static __device__ __forceinline__ void run(cplx (&a)[4]) { dft4(a[0], a[1], a[2], a[3]); }
};
template <>
struct Dft<8> {
    static __device__ __forceinline__ void run(cplx (&a)[8]) {
        dft4(a[0], a[2], a[4], a[6]);  // E_k in a[2k]
        dft4(a[1], a[3], a[5], a[7]);  // O_k in a[2k+1]
        cplx o0 = a[1];
        cplx o1 = mk((a[3].x + a[3].y) * GOPF_SQRT1_2, (a[3].y - a[3].x) * GOPF_SQRT1_2);
        cplx o2 = mul_mi(a[5]);
        cplx o3 = mk((a[7].y - a[7].x) * GOPF_SQRT1_2, (-a[7].x - a[7].y) * GOPF_SQRT1_2);
        cplx e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
        a[0] = e0 + o0;
        a[4] = e0 - o0;
        a[1] = e1 + o1;
        a[5] = e1 - o1;
        a[2] = e2 + o2;
        a[6] = e2 - o2;
        a[3] = e3 + o3;
        a[7] = e3 - o3;
    }
};
template <>
struct Dft<16> {
    static __device__ __forceinline__ void run(cplx (&a)[16]) {
        // step 1: y[n2][k1] = DFT4 over n1 of x[4*n1 + n2]; kept in a[4*k1 + n2]
        dft4(a[0], a[4], a[8], a[12]);
        dft4(a[1], a[5], a[9], a[13]);
        dft4(a[2], a[6], a[10], a[14]);
        dft4(a[3], a[7], a[11], a[15]);
        // step 2: times W16^{n2*k1}
        a[5] = cmulc(a[5], GOPF_C1_16, -GOPF_S1_16);                                   // 1
        a[6] = mk((a[6].x + a[6].y) * GOPF_SQRT1_2, (a[6].y - a[6].x) * GOPF_SQRT1_2);  // 2
        a[7] = cmulc(a[7], GOPF_S1_16, -GOPF_C1_16);                                   // 3
        a[9] = mk((a[9].x + a[9].y) * GOPF_SQRT1_2, (a[9].y - a[9].x) * GOPF_SQRT1_2);  // 2
        a[10] = mul_mi(a[10]);                                                         // 4
        a[11] = mk((a[11].y - a[11].x) * GOPF_SQRT1_2, (-a[11].x - a[11].y) * GOPF_SQRT1_2);  // 6
        a[13] = cmulc(a[13], GOPF_S1_16, -GOPF_C1_16);                                 // 3
        a[14] = mk((a[14].y - a[14].x) * GOPF_SQRT1_2, (-a[14].x - a[14].y) * GOPF_SQRT1_2);  // 6
        a[15] = cmulc(a[15], -GOPF_C1_16, GOPF_S1_16);                                 // 9
        // step 3: X[k1 + 4*k2] = DFT4 over n2 of y[n2][k1]; result k2 lands in a[4*k1 + k2]
        dft4(a[0], a[1], a[2], a[3]);
        dft4(a[4], a[5], a[6], a[7]);
        dft4(a[8], a[9], a[10], a[11]);
        dft4(a[12], a[13], a[14], a[15]);
        // transpose 4x4 so that X[k] sits in a[k]
        cplx t;
#define GOPF_SWAP(i, j) t = a[i]; a[i] = a[j]; a[j] = t;
        GOPF_SWAP(1, 4) GOPF_SWAP(2, 8) GOPF_SWAP(3, 12) GOPF_SWAP(6, 9) GOPF_SWAP(7, 13) GOPF_SWAP(11, 14)
#undef GOPF_SWAP
    }
};

// ---- per-length plans ------------------------------------------------------------
template <int N>
struct PlanFor;
#define GOPF_PLAN(N_, E_, NS_, R0_, R1_, R2_)                         \
    template <>                                                       \
    struct PlanFor<N_> {                                              \
        enum { E = E_, NS = NS_, R0 = R0_, R1 = R1_, R2 = R2_, T = N_ / E_ }; \
    };
GOPF_PLAN(2, 2, 1, 2, 1, 1)
GOPF_PLAN(4, 4, 1, 4, 1, 1)
GOPF_PLAN(8, 8, 1, 8, 1, 1)
GOPF_PLAN(16, 16, 1, 16, 1, 1)
GOPF_PLAN(32, 8, 2, 8, 4, 1)
GOPF_PLAN(64, 8, 2, 8, 8, 1)
GOPF_PLAN(128, 16, 2, 16, 8, 1)
GOPF_PLAN(256, 16, 2, 16, 16, 1)
GOPF_PLAN(512, 16, 3, 16, 2, 16)
GOPF_PLAN(1024, 16, 3, 16, 4, 16)
GOPF_PLAN(2048, 16, 3, 16, 8, 16)
GOPF_PLAN(4096, 16, 3, 16, 16, 16)
#undef GOPF_PLAN

template <int N, int S>
struct StageInfo {
    typedef PlanFor<N> P;
    enum {
        R = (S == 0 ? P::R0 : (S == 1 ? P::R1 : P::R2)),
        L = (S == 0 ? 1 : (S == 1 ? P::R0 : P::R0 * P::R1)),
        LAST = (S == P::NS - 1)
    };
};

// ---- shared-memory layouts -------------------------------------------------------
// Strided-axis tiles: TX adjacent lines, line index fastest.  With TX = 8 a quarter-warp (the unit
// of a 128-bit shared-memory access) holds one position and 8 lines = one 128-B row: conflict free
// for every stage without padding.  With TX = 4 (2) it holds W = 2 (4) consecutive positions; the
// gathers are still one row, but the radix-16 scatter of the first stage sends them 16 positions
// apart, onto the same bank groups (2-way / 4-way conflicts: 21 % of the wavefronts of the
// 1024-point pass in ncu).  XOR-ing the position's low bits (which pick the 64-B / 32-B part of
// the row) with bits 4.. of the position keeps every access pattern of every plan on 8 distinct
// bank groups (brute-force check over all plans and stages).  The map is a bijection inside each
// aligned 8-cell group, so the tile size does not change.
template <int TX>
struct LayoutInterleaved {
    enum { W = (TX >= 8 ? 1 : 8 / TX) };
    static __device__ __forceinline__ int at(int pos, int l) {
        const int a = pos * TX + l;
        return W > 1 ? (a ^ (((pos >> 4) & (W - 1)) * TX)) : a;
    }
    static constexpr int elems(int n, int lines) { return n * lines; }
};
// Contiguous-axis lines: position fastest, one pad cell per 16 so that the radix-16
// scatter (stride 16 cells = 256 B) spreads over all bank groups.
template <int N>
struct LayoutPadded {
    enum { LS = N + N / 16 };
    static __device__ __forceinline__ int at(int pos, int l) { return l * LS + pos + (pos >> 4); }
    static constexpr int elems(int n, int lines) { return (n + n / 16) * lines; }
};

struct SyncCta {
    static __device__ __forceinline__ void run() { __syncthreads(); }
};
struct SyncWarp {
    static __device__ __forceinline__ void run() { __syncwarp(); }
};

// a[k] *= W_N^{k*bh}, k = 1 .. R-1.  Only the power-of-two multiples W^{bh}, W^{2bh}, W^{4bh},
// W^{8bh} come from the table (4 loads instead of 15 at radix 16); the others are products.  The
// passes are limited by the load/store queues, not by the fp64 pipe (ncu: fp64 pipe 24 % busy,
// lg_throttle + mio_throttle stalls dominate at 1024-point lines), so trading loads for DFMAs pays.
// Each product costs about 1.5 ulp of the unit-modulus twiddle, far inside the 1e-10 tolerance.
// Where the twiddle table lives.  TwGlobal: global memory through the read-only path (L1-resident when the
// kernel leaves L1 some room).  TwShared: a copy in shared memory for the kernels whose tile buffers take the
// whole unified L1 / shared array (tma_kernels.cuh: with 209 KB of buffers the table fell out of L1 and every
// twiddle became an L2 round trip, long_scoreboard 12 cycles per issue).  Entry j sits at j + (j >> 3), so the
// four loads of a radix-16 butterfly (strides 1, 2, 4, 8 over eight consecutive threads) are bank-conflict free.
struct TwGlobal {
    static __device__ __forceinline__ cplx ld(const cplx* __restrict__ tw, int j) { return ld_tab(tw + j); }
};
struct TwShared {
    static constexpr int elems(int n) { return n + n / 8; }
    static __device__ __forceinline__ int at(int j) { return j + (j >> 3); }
    static __device__ __forceinline__ cplx ld(const cplx* tw, int j) { return tw[at(j)]; }
};

template <int R, class Tw>
__device__ __forceinline__ void twiddle_mul(cplx (&a)[R], const cplx* __restrict__ tw, int bh) {
    if (R == 2) {
        a[1] = a[1] * Tw::ld(tw, bh);
    } else if (R == 4) {
        const cplx w1 = Tw::ld(tw, bh), w2 = Tw::ld(tw, 2 * bh);
        a[1] = a[1] * w1;
        a[2] = a[2] * w2;
        a[3] = a[3] * (w1 * w2);
    } else if (R == 8) {
        const cplx w1 = Tw::ld(tw, bh), w2 = Tw::ld(tw, 2 * bh), w4 = Tw::ld(tw, 4 * bh);
        const cplx w3 = w1 * w2;
        a[1] = a[1] * w1;
        a[2] = a[2] * w2;
        a[3] = a[3] * w3;
        a[4] = a[4] * w4;
        a[5] = a[5] * (w4 * w1);
        a[6] = a[6] * (w4 * w2);
        a[7] = a[7] * (w4 * w3);
    } else {  // 16
        const cplx w1 = Tw::ld(tw, bh), w2 = Tw::ld(tw, 2 * bh), w4 = Tw::ld(tw, 4 * bh), w8 = Tw::ld(tw, 8 * bh);
        const cplx w3 = w1 * w2;
        a[1] = a[1] * w1;
        a[2] = a[2] * w2;
        a[3] = a[3] * w3;
        a[4] = a[4] * w4;
        a[8] = a[8] * w8;
        a[12] = a[12] * (w8 * w4);
        {
            const cplx w5 = w4 * w1;
            a[5] = a[5] * w5;
            a[13] = a[13] * (w8 * w5);
        }
        {
            const cplx w6 = w4 * w2;
            a[6] = a[6] * w6;
            a[14] = a[14] * (w8 * w6);
        }
        {
            const cplx w7 = w4 * w3;
            a[7] = a[7] * w7;
            a[15] = a[15] * (w8 * w7);
        }
        a[9] = a[9] * (w8 * w1);
        a[10] = a[10] * (w8 * w2);
        a[11] = a[11] * (w8 * w3);
    }
}

// ---- one stage -------------------------------------------------------------------
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};

// `pre` runs once, after the first butterfly of the stage and before its first exchange write: a kernel whose
// exchange buffer is still busy (a bulk store reading it, other lines still reading the landed tile) puts the
// wait there, behind the register-only arithmetic, instead of in front of the transform (tma_kernels.cuh).
template <int N, int S, class Layout, class Sync, class Tw = TwGlobal, class Pre = NoHook>
__device__ __forceinline__ void fft_stage(cplx (&v)[PlanFor<N>::E], int t, int l, cplx* sm,
                                          const cplx* __restrict__ tw, Pre pre = Pre()) {
    typedef PlanFor<N> P;
    typedef StageInfo<N, S> SI;
    constexpr int E = P::E, T = P::T, R = SI::R, L = SI::L, Q = E / R;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        cplx a[R];
#pragma unroll
        for (int n = 0; n < R; ++n) a[n] = v[i + n * Q];
        Dft<R>::run(a);
        if (SI::LAST) {
#pragma unroll
            for (int k = 0; k < R; ++k) v[i + k * Q] = a[k];
        } else {
            const int b = t + T * i;
            const int bl = b & (L - 1);
            const int bh = b - bl;  // (b div L) * L
            twiddle_mul<R, Tw>(a, tw, bh);
            if (i == 0) pre();
            const int base = bl + (bh * R);
#pragma unroll
            for (int k = 0; k < R; ++k) sm[Layout::at(base + L * k, l)] = a[k];
        }
    }
    if (!SI::LAST) {
        Sync::run();
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = sm[Layout::at(t + T * m, l)];
        Sync::run();
    }
}

// Forward DFT of one line.  v[m] <-> position t + T*m on entry and on exit.
// sm: the CTA's exchange buffer (unused when NS == 1).  tw: W_N^j = exp(-2 pi i j / N).
template <int N, class Layout, class Sync, class Tw = TwGlobal>
__device__ __forceinline__ void line_fft(cplx (&v)[PlanFor<N>::E], int t, int l, cplx* sm,
                                         const cplx* __restrict__ tw) {
    typedef PlanFor<N> P;
    fft_stage<N, 0, Layout, Sync, Tw>(v, t, l, sm, tw);
    if (P::NS > 1) fft_stage<N, (P::NS > 1 ? 1 : 0), Layout, Sync, Tw>(v, t, l, sm, tw);
    if (P::NS > 2) fft_stage<N, (P::NS > 2 ? 2 : 0), Layout, Sync, Tw>(v, t, l, sm, tw);
}

// Stages S0 .. NS-1 of the same transform (S0 = 0: all of it).  Lets a kernel do something between two
// stages (tma_kernels.cuh takes the shared spectrum buffer after the first exchange).
template <int N, int S0, class Layout, class Sync>
__device__ __forceinline__ void line_fft_from(cplx (&v)[PlanFor<N>::E], int t, int l, cplx* sm,
                                              const cplx* __restrict__ tw) {
    typedef PlanFor<N> P;
    if (S0 <= 0) fft_stage<N, 0, Layout, Sync>(v, t, l, sm, tw);
    if (S0 <= 1 && P::NS > 1) fft_stage<N, (P::NS > 1 ? 1 : 0), Layout, Sync>(v, t, l, sm, tw);
    if (S0 <= 2 && P::NS > 2) fft_stage<N, (P::NS > 2 ? 2 : 0), Layout, Sync>(v, t, l, sm, tw);
}

// The same transform in two parts, for kernels that want the exchange buffer back early:
// after line_fft_head every thread has passed the last barrier of the last exchange, so `sm`
// is free while line_fft_tail (register-only butterflies of the last stage) runs.
template <int N, class Layout, class Sync, class Tw = TwGlobal, class Pre = NoHook>
__device__ __forceinline__ void line_fft_head(cplx (&v)[PlanFor<N>::E], int t, int l, cplx* sm,
                                              const cplx* __restrict__ tw, Pre pre = Pre()) {
    typedef PlanFor<N> P;
    if (P::NS > 1) fft_stage<N, 0, Layout, Sync, Tw, Pre>(v, t, l, sm, tw, pre);
    if (P::NS > 2) fft_stage<N, (P::NS > 2 ? 1 : 0), Layout, Sync, Tw>(v, t, l, sm, tw);
}
template <int N, class Layout, class Sync, class Tw = TwGlobal>
__device__ __forceinline__ void line_fft_tail(cplx (&v)[PlanFor<N>::E], int t, int l, cplx* sm,
                                              const cplx* __restrict__ tw) {
    fft_stage<N, PlanFor<N>::NS - 1, Layout, Sync, Tw>(v, t, l, sm, tw);
}
// whole transform with the hook of its first stage (NS >= 2)
template <int N, class Layout, class Sync, class Tw, class Pre>
__device__ __forceinline__ void line_fft_pre(cplx (&v)[PlanFor<N>::E], int t, int l, cplx* sm,
                                             const cplx* __restrict__ tw, Pre pre) {
    static_assert(PlanFor<N>::NS >= 2, "line_fft_pre: single-stage plans never touch the exchange buffer");
    line_fft_head<N, Layout, Sync, Tw, Pre>(v, t, l, sm, tw, pre);
    line_fft_tail<N, Layout, Sync, Tw>(v, t, l, sm, tw);
}

}  // namespace gopf
