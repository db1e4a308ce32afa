// Fused pass kernels of the single-field fast path.
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
// Freq of position j = t + T*m along a line of power-of-two length N (every fused length is one): wrap(j / N) of
// FFTWWrapper.Freq (pfutil/fftWrap.go:57-74) is exact in binary floating point, so it is computed instead of read
// from the axis table -- one DADD per cell instead of a dependent global load (the k-space kernels sat on
// long_scoreboard stalls for these: 5.2 per issue at 256^3, profiles/r1c_ncu_full_step.md).  For unrolled m the
// wrap test is a compile-time constant except at j = N/2 + t.
template <int N, int T>
struct LineFreq {
    double f0;
    int t;
    __device__ __forceinline__ explicit LineFreq(int t_) : f0((double)t_ * (1.0 / N)), t(t_) {}
    __device__ __forceinline__ double at(int m) const {
        const double f = f0 + (double)(T * m) * (1.0 / N);
        const bool wrap = T * m > N / 2 ? true : (T * m + (T - 1) <= N / 2 ? false : t + T * m > N / 2);
        return wrap ? f - 1.0 : f;
    }
};


struct FreqTabs {
    const double* f0;  // per normalised FFTW axis 0, 1, 2: wrap(idx / n), fftWrap.go:57-74
    const double* f1;
    const double* f2;
    int rank;
    int off1;  // slab-sharded k-space: this rank's first k1 (0 on a single GPU)
};

// 16-byte asynchronous global -> shared copies (cp.async.cg).  tests/host_emul (GOPF_HOST_EMUL) runs these
// kernels on the host, where the "shared address" is an offset into the block's buffer and the copy is immediate.
#ifdef GOPF_HOST_EMUL
inline unsigned smem_address(const void* p) { return (unsigned)(reinterpret_cast<const unsigned char*>(p) - gopf_smem_raw); }
inline void cp_async16(unsigned dst, const void* src) { memcpy(gopf_smem_raw + dst, src, 16); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
#else
__device__ __forceinline__ unsigned smem_address(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
#endif

// ---- fused real-space kernel (contiguous axis) ------------------------------------
// MODE 0: inverse + /N + g + forward (steady state)
// MODE 1: inverse + /N, store real field only
template <int N, int MODE>
__global__ void __launch_bounds__(ContigCfg<N>::T* ContigCfg<N>::LINES, GOPF_MINB(ContigCfg<N>::T* ContigCfg<N>::LINES))
    k_fused_real(PassGeom g, cplx* __restrict__ W, cplx* __restrict__ real_out, const __grid_constant__ DevDerived D,
                 double inv_n, unsigned long long step, const cplx* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char gopf_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(gopf_smem_raw);
    constexpr int E = PlanFor<N>::E, T = PlanFor<N>::T, LINES = ContigCfg<N>::LINES;
    typedef typename std::conditional<(ContigCfg<N>::WARP_SYNC != 0), SyncWarp, SyncCta>::type Sync;
    typedef LayoutPadded<N> Lay;
    const int tid = threadIdx.x;
    const int p = tid % T, l = tid / T;
    long long line = (long long)blockIdx.x * LINES + l;
    const bool live = line < g.A;
    if (!live) line = g.A - 1;
    const size_t base = (size_t)line * N;
    cplx v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = W[base + p + T * m];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
    line_fft<N, Lay, Sync>(v, p, l, sm, tw);
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
    if (derived_is_monomial_fast(D)) {
        const int pw = D.ipower[0];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = derived_fast(pw, v[m]);
    } else if (derived_is_poly_fast(D)) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = derived_poly(D, v[m]);
    } else {
        // general derived field: cells staged in shared memory, interpreter in a rolled loop
#pragma unroll
        for (int m = 0; m < E; ++m) sm[Lay::at(p + T * m, l)] = v[m];
#pragma unroll 1
        for (int m = 0; m < E; ++m) {
            const int pos = Lay::at(p + T * m, l);
            const cplx c = sm[pos];
            sm[pos] = eval_derived(D, [&](int) -> cplx { return c; }, step, (unsigned long long)g.node0 + base + p + T * m);
        }
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = sm[Lay::at(p + T * m, l)];
        Sync::run();  // other threads' next exchange writes must not land on cells still being read
    }
    line_fft<N, Lay, Sync>(v, p, l, sm, tw);
    if (live) {
#pragma unroll
        for (int m = 0; m < E; ++m) W[base + p + T * m] = v[m];
    }
}

// ---- fused k-space kernel (slowest active axis, strided) ---------------------------
// W holds the partial forward transform of the derived field: finish it, apply the Euler
// update to S, start the next inverse transform from the new S and leave it in W.
// Shared memory: [exchange tile N*TX][spectrum tile N*TX].
// resident CTAs are bounded by the two shared-memory tiles; do not squeeze registers below that
#define GOPF_MINB_K(threads, smem) \
    (GOPF_MINB(threads) < (200 * 1024 / (smem)) ? GOPF_MINB(threads) : ((200 * 1024 / (smem)) < 1 ? 1 : (200 * 1024 / (smem))))
// PEER: the inverse pass of the new spectrum is stored straight into the owning ranks'
// receive buffers (g.peer, fft_kernels.cuh) instead of W.
// LATE = false: the spectrum tile is prefetched into a second shared-memory tile at kernel start
//   (its latency hides behind the W loads and the whole forward FFT); 2 tiles of shared memory.
// LATE = true: the spectrum tile lands in the exchange tile itself once the forward FFT has made
//   its last exchange (its latency hides behind the last register-only butterfly stage and the
//   other resident CTAs); 1 tile of shared memory, so long lines keep wide tiles / two CTAs per SM.
// SPLIT: W / S / Wout are addressed through the row maps of the launch (blocked k-space layout, solver.cu);
//   otherwise they are plain [A][N][B] arrays and a row is one multiply away (the maps cost ~10 % more
//   instructions, measured 5073 -> 4643 GB/s at 256^3, so the uniform case keeps its own instantiation).
// TAB: the tabulated single-field form (DevKProgram::fast == 2) instead of the polynomial one; its own
//   instantiation because the noise generator and the extra polynomials cost the plain kernels registers
//   (256-length kernel: 5073 -> 3730 GB/s when both forms shared one kernel).
#define GOPF_KMODE_PLAIN 0
#define GOPF_KMODE_SPLIT 1
#define GOPF_KMODE_TAB 2
template <int N, int TX, bool PEER, bool LATE, int MODE = GOPF_KMODE_PLAIN>
__global__ void __launch_bounds__(PlanFor<N>::T* TX, GOPF_MINB_K(PlanFor<N>::T* TX, (LATE ? 1 : 2) * N * TX * 16))
    k_fused_kspace(const __grid_constant__ PassGeom g, const cplx* W, cplx* Wout, cplx* __restrict__ S,
                   const __grid_constant__ DevKProgram P, FreqTabs ft, const cplx* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char gopf_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(gopf_smem_raw);
    cplx* sS = LATE ? sm : sm + N * TX;
    constexpr int E = PlanFor<N>::E, T = PlanFor<N>::T;
    constexpr bool SPLIT = MODE == GOPF_KMODE_SPLIT, TAB = MODE == GOPF_KMODE_TAB;
    typedef LayoutInterleaved<TX> Lay;
    const int tid = threadIdx.x;
    const int l = tid % TX, t = tid / TX;
    const long long tilesB = g.bcount / TX;
    const long long tile = blockIdx.x;
    const long long a = tile / tilesB;
    const long long b = window_col(g, tile - a * tilesB, TX) + l;
    // W and S share the input row map, Wout the output one (uniform [A][N][B] unless the launch says otherwise:
    // blocked k-space layout, solver.cu)
    const size_t base = SPLIT ? (size_t)slab_off(g.in, a) + b : (size_t)a * N * g.B + b;
    const size_t obase = SPLIT ? (size_t)slab_off(g.out, a) + b : base;
    const size_t strideB = (size_t)g.B;
    auto roff = [&](int j) -> size_t { return SPLIT ? (size_t)row_off(g.in, j) : (size_t)j * strideB; };
    auto roff_out = [&](int j) -> size_t { return SPLIT ? (size_t)row_off(g.out, j) : (size_t)j * strideB; };
    // Each thread later reads back exactly the spectrum cells it copied, so cp.async.wait_group
    // is the only synchronisation the copy needs.
    auto prefetch_spectrum = [&]() {
        if (LATE) {
            // the line is live in registers here: walk the rows with two running addresses in a
            // rolled loop instead of materialising E address pairs
            const unsigned sbase = smem_address(sS);
            if (SPLIT) {
                const cplx* src = S + base;
#pragma unroll 1
                for (int m = 0; m < E; ++m) {
                    const unsigned dst = sbase + (unsigned)(Lay::at(t + T * m, l) * (int)sizeof(cplx));
                    cp_async16(dst, src + roff(t + T * m));
                }
            } else {
                const cplx* src = S + base + (size_t)t * strideB;
                const size_t src_step = (size_t)T * strideB;
#pragma unroll 1
                for (int m = 0; m < E; ++m) {
                    const unsigned dst = sbase + (unsigned)(Lay::at(t + T * m, l) * (int)sizeof(cplx));
                    cp_async16(dst, src);
                    src += src_step;
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const unsigned dst = smem_address(sS + Lay::at(t + T * m, l));
                const cplx* src = S + base + roff(t + T * m);
                cp_async16(dst, src);
            }
        }
        cp_async_commit();
    };
    if (!LATE) prefetch_spectrum();
    if (TAB) {
        // the tabulated factors of this tile into L2 now: they are read after both the forward transform and the
        // noise loop, and a DRAM round trip there is exposed (ncu, 512^3 cfg 5: a third of the kernel's stall
        // samples sat on the multiplies that consume them)
        const double* __restrict__ dtab = P.dtab;
#pragma unroll 4
        for (int m = 0; m < E; ++m) prefetch_l2(dtab + base + roff(t + T * m));
    }
    cplx v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = W[base + roff(t + T * m)];

    // Reference Freq components [row, col, depth] = FFTW axes [1, 2, 0] (fftWrap.go:42-74).
    // Two of them are fixed along this thread's line, the third runs with j (LineFreq).
    double fa, fb;        // the two fixed components
    if (g.axis == 0) {
        fa = ft.f1[ft.off1 + (int)(b / g.n2)];
        fb = ft.f2[(int)(b % g.n2)];
    } else if (g.axis == GOPF_AXIS0_BY_PLANE) {  // lines along axis 0, slab = axis-1 index, column = axis-2 index
        fa = ft.f1[ft.off1 + (int)a];
        fb = ft.f2[(int)b];
    } else {  // axis 1
        fa = ft.f2[(int)b];
        fb = ft.rank > 2 ? ft.f0[(int)a] : 0.0;
    }
    const LineFreq<N, T> lf(t);
    if (LATE) {
        line_fft_head<N, Lay, SyncCta>(v, t, l, sm, tw);
        prefetch_spectrum();  // the exchange tile is free from here until the next transform's first exchange
        line_fft_tail<N, Lay, SyncCta>(v, t, l, sm, tw);
    } else {
        line_fft<N, Lay, SyncCta>(v, t, l, sm, tw);
    }
    cp_async_wait_all();
    if (LATE || P.fast) {  // LATE is launched for fast-form programs only (fused_launch.h)
        const double s2 = fa * fa + fb * fb;
        if (TAB) {
            // Tabulated form: one real per k-point read next to the spectrum cell (8 more bytes per cell).  The
            // noise term is folded into the staged spectrum cell first, in a ROLLED loop: sixteen inlined copies
            // of the Philox / Box-Muller chain next to the register-resident line spill (532 bytes measured).
            const double* __restrict__ dtab = P.dtab;
#ifdef GOPF_KNOISE
            if (P.noise_param >= 0) {
                const KnComp ca = knoise_comp(fa), cb = knoise_comp(fb);  // fixed along the line
#pragma unroll 1
                for (int m = 0; m < E; ++m) {
                    const int j = t + T * m;
                    const double fl = lf.at(m);
                    const KnComp cl = knoise_comp_index<N>(j > N / 2 ? j - N : j);
                    const int pos = Lay::at(j, l);
                    sS[pos] = (g.axis != 1) ? tab_self_and_noise(P, fma(fl, fl, s2), ca, cb, cl, sS[pos])
                                            : tab_self_and_noise(P, fma(fl, fl, s2), cl, ca, cb, sS[pos]);
                }
            }
#endif
            const bool folded = P.noise_param >= 0;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int j = t + T * m;
                const double fl = lf.at(m);
                const double dk = dtab[base + roff(j)];
                const cplx cur = tab_update(P, fma(fl, fl, s2), dk, sS[Lay::at(j, l)], v[m], folded);
                S[base + roff(j)] = cur;
                v[m] = cswap(cur);
            }
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int j = t + T * m;
                const double fl = lf.at(m);
                const cplx cur = fast_update(P, fma(fl, fl, s2), sS[Lay::at(j, l)], v[m]);
                S[base + roff(j)] = cur;
                v[m] = cswap(cur);
            }
        }
        if (LATE) __syncthreads();  // spectrum cells were read from the exchange tile
    } else {
        // general program: cells staged in shared memory, term interpreter in a rolled loop
#pragma unroll
        for (int m = 0; m < E; ++m) sm[Lay::at(t + T * m, l)] = v[m];
#pragma unroll 1
        for (int m = 0; m < E; ++m) {
            const int j = t + T * m;
            const int pos = Lay::at(j, l);
            const double fl = lf.at(m);
            const KPoint kp = (g.axis != 1) ? make_kpoint(fa, fb, fl) : make_kpoint(fl, fa, fb);  // row, col, depth
            const cplx old = sS[pos], nl = sm[pos];
            const cplx cur = euler_update(P, 0, kp, old, [&](int bi) -> cplx { return bi == 0 ? old : nl; });
            S[base + roff(j)] = cur;
            sm[pos] = cswap(cur);
        }
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = sm[Lay::at(t + T * m, l)];
        __syncthreads();
    }
    line_fft<N, Lay, SyncCta>(v, t, l, sm, tw);
    if (PEER) {
#pragma unroll
        for (int m = 0; m < E; ++m) *peer_row(g.peer, a, t + T * m, b) = cswap(v[m]);
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) Wout[obase + roff_out(t + T * m)] = cswap(v[m]);
    }
}

}  // namespace gopf
