// Axis-pass kernels: one HBM read + one HBM write of every 16-byte cell per pass
// (B_T = 32*rank bytes per cell per transform).  Pointwise work is
// attached through load/store functors so it costs no extra HBM traffic.
//
// Array view for a pass along `axis` of a row-major [n0][n1][n2] array:
//   [A][N][B], element (a, j, b) at (a*N + j)*B + b, N = n[axis],
//   B = product of faster extents (inner stride), A = product of slower extents.
#pragma once
#include "fft_engine.cuh"

namespace gopf {

struct PassGeom {
    int n0, n1, n2;  // extents in FFTW order (2-D: n0 == 1; 1-D: n0 == n1 == 1)
    int axis;        // 0, 1 or 2
    long long A, B;  // outer count / inner stride for this axis
    int N;           // n[axis]
};

inline PassGeom make_geom(int n0, int n1, int n2, int axis) {
    PassGeom g;
    g.n0 = n0; g.n1 = n1; g.n2 = n2; g.axis = axis;
    if (axis == 2) { g.N = n2; g.B = 1; g.A = (long long)n0 * n1; }
    else if (axis == 1) { g.N = n1; g.B = n2; g.A = n0; }
    else { g.N = n0; g.B = (long long)n1 * n2; g.A = 1; }
    return g;
}

// ---- functors ------------------------------------------------------------------
struct LoadPlain {
    const cplx* __restrict__ in;
    __device__ __forceinline__ void begin(const PassGeom&, long long, long long) {}
    __device__ __forceinline__ cplx operator()(size_t idx, int) const { return in[idx]; }
};
struct StorePlain {
    cplx* __restrict__ out;
    double scale;
    __device__ __forceinline__ void begin(const PassGeom&, long long, long long) {}
    __device__ __forceinline__ void operator()(size_t idx, int, cplx v) const { out[idx] = mk(v.x * scale, v.y * scale); }
};

// ---- strided axis (B > 1): tile = N x TX, TX adjacent lines ---------------------
template <int N, int TX, bool INV, class Ld, class St>
__global__ void __launch_bounds__(PlanFor<N>::T* TX)
    k_pass_strided(PassGeom g, Ld ld, St st, const cplx* __restrict__ tw) {
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
    ld.begin(g, a, b);
    st.begin(g, a, b);
    cplx v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int j = t + T * m;
        cplx x = ld(base + (size_t)j * g.B, j);
        v[m] = INV ? cswap(x) : x;
    }
    line_fft<N, LayoutInterleaved<TX>, SyncCta>(v, t, l, sm, tw);
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int j = t + T * m;
        st(base + (size_t)j * g.B, j, INV ? cswap(v[m]) : v[m]);
    }
}

// ---- contiguous axis (B == 1): LINES lines per CTA, position fastest ------------
template <int N>
struct ContigCfg {
    enum {
        T = PlanFor<N>::T,
        LINES = (T >= 128 ? 1 : 128 / T),
        WARP_SYNC = (T <= 32)
    };
};

template <int N, bool INV, class Ld, class St>
__global__ void __launch_bounds__(ContigCfg<N>::T* ContigCfg<N>::LINES)
    k_pass_contig(PassGeom g, Ld ld, St st, const cplx* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char gopf_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(gopf_smem_raw);
    constexpr int E = PlanFor<N>::E, T = PlanFor<N>::T, LINES = ContigCfg<N>::LINES;
    const int tid = threadIdx.x;
    const int p = tid % T, l = tid / T;
    long long line = (long long)blockIdx.x * LINES + l;
    const bool live = line < g.A;
    if (!live) line = g.A - 1;
    const size_t base = (size_t)line * N;
    ld.begin(g, line, 0);
    st.begin(g, line, 0);
    cplx v[E];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int j = p + T * m;
        cplx x = ld(base + j, j);
        v[m] = INV ? cswap(x) : x;
    }
    if (ContigCfg<N>::WARP_SYNC)
        line_fft<N, LayoutPadded<N>, SyncWarp>(v, p, l, sm, tw);
    else
        line_fft<N, LayoutPadded<N>, SyncCta>(v, p, l, sm, tw);
    if (live) {
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int j = p + T * m;
            st(base + j, j, INV ? cswap(v[m]) : v[m]);
        }
    }
}

// ---- any length (non power of two): O(n^2) DFT per line, out of place -----------
// Completes the FFTWWrapper contract (FFTW accepts any n); not a performance path.
template <bool INV>
__global__ void k_pass_dft_generic(PassGeom g, const cplx* __restrict__ in, cplx* __restrict__ out,
                                   const cplx* __restrict__ tw, double scale) {
    const long long total = g.A * g.N * g.B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i % g.B;
        const long long k = (i / g.B) % g.N;
        const long long a = i / (g.B * g.N);
        const cplx* src = in + (size_t)a * g.N * g.B + b;
        double sx = 0.0, sy = 0.0;
        long long e = 0;  // (j*k) mod N
        for (int j = 0; j < g.N; ++j) {
            cplx w = tw[e];
            if (INV) w.y = -w.y;
            cplx x = src[(size_t)j * g.B];
            sx += x.x * w.x - x.y * w.y;
            sy += x.x * w.y + x.y * w.x;
            e += k;
            if (e >= g.N) e -= g.N;
        }
        out[i] = mk(sx * scale, sy * scale);
    }
}

// ---- launch helpers --------------------------------------------------------------
inline int pick_tx(int N, long long B, int want) {
    int tx = want;
    while (tx > 1 && ((long long)N * tx * 16 > 128 * 1024)) tx >>= 1;  // keep >= 1 CTA of headroom
    while (tx > 1 && (B % tx) != 0) tx >>= 1;
    if (tx < 2) tx = (B % 2 == 0) ? 2 : 1;
    return tx;
}

template <int N, int TX, bool INV, class Ld, class St>
cudaError_t launch_strided_n_tx(const PassGeom& g, Ld ld, St st, const cplx* tw, cudaStream_t s) {
    constexpr int T = PlanFor<N>::T;
    const size_t smem = PlanFor<N>::NS > 1 ? (size_t)N * TX * sizeof(cplx) : 0;
    auto kern = k_pass_strided<N, TX, INV, Ld, St>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long tiles = g.A * (g.B / TX);
    kern<<<(unsigned)tiles, T * TX, smem, s>>>(g, ld, st, tw);
    return cudaGetLastError();
}

template <int N, bool INV, class Ld, class St>
cudaError_t launch_strided_n(const PassGeom& g, int tx, Ld ld, St st, const cplx* tw, cudaStream_t s) {
    constexpr size_t line_bytes = (size_t)N * sizeof(cplx);
    switch (tx) {
        case 2: return launch_strided_n_tx<N, 2, INV>(g, ld, st, tw, s);
        case 4:
            if constexpr (line_bytes * 4 <= 200 * 1024) return launch_strided_n_tx<N, 4, INV>(g, ld, st, tw, s);
            break;
        case 8:
            if constexpr (line_bytes * 8 <= 200 * 1024) return launch_strided_n_tx<N, 8, INV>(g, ld, st, tw, s);
            break;
        case 16:
            if constexpr (line_bytes * 16 <= 200 * 1024 && PlanFor<N>::T * 16 <= 1024)
                return launch_strided_n_tx<N, 16, INV>(g, ld, st, tw, s);
            break;
        default: break;
    }
    return cudaErrorInvalidConfiguration;
}

template <int N, bool INV, class Ld, class St>
cudaError_t launch_contig_n(const PassGeom& g, Ld ld, St st, const cplx* tw, cudaStream_t s) {
    constexpr int T = ContigCfg<N>::T, LINES = ContigCfg<N>::LINES;
    const size_t smem = PlanFor<N>::NS > 1 ? (size_t)LayoutPadded<N>::elems(N, LINES) * sizeof(cplx) : 0;
    auto kern = k_pass_contig<N, INV, Ld, St>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long blocks = (g.A + LINES - 1) / LINES;
    kern<<<(unsigned)blocks, T * LINES, smem, s>>>(g, ld, st, tw);
    return cudaGetLastError();
}

inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
inline bool fast_length(int n) { return is_pow2(n) && n >= 2 && n <= 4096; }

#define GOPF_FOR_EACH_N(X) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)

// One pass along g.axis with the given functors (power-of-two length only).
template <bool INV, class Ld, class St>
cudaError_t launch_pass(const PassGeom& g, int tx_want, Ld ld, St st, const cplx* tw, cudaStream_t s) {
    if (g.B == 1) {
        switch (g.N) {
#define X(n) case n: return launch_contig_n<n, INV>(g, ld, st, tw, s);
            GOPF_FOR_EACH_N(X)
#undef X
            default: return cudaErrorInvalidValue;
        }
    }
    const int tx = pick_tx(g.N, g.B, tx_want);
    if (tx < 2) return cudaErrorInvalidValue;
    switch (g.N) {
#define X(n) case n: return launch_strided_n<n, INV>(g, tx, ld, st, tw, s);
        GOPF_FOR_EACH_N(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace gopf
