// Axis-pass kernels: one HBM read + one HBM write of every 16-byte cell per pass
// (B_T = 32*rank bytes per cell per transform).  Pointwise work rides
// on the load (PassIO::load_kind) or the store (scale) so it costs no extra HBM
// traffic.
//
// Array view for a pass along `axis` of a row-major [n0][n1][n2] array:
//   [A][N][B], element (a, j, b) at (a*N + j)*B + b, N = n[axis],
//   B = product of faster extents (inner stride), A = product of slower extents.
#pragma once
#include "fft_engine.cuh"
#include "kupdate.cuh"
#include "step_program.h"

namespace gopf {

// Address of row j of the strided tile (a, b):
//   a*a_stride + b + (j >> split_log)*split_stride + (j & split_mask)*row_stride
// Uniform arrays use split_log = 31 (no split).  The slab-sharded transform reads the
// all-to-all receive buffer / writes the send buffer through a split map, so the
// pack/unpack of the transpose costs no pass of its own (dist_solver.cu).
// The slab index splits the same way (a_split_*): slab a sits at
//   (a >> a_split_log)*a_split_stride + (a & a_split_mask)*a_stride,
// which is how the blocked k-space layout of large 3-D grids ([n0/2^s][n1][2^s][n2], solver.cu) looks to the
// middle-axis passes (slab = axis-0 index, split into block and position in block).
struct RowMap {
    long long a_stride, row_stride, split_stride;
    int split_log, split_mask;
    long long a_split_stride = 0;
    int a_split_log = 31, a_split_mask = 0x7fffffff;
};
__host__ __device__ __forceinline__ long long slab_off(const RowMap& r, long long a) {
    return (a >> r.a_split_log) * r.a_split_stride + (a & r.a_split_mask) * r.a_stride;
}
__host__ __device__ __forceinline__ long long row_off(const RowMap& r, int j) {
    return (long long)(j >> r.split_log) * r.split_stride + (long long)(j & r.split_mask) * r.row_stride;
}

// Peer-store output of the slab-sharded transform: row j of tile (a, b) belongs to rank
// q = j >> log and is written straight into rank q's receive buffer over NVLink,
//   base[q] + block_off + a*a_stride + (j & mask)*row_stride + b,
// so the transpose exchange is fused into the pass that produces the data (no send buffer, no
// separate collective).  base[own rank] is the local buffer.  n == 0: ordinary output.
#define GOPF_MAX_PEERS 8
struct PeerOut {
    cplx* base[GOPF_MAX_PEERS];
    long long block_off, a_stride, row_stride;
    int log, mask;
    int n;
    int max_ctas;  // > 0: launch at most this many (persistent) CTAs
};

// PassGeom::axis of the fused k-space kernel on the blocked layout: lines run along axis 0, but the tiles are
// numbered (slab = axis-1 index, column = axis-2 index) instead of (slab 0, column = n2*i1 + i2)
#define GOPF_AXIS0_BY_PLANE 3
struct PassGeom {
    int n0, n1, n2;  // extents in FFTW order (2-D: n0 == 1; 1-D: n0 == n1 == 1)
    int axis;        // 0, 1 or 2 (GOPF_AXIS0_BY_PLANE: see above)
    long long A, B;  // outer count / inner stride for this axis
    int N;           // n[axis]
    RowMap in, out;  // strided kernels only
    long long node0; // reference node number of this array's first cell (slab offset when sharded)
    PeerOut peer;    // strided kernels only
    // Column window (chunked launches): the pass covers `bcount` of the B columns, laid out as rows of `bw`
    // consecutive columns starting at b0, one row every `bpitch` columns: window column w is column
    // (w / bw) * bpitch + b0 + w % bw.  bw == bcount: one contiguous run [b0, b0 + bcount).  The slab-sharded
    // k-space kernel uses rows = k1_local, bpitch = n2 and a k2 window, so that the inverse middle pass of that
    // k2 window can start while the next window is still being exchanged (dist_solver.cu).
    long long b0, bcount, bw, bpitch;
    // Next-wave L2 prefetch (plain axis passes): a CTA asks L2 for the input tile of CTA
    // blockIdx.x + pf_tiles (the one that takes its place on the SM, = resident CTAs of the launch),
    // so DRAM keeps streaming while this tile is in its compute phase and the next wave's loads hit
    // L2.  Off by default (GOPF_PREFETCH=1 enables it): inside the fused Cahn-Hilliard step the
    // middle-axis passes gained 4-7 % (scripts/tune_prefetch.py), but isolated passes along the
    // slowest axis lost up to 30 % (scripts/tune_axis0.py: 512^3 2996 -> 2017 GB/s at TX 4) and the
    // fused kernels 15 %.  0: off.
    long long pf_tiles;
    // Persistent kernels (tma_kernels.cuh) launch at most this many CTAs (0: one per SM).  The slab-sharded step
    // sets it while an NVLink-bound pass is confined to a few SMs on the second stream, so that the statically
    // partitioned tiles of the compute-stream kernel are not queued behind it.
    int grid_cap;
    // The caller vouches that the field is REAL in real space (its spectrum Hermitian, every term keeps it so):
    // the copy-engine real-space kernel may then carry two lines through one complex transform each way
    // (k_fused_real_pair_tma, tma_kernels.cuh).  0: lines are general complex data.
    int real_pairs;
};

// GOPF_HOST_EMUL: tests/host_emul compiles this header with g++ and runs the kernels on the host (one OS
// thread per CUDA thread); it sets the macro to skip what only exists on the device or in the CUDA runtime.
#ifdef GOPF_HOST_EMUL
inline void prefetch_l2(const void*) {}
#else
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// resident CTAs of a launch = prefetch distance (host)
int prefetch_enabled();  // pass_launch.cu: GOPF_PREFETCH (default 1)
template <class Kern>
inline long long prefetch_distance(Kern kern, int threads, size_t smem) {
    if (!prefetch_enabled()) return 0;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess || nb < 1) return 0;
    return (long long)nb * sms;
}
#endif  // GOPF_HOST_EMUL

__device__ __forceinline__ cplx* peer_row(const PeerOut& p, long long a, int j, long long b) {
    return p.base[j >> p.log] + (p.block_off + a * p.a_stride + (long long)(j & p.mask) * p.row_stride + b);
}

inline RowMap uniform_rows(long long a_stride, long long row_stride) {
    RowMap r;
    r.a_stride = a_stride;
    r.row_stride = row_stride;
    r.split_stride = 0;
    r.split_log = 31;
    r.split_mask = 0x7fffffff;
    r.a_split_stride = 0;
    r.a_split_log = 31;
    r.a_split_mask = 0x7fffffff;
    return r;
}

inline PassGeom make_geom(int n0, int n1, int n2, int axis) {
    PassGeom g;
    g.n0 = n0; g.n1 = n1; g.n2 = n2; g.axis = axis;
    if (axis == 2) { g.N = n2; g.B = 1; g.A = (long long)n0 * n1; }
    else if (axis == 1) { g.N = n1; g.B = n2; g.A = n0; }
    else { g.N = n0; g.B = (long long)n1 * n2; g.A = 1; }
    g.in = g.out = uniform_rows((long long)g.N * g.B, g.B);
    g.node0 = 0;
    g.peer = PeerOut{};
    g.b0 = 0;
    g.bcount = g.B;
    g.bw = g.B;
    g.bpitch = 0;
    g.pf_tiles = 0;
    g.grid_cap = 0;
    g.real_pairs = 0;
    return g;
}

// first column of tile `t` (TX columns wide, TX divides bw) of the column window
__host__ __device__ __forceinline__ long long window_col(const PassGeom& g, long long t, int tx) {
    const long long w = t * tx;
    if (g.bw == g.bcount) return g.b0 + w;
    const long long r = w / g.bw;
    return r * g.bpitch + g.b0 + (w - r * g.bw);
}

struct RealPtrs {
    const cplx* r[GOPF_MAX_FIELDS];  // real-space fields (complex-carrying, like the reference)
};

// What a pass reads and writes.  One POD for every use keeps the number of kernel
// instantiations at (lengths x tile widths) instead of multiplying it by functor types.
enum LoadKind {
    LK_PLAIN = 0,     // in[idx]
    LK_DERIVED = 1,   // derived-field value from the real-space fields (model.go:237-241)
    LK_GRADIENT = 2,  // i 2 pi f_comp * in[idx], +0.5 Nyquist zeroed (squareGradientTerm.go:45-50)
    LK_SUM_SQUARES = 3,  // sum_d g_d[idx]^2 (squareGradientTerm.go:53-62, summed before the transform)
    LK_ELAST_H = 4,      // H(re phi[idx]), phi = g[0]                       (homoLinElast.go:53-57)
    LK_ELAST_R = 5,      // H'(phi) * in[idx] - aux * H(phi) * H'(phi)        (homoLinElast.go:64-97, by linearity)
    LK_MUL_TABLE = 6,    // rtab[idx] * in[idx]: tabulated real k-space multiplier (elastic.cuh M(k))
    // LK_GRADIENT applied in the pass that runs ALONG the component's axis: the multiplier i 2 pi f_comp depends on
    // the row of the line only (rtab = that axis' Freq table, FftPlan::freq_axis), and it commutes with the
    // transforms along the other axes, so it need not sit in the first pass of the inverse transform.  No index
    // decomposition per cell (LK_GRADIENT: two 64-bit divisions, 0.41 instead of 0.09 ms per 256^3 pass).
    LK_GRADIENT_LINE = 7
};

struct PassIO {
    const cplx* in;
    cplx* out;
    double scale;  // applied on store (1/N on the last inverse pass, sliceOperations.go:34-40)
    int inv;       // 0: forward (sign -1), 1: inverse (sign +1)
    int load_kind;
    // LK_DERIVED
    DevDerived D;
    RealPtrs R;
    unsigned long long step;
    // LK_GRADIENT
    FreqGeom fg;
    int comp;
    // LK_SUM_SQUARES, LK_ELAST_*
    int dim;
    const cplx* g[3];
    double aux;
    // LK_MUL_TABLE
    const double* rtab;
};

inline PassIO plain_io(const cplx* in, cplx* out, bool inverse, double scale) {
    PassIO io;
    io.in = in;
    io.out = out;
    io.scale = scale;
    io.inv = inverse ? 1 : 0;
    io.load_kind = LK_PLAIN;
    io.step = 0;
    io.comp = 0;
    io.dim = 0;
    io.D.kind = 0;
    io.D.n_factors = 0;
    io.D.n_ops = 0;
    io.aux = 0.0;
    io.rtab = nullptr;
    return io;
}

// Non-plain loads (run from a rolled loop, see pass_load_line).
__device__ __forceinline__ cplx pass_load_slow(const PassIO& io, size_t idx) {
    cplx x;
    if (io.load_kind == LK_DERIVED) {
        x = eval_derived(io.D, [&](int f) -> cplx { return io.R.r[f][idx]; }, io.step, idx);
    } else if (io.load_kind == LK_GRADIENT) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(io.fg, (long long)idx, f);
        double fd = f[io.comp];
        if (fabs(fd - 0.5) < 1e-10) fd = 0.0;
        const double w = 2.0 * GOPF_PI * fd;
        const cplx u = io.in[idx];
        x = mk(-u.y * w, u.x * w);
    } else if (io.load_kind == LK_ELAST_H) {
        const double p = io.g[0][idx].x;
        x = mk(3.0 * p * p - 2.0 * p * p * p, 0.0);
    } else if (io.load_kind == LK_ELAST_R) {
        const double p = io.g[0][idx].x;
        const double h = 3.0 * p * p - 2.0 * p * p * p, dh = 6.0 * p - 6.0 * p * p;
        const cplx e = io.in[idx];
        x = mk(dh * e.x - io.aux * (h * dh), dh * e.y);
    } else if (io.load_kind == LK_MUL_TABLE) {
        const double w = io.rtab[idx];
        const cplx u = io.in[idx];
        x = mk(u.x * w, u.y * w);
    } else {
        cplx a = io.g[0][idx];
        x = a * a;
        a = io.g[1][idx];
        x += a * a;
        if (io.dim > 2) {
            a = io.g[2][idx];
            x += a * a;
        }
    }
    return x;
}

// Loads the E cells of one thread.  `at(m)` maps slot m to its array index.  The plain case
// issues all E independent 128-bit loads back to back before any is consumed.
// Every other load kind evaluates its interpreter in a ROLLED loop whose results are staged
// in the thread's own shared-memory cells (`sat(m)`), so the interpreter is instantiated once
// and the register-resident line never spills.
// `row(m)` is the position of slot m along the line.
template <int E, class At, class SAt, class Row>
__device__ __forceinline__ void pass_load_line(const PassIO& io, cplx (&v)[E], At at, cplx* sm, SAt sat, Row row) {
    if (io.load_kind == LK_PLAIN) {
        const cplx* __restrict__ in = io.in;
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = in[at(m)];
    } else if (io.load_kind == LK_GRADIENT_LINE) {
        const cplx* __restrict__ in = io.in;
        const double* __restrict__ ftab = io.rtab;
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = in[at(m)];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            double fd = ftab[row(m)];
            if (fabs(fd - 0.5) < 1e-10) fd = 0.0;  // squareGradientTerm.go:45-50
            const double w = 2.0 * GOPF_PI * fd;
            v[m] = mk(-v[m].y * w, v[m].x * w);
        }
    } else if (io.load_kind == LK_SUM_SQUARES) {
        // two or three loads per cell, all in flight (the rolled loop below serialises them)
        const cplx* __restrict__ g0 = io.g[0];
        const cplx* __restrict__ g1 = io.g[1];
        const cplx* __restrict__ g2 = io.g[2];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = g0[at(m)];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = v[m] * v[m];
        for (int d = 1; d < io.dim; ++d) {
            const cplx* __restrict__ gd = d == 1 ? g1 : g2;
            cplx u[E];
#pragma unroll
            for (int m = 0; m < E; ++m) u[m] = gd[at(m)];
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] += u[m] * u[m];
        }
    } else if (io.load_kind == LK_ELAST_H) {
        // simple arithmetic on one or two loads per cell: keep every load of the thread in flight
        // like the plain case (the rolled loop below serialises them: 1.3 vs 1.0 ms per 512^3 pass)
        const cplx* __restrict__ phi = io.g[0];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = phi[at(m)];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const double p = v[m].x;
            v[m] = mk(3.0 * p * p - 2.0 * p * p * p, 0.0);
        }
    } else if (io.load_kind == LK_ELAST_R) {
        const cplx* __restrict__ in = io.in;
        const double* __restrict__ phi = reinterpret_cast<const double*>(io.g[0]);
        double p[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = in[at(m)];
#pragma unroll
        for (int m = 0; m < E; ++m) p[m] = phi[2 * at(m)];  // real part only
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const double h = 3.0 * p[m] * p[m] - 2.0 * p[m] * p[m] * p[m], dh = 6.0 * p[m] - 6.0 * p[m] * p[m];
            v[m] = mk(dh * v[m].x - io.aux * (h * dh), dh * v[m].y);
        }
    } else if (io.load_kind == LK_MUL_TABLE) {
        const cplx* __restrict__ in = io.in;
        const double* __restrict__ tab = io.rtab;
        double w[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = in[at(m)];
#pragma unroll
        for (int m = 0; m < E; ++m) w[m] = tab[at(m)];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = mk(v[m].x * w[m], v[m].y * w[m]);
    } else if (io.load_kind == LK_DERIVED && derived_is_monomial_fast(io.D)) {
        // field^p, p a small integer (pf/model.go:237-241): square-and-multiply on register-resident cells, all
        // loads in flight (the interpreter loop below ran this pass at 2.1 TB/s, 256^3)
        const cplx* __restrict__ in = io.R.r[io.D.field[0]];
        const int pw = io.D.ipower[0];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = in[at(m)];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = derived_fast(pw, v[m]);
#ifdef GOPF_JIT_LOAD_LINE
    } else if (io.load_kind == LK_DERIVED) {
        // run-time specialisation (jit.cu): the registered function as straight-line code on
        // register-resident cells, all loads of a batch in flight like the plain case
        GOPF_JIT_LOAD_LINE(io, v, at)
#endif
    } else {
#pragma unroll 1
        for (int m = 0; m < E; ++m) sm[sat(m)] = pass_load_slow(io, at(m));
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = sm[sat(m)];
        __syncthreads();  // before any thread's first exchange write can land on these cells
    }
    if (io.inv) {
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
    }
}

template <int E, class At>
__device__ __forceinline__ void pass_store_line(const PassIO& io, cplx (&v)[E], At at) {
    cplx* __restrict__ out = io.out;
    const double sc = io.scale;
    if (io.inv) {
#pragma unroll
        for (int m = 0; m < E; ++m) out[at(m)] = mk(v[m].y * sc, v[m].x * sc);
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) out[at(m)] = mk(v[m].x * sc, v[m].y * sc);
    }
}

// same with one destination pointer per cell (peer-store output)
template <int E, class Ptr>
__device__ __forceinline__ void pass_store_ptr(const PassIO& io, cplx (&v)[E], Ptr ptr) {
    const double sc = io.scale;
    if (io.inv) {
#pragma unroll
        for (int m = 0; m < E; ++m) *ptr(m) = mk(v[m].y * sc, v[m].x * sc);
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) *ptr(m) = mk(v[m].x * sc, v[m].y * sc);
    }
}

// ---- strided axis (B > 1): tile = N x TX, TX adjacent lines ---------------------
template <int N, int TX, bool PEER>
__device__ __forceinline__ void pass_strided_tile(const PassGeom& g, const PassIO& io, const cplx* __restrict__ tw,
                                                  cplx* sm, long long tile) {
    constexpr int E = PlanFor<N>::E, T = PlanFor<N>::T;
    const int tid = threadIdx.x;
    const int l = tid % TX, t = tid / TX;
    const long long tilesB = g.bcount / TX;
    const long long a = tile / tilesB;
    const long long b = window_col(g, tile - a * tilesB, TX) + l;
    cplx v[E];
    const size_t ibase = (size_t)slab_off(g.in, a) + b, obase = (size_t)slab_off(g.out, a) + b;
    auto at_in = [&](int m) -> size_t {
        const int j = t + T * m;
        return ibase + (size_t)(j >> g.in.split_log) * g.in.split_stride + (size_t)(j & g.in.split_mask) * g.in.row_stride;
    };
    auto at_out = [&](int m) -> size_t {
        const int j = t + T * m;
        return obase + (size_t)(j >> g.out.split_log) * g.out.split_stride +
               (size_t)(j & g.out.split_mask) * g.out.row_stride;
    };
    pass_load_line<E>(io, v, at_in, sm, [&](int m) -> int { return LayoutInterleaved<TX>::at(t + T * m, l); },
                      [&](int m) -> int { return t + T * m; });
    if (!PEER && g.pf_tiles > 0 && io.load_kind == LK_PLAIN) {
        const long long tile2 = tile + g.pf_tiles;
        if (tile2 < g.A * tilesB) {
            const long long a2 = tile2 / tilesB;
            const size_t ib2 = (size_t)slab_off(g.in, a2) + (size_t)(window_col(g, tile2 - a2 * tilesB, TX) + l);
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int j = t + T * m;
                prefetch_l2(io.in + ib2 + (size_t)(j >> g.in.split_log) * g.in.split_stride +
                            (size_t)(j & g.in.split_mask) * g.in.row_stride);
            }
        }
    }
    line_fft<N, LayoutInterleaved<TX>, SyncCta>(v, t, l, sm, tw);
    if (PEER)
        pass_store_ptr<E>(io, v, [&](int m) -> cplx* { return peer_row(g.peer, a, t + T * m, b); });
    else
        pass_store_line<E>(io, v, at_out);
}

// PEER: persistent over tiles (grid-stride), so the launch can confine the NVLink-bound pass to
// a few SMs (PeerOut::max_ctas) and leave the rest to the HBM-bound kernels of the next chunk
// running on the compute stream.  Every thread has passed the barrier after the last exchange
// read before it stores, so the next tile's first exchange write needs no extra barrier.
template <int N, int TX, bool PEER>
__global__ void __launch_bounds__(PlanFor<N>::T* TX, GOPF_MINB(PlanFor<N>::T* TX))
    k_pass_strided(const __grid_constant__ PassGeom g, const __grid_constant__ PassIO io, const cplx* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char gopf_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(gopf_smem_raw);
    if (PEER) {
        const long long tiles = g.A * (g.bcount / TX);
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) pass_strided_tile<N, TX, true>(g, io, tw, sm, tile);
    } else {
        pass_strided_tile<N, TX, false>(g, io, tw, sm, blockIdx.x);
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

template <int N>
__global__ void __launch_bounds__(ContigCfg<N>::T* ContigCfg<N>::LINES, GOPF_MINB(ContigCfg<N>::T* ContigCfg<N>::LINES))
    k_pass_contig(const PassGeom g, const __grid_constant__ PassIO io, const cplx* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char gopf_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(gopf_smem_raw);
    constexpr int E = PlanFor<N>::E, T = PlanFor<N>::T, LINES = ContigCfg<N>::LINES;
    const int tid = threadIdx.x;
    const int p = tid % T, l = tid / T;
    long long line = (long long)blockIdx.x * LINES + l;
    const bool live = line < g.A;
    if (!live) line = g.A - 1;
    const size_t base = (size_t)line * N;
    cplx v[E];
    auto at = [&](int m) -> size_t { return base + p + T * m; };
    pass_load_line<E>(io, v, at, sm, [&](int m) -> int { return LayoutPadded<N>::at(p + T * m, l); },
                      [&](int m) -> int { return p + T * m; });
    if (g.pf_tiles > 0 && io.load_kind == LK_PLAIN) {
        const long long line2 = line + g.pf_tiles * LINES;
        if (line2 < g.A) {
#pragma unroll
            for (int m = 0; m < E; ++m) prefetch_l2(io.in + (size_t)line2 * N + p + T * m);
        }
    }
    if (ContigCfg<N>::WARP_SYNC)
        line_fft<N, LayoutPadded<N>, SyncWarp>(v, p, l, sm, tw);
    else
        line_fft<N, LayoutPadded<N>, SyncCta>(v, p, l, sm, tw);
    if (live) pass_store_line<E>(io, v, at);
}

// ---- launch-side helpers ------------------------------------------------------------
int pass_tx_override();  // pass_launch.cu: GOPF_PASS_TX (tuning; 0 = none)

inline int pick_tx(int N, long long B, int want, bool peer = false) {
    const int forced = pass_tx_override();
    if (forced > 0 && !peer) {
        int tx = forced;
        while (tx > 1 && ((long long)N * tx * 16 > 200 * 1024 || (B % tx) != 0)) tx >>= 1;
        return tx;
    }
    // Peer stores go out over NVLink in row segments of TX cells: 128-B segments reach 717 GB/s,
    // 64-B ones 438 GB/s (scripts/peer_store_probe.cu), so the peer-writing passes take TX >= 8
    // even where a narrower tile is the faster local choice.
    if (peer) {
        int tx = want > 8 ? want : 8;
        while (tx > 1 && ((long long)N * tx * 16 > 128 * 1024 || (B % tx) != 0)) tx >>= 1;
        return tx < 2 ? ((B % 2 == 0) ? 2 : 1) : tx;
    }
    // Tile width by measurement (scripts/tune_axis0.py, B200, GB/s of the 32 B/cell pass, TX 4 / 8 / 16):
    //   256^3  slowest axis 4656 / 5762 / 5972   middle axis 5428 / 6347 / 6168
    //   512^3  slowest axis 2996 / 5073 / 4446   middle axis 5569 / 6023 / 4469
    //   1024   slowest axis 3874 / 4028 / 4013   middle axis 4797 / 3900 / 3888
    // Row segments of 128 B matter most where the row stride is large (slowest axis); 1024-cell
    // lines along the middle axis prefer two 64-KB tiles per SM over one 128-KB tile.
    if (N >= 1024 && B < 16384 && want > 4) want = 4;
    else if (N <= 256 && B >= 16384 && want == 8) want = 16;
    int tx = want;
    while (tx > 1 && ((long long)N * tx * 16 > 128 * 1024)) tx >>= 1;
    while (tx > 1 && (B % tx) != 0) tx >>= 1;
    if (tx < 2) tx = (B % 2 == 0) ? 2 : 1;
    return tx;
}

#ifndef GOPF_HOST_EMUL
template <int N, int TX>
cudaError_t launch_strided_n_tx(const PassGeom& g, const PassIO& io, const cplx* tw, cudaStream_t s) {
    constexpr int T = PlanFor<N>::T;
    const size_t smem = (size_t)N * TX * sizeof(cplx);
    auto kern = g.peer.n > 0 ? k_pass_strided<N, TX, true> : k_pass_strided<N, TX, false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    long long tiles = g.A * (g.bcount / TX);
    if (g.peer.n > 0 && g.peer.max_ctas > 0 && tiles > g.peer.max_ctas) tiles = g.peer.max_ctas;
    PassGeom gl = g;
    if (g.peer.n == 0) gl.pf_tiles = prefetch_distance(kern, T * TX, smem);
    kern<<<(unsigned)tiles, T * TX, smem, s>>>(gl, io, tw);
    return cudaGetLastError();
}

template <int N>
cudaError_t launch_strided_n(const PassGeom& g, int tx, const PassIO& io, const cplx* tw, cudaStream_t s) {
    constexpr size_t line_bytes = (size_t)N * sizeof(cplx);
    switch (tx) {
        case 2: return launch_strided_n_tx<N, 2>(g, io, tw, s);
        case 4:
            if constexpr (line_bytes * 4 <= 200 * 1024) return launch_strided_n_tx<N, 4>(g, io, tw, s);
            break;
        case 8:
            if constexpr (line_bytes * 8 <= 200 * 1024) return launch_strided_n_tx<N, 8>(g, io, tw, s);
            break;
        case 16:
            if constexpr (line_bytes * 16 <= 200 * 1024 && PlanFor<N>::T * 16 <= 1024)
                return launch_strided_n_tx<N, 16>(g, io, tw, s);
            break;
        default: break;
    }
    return cudaErrorInvalidConfiguration;
}

template <int N>
cudaError_t launch_contig_n(const PassGeom& g, const PassIO& io, const cplx* tw, cudaStream_t s) {
    constexpr int T = ContigCfg<N>::T, LINES = ContigCfg<N>::LINES;
    const size_t smem = (size_t)LayoutPadded<N>::elems(N, LINES) * sizeof(cplx);
    auto kern = k_pass_contig<N>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long blocks = (g.A + LINES - 1) / LINES;
    PassGeom gl = g;
    gl.pf_tiles = prefetch_distance(kern, T * LINES, smem);
    kern<<<(unsigned)blocks, T * LINES, smem, s>>>(gl, io, tw);
    return cudaGetLastError();
}

#endif  // GOPF_HOST_EMUL

// launch shape of k_pass_contig<N> for A lines (the run-time specialised copy of the kernel is launched
// through the driver API with the same shape, jit.cu)
template <int N>
inline void contig_config_n(long long A, unsigned* grid, unsigned* block, size_t* smem) {
    constexpr int T = ContigCfg<N>::T, LINES = ContigCfg<N>::LINES;
    *smem = (size_t)LayoutPadded<N>::elems(N, LINES) * sizeof(cplx);
    *grid = (unsigned)((A + LINES - 1) / LINES);
    *block = (unsigned)(T * LINES);
}

#ifndef GOPF_HOST_EMUL
template <int N>
cudaError_t launch_pass_n(const PassGeom& g, int tx_want, const PassIO& io, const cplx* tw, cudaStream_t s) {
    if (g.B == 1) return launch_contig_n<N>(g, io, tw, s);
    int tx = pick_tx(N, g.B, tx_want, g.peer.n > 0);
    while (tx > 1 && (g.bw % tx) != 0) tx >>= 1;  // tiles must not straddle a row of the column window
    if (tx < 2) return cudaErrorInvalidValue;
    return launch_strided_n<N>(g, tx, io, tw, s);
}

#endif  // GOPF_HOST_EMUL

inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
inline bool fast_length(int n) { return is_pow2(n) && n >= 2 && n <= 4096; }

// One pass along g.axis (power-of-two length 2..4096).  Defined in pass_launch.cu; the
// instantiations are spread over pass_inst_*.cu so they compile in parallel.
#ifndef GOPF_HOST_EMUL
cudaError_t launch_pass(const PassGeom& g, int tx_want, const PassIO& io, const cplx* tw, cudaStream_t s);
#endif
// pass_launch.cu: contig_config_n for a run-time N; false when N is not a fast length
bool contig_launch_config(int N, long long A, unsigned* grid, unsigned* block, size_t* smem);

}  // namespace gopf
