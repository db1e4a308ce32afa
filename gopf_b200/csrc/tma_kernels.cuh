// Warp-specialised, copy-engine-fed line kernels for long lines (N >= 512 cells).
//
// Why: a tile of TX lines of 1024 complex128 cells is 16 KB * TX.  With the register-resident kernels of
// fft_kernels.cuh / step_kernels.cuh an SM holds two 64-KB tiles (512 threads at 128 registers) and every CTA
// runs load -> FFT -> store in lock-step behind CTA-wide barriers, so HBM idles while the tile computes, the
// SM idles while it loads, and the shared-memory, fp64 and issue phases of the 16 resident warps line up
// instead of overlapping (ncu, 1024^3: 55 % of DRAM peak, issue slots 32 % busy, fp64 pipe 24 %;
// profiles/r1b_ncu_full_1024.md; measured on B200: each resource is 30-40 % busy and the tile time is
// their SUM).  Here ONE persistent CTA per SM owns a ring of shared-memory buffers fed by the copy engine:
//
//   copy engine (TMA):   global -> buffer ................................. UTMALDG / UBLKCP, mbarrier complete_tx
//   line worker:         buffer -> registers, Stockham FFT (exchanges in the same buffer), registers -> buffer
//   copy engine:         buffer -> global ................................. UTMASTG / UBLKCP, bulk_group
//
// No thread issues a global load or store, and -- because coalescing no longer constrains which thread holds
// which cell -- a line belongs to T = N/16 consecutive threads (two warps at N = 1024, one at 512), which
// synchronise among themselves only (named barrier per line / __syncwarp): eight to sixteen independent
// workers per SM whose phases drift apart and overlap.  The FFT is the engine of fft_engine.cuh, unchanged:
// results are bitwise those of the register-resident kernels.
//
// Strided tiles land as [row][TX] through a tensor map with the 64-B / 128-B shared-memory swizzle, so the
// line-major reads (eight consecutive rows of one column per quarter-warp) are bank-conflict free.  The tensor
// maps are rank 4 over doubles: (2*column, row_low, row_high, slab); row_high serves the split row addressing
// of the slab-sharded layouts (RowMap in fft_kernels.cuh).
#pragma once
#include "fft_kernels.cuh"
#include "step_kernels.cuh"
#include "tma.cuh"

namespace gopf {

template <int GT>
struct SyncGroup {
    static __device__ __forceinline__ void run() { tma::group_sync(1 + (int)(threadIdx.x / GT), GT); }
};
// one line = T consecutive threads: a warp (or less) synchronises with __syncwarp, two or more warps with the
// named barrier of the line's slot in the CTA (ids 1 .. THREADS/T <= 8)
template <int T>
struct SyncLine {
    static __device__ __forceinline__ void run() {
        if (T <= 32)
            __syncwarp();
        else
            tma::group_sync(1 + (int)(threadIdx.x / T), T);
    }
};

// row addressing of one side (input or output) of a pass: tile row j -> coordinates (j & mask, j >> log)
// The three outer tensor dimensions (row_low, row_high, slab) are encoded in ascending stride order; p_lo, p_hi,
// p_a give the coordinate slot (1..3) of each.
struct TmaRows {
    int log, mask;
    int box_rows;  // rows per bulk copy (<= 256, divides the low row dimension)
    int p_lo, p_hi, p_a;
};

struct TmaCoord {
    int c[4];
};
__device__ __forceinline__ TmaCoord tma_coord(const TmaRows& R, int c0, int row, int a) {
    const int lo = row & R.mask, hi = row >> R.log;
    TmaCoord q;  // selects, not indexed stores: the coordinates stay in registers
    q.c[0] = c0;
    q.c[1] = R.p_lo == 1 ? lo : (R.p_hi == 1 ? hi : a);
    q.c[2] = R.p_lo == 2 ? lo : (R.p_hi == 2 ? hi : a);
    q.c[3] = R.p_lo == 3 ? lo : (R.p_hi == 3 ? hi : a);
    return q;
}
__device__ __forceinline__ void tma_load_rows(void* dst, const CUtensorMap* map, const TmaRows& R, int c0, int row, int a,
                                              uint64_t* bar) {
    const TmaCoord q = tma_coord(R, c0, row, a);
    tma::load_4d(dst, map, q.c[0], q.c[1], q.c[2], q.c[3], bar);
}
__device__ __forceinline__ void tma_store_rows(const CUtensorMap* map, const TmaRows& R, int c0, int row, int a, const void* src) {
    const TmaCoord q = tma_coord(R, c0, row, a);
    tma::store_4d(map, q.c[0], q.c[1], q.c[2], q.c[3], src);
}
__device__ __forceinline__ void tma_prefetch_rows(const CUtensorMap* map, const TmaRows& R, int c0, int row, int a) {
    const TmaCoord q = tma_coord(R, c0, row, a);
    tma::prefetch_4d(map, q.c[0], q.c[1], q.c[2], q.c[3]);
}

// control block behind the buffers
#define GOPF_TMA_MAX_STAGES 24
struct TmaCtl {
    unsigned long long full[GOPF_TMA_MAX_STAGES];  // per stage: the bytes have landed
    volatile unsigned issued[GOPF_TMA_MAX_STAGES];  // per stage: loads issued so far (guards the parity wait, see tma_wait_tile)
};

// The workers interleave on the ring, so a worker can reach stage s for its use u before the worker that
// held the stage last has even issued that load; a bare parity wait would then pass on the phase of use u-2.
// The issue counter closes that window.
__device__ __forceinline__ void tma_wait_tile(TmaCtl* ctl, int stage, unsigned use) {
    while (ctl->issued[stage] <= use) {
    }
    tma::mbar_wait(reinterpret_cast<uint64_t*>(&ctl->full[stage]), use & 1u);
}

// ---- strided axis pass: out = FFT(in) along rows, TX adjacent lines per tile ---------------------------
// Tile = N rows x TX columns (64 KB), three buffers, two tiles in compute.  Threads are line-major: thread
// gtid of a group holds line l = gtid / T, position t = gtid % T (+ T*m).  Buffer = LayoutPadded exchange
// regions of the TX lines (N + N/16 cells each); the tile lands swizzled in its first N*TX cells.
template <int N, int TX>
struct TmaCfg {
    enum {
        E = PlanFor<N>::E,
        T = PlanFor<N>::T,
        GT = T * TX,          // threads of one tile
        GROUPS = 2,
        STAGES = 3,
        CELLS = N * TX,       // cells of one tile
        PADDED = (N + N / 16) * TX,
        THREADS = GROUPS * GT,
        ROW_SHIFT = (TX == 4 ? 1 : 0),  // smem address bits 7.. = row >> ROW_SHIFT (rows of 64 B / 128 B)
        SWIZZLE = (TX == 4 ? 2 : (TX == 8 ? 3 : 0))  // CU_TENSOR_MAP_SWIZZLE_64B / _128B
    };
    static constexpr size_t tile_bytes() { return (size_t)CELLS * sizeof(cplx); }
    static constexpr size_t buf_bytes() { return (size_t)PADDED * sizeof(cplx); }
    static constexpr size_t smem_bytes() { return STAGES * buf_bytes() + sizeof(TmaCtl) + 128; }
    // cell (row, line) of the landed / outgoing tile (hardware swizzle: 16-B chunk index ^= address bits 7..)
    static __device__ __forceinline__ int sw(int row, int l) { return row * TX + (l ^ ((row >> ROW_SHIFT) & (TX - 1))); }
};

template <int N, int TX>
__global__ void __launch_bounds__(TmaCfg<N, TX>::THREADS, 1)
    k_pass_strided_tma(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout,
                       const __grid_constant__ PassGeom g, TmaRows rin, TmaRows rout, int inv, double scale,
                       const cplx* __restrict__ tw) {
    typedef TmaCfg<N, TX> C;
    typedef LayoutPadded<N> Lay;
    constexpr int E = C::E, T = C::T, GT = C::GT, STAGES = C::STAGES;
    extern __shared__ __align__(1024) unsigned char gopf_smem_raw[];
    TmaCtl* ctl = reinterpret_cast<TmaCtl*>(gopf_smem_raw + STAGES * C::buf_bytes());
    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT;
    const int l = gtid / T, t = gtid - l * T;
    const long long tilesB = g.bcount / TX, tiles = g.A * tilesB;
    const long long first = blockIdx.x, hop = gridDim.x;
    const long long mine = first < tiles ? (tiles - first + hop - 1) / hop : 0;  // tiles of this CTA

    auto buffer = [&](int s) -> cplx* { return reinterpret_cast<cplx*>(gopf_smem_raw + (size_t)s * C::buf_bytes()); };
    auto tile_coords = [&](long long k, int* c0, int* c3) {
        const long long tile = first + k * hop;
        const long long a = tile / tilesB;
        *c0 = (int)(2 * window_col(g, tile - a * tilesB, TX));
        *c3 = (int)a;
    };
    auto issue_load = [&](long long k) {  // one thread
        const int s = (int)(k % STAGES);
        int c0, c3;
        tile_coords(k, &c0, &c3);
        uint64_t* bar = reinterpret_cast<uint64_t*>(&ctl->full[s]);
        tma::mbar_arrive_expect_tx(bar, (unsigned)C::tile_bytes());
        for (int r = 0; r < N; r += rin.box_rows) tma_load_rows(buffer(s) + (size_t)r * TX, &tin, rin, c0, r, c3, bar);
        __threadfence_block();
        ctl->issued[s] = ctl->issued[s] + 1;
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tma::mbar_init(reinterpret_cast<uint64_t*>(&ctl->full[s]), 1);
            ctl->issued[s] = 0;
        }
        tma::fence_barrier_init();
        tma::prefetch_descriptor(&tin);
        tma::prefetch_descriptor(&tout);
    }
    __syncthreads();
    if (tid == 0)
        for (long long k = 0; k < mine && k < STAGES; ++k) issue_load(k);

    for (long long k = grp; k < mine; k += C::GROUPS) {
        const int s = (int)(k % STAGES);
        cplx* buf = buffer(s);
        tma_wait_tile(ctl, s, (unsigned)(k / STAGES));
        cplx v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = buf[C::sw(t + T * m, l)];
        if (inv) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
        }
        // tile-wide: the padded exchange regions of the lines overlay the landed tile
        tma::group_sync(9 + grp, GT);
        line_fft<N, Lay, SyncLine<T> >(v, t, l, buf, tw);
        tma::group_sync(9 + grp, GT);  // every line has read its last exchange: the outgoing tile may overwrite them
        if (inv) {
#pragma unroll
            for (int m = 0; m < E; ++m) buf[C::sw(t + T * m, l)] = mk(v[m].y * scale, v[m].x * scale);
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) buf[C::sw(t + T * m, l)] = mk(v[m].x * scale, v[m].y * scale);
        }
        tma::fence_proxy_async();
        tma::group_sync(9 + grp, GT);
        if (gtid == 0) {
            int c0, c3;
            tile_coords(k, &c0, &c3);
            for (int r = 0; r < N; r += rout.box_rows) tma_store_rows(&tout, rout, c0, r, c3, buf + (size_t)r * TX);
            tma::store_commit();
            tma::store_wait_read();  // the copy engine has read the buffer: refill it for the other group
            if (k + STAGES < mine) issue_load(k + STAGES);
        }
        __syncwarp();
    }
    if (gtid == 0) tma::store_wait_all();
}

// ---- fused k-space kernel (slowest active axis), fast-form programs -------------------------------------------
// k_fused_kspace (step_kernels.cuh) on the same ring and with the same line-major workers.  One buffer serves a
// tile through its whole life: W lands -> forward FFT (exchanges) -> the spectrum tile lands in the very same
// buffer once the last exchange has been read (it was prefetched into L2 when the tile started) -> Euler update
// in place (pf/euler.go:28-38) -> the new spectrum leaves by bulk store -> inverse FFT of it (exchanges again)
// -> the first inverse pass of the next step leaves as W.  Traffic 64 B per cell, no thread touches global memory
// except for the k-tables.
struct TmaKCtl {
    TmaCtl ring;
    unsigned long long sfull[2];  // per group: the spectrum tile has landed
};

template <int N, int TX>
__global__ void __launch_bounds__(TmaCfg<N, TX>::THREADS, 1)
    k_fused_kspace_tma(const __grid_constant__ CUtensorMap tw_in, const __grid_constant__ CUtensorMap tw_out,
                       const __grid_constant__ CUtensorMap ts, const __grid_constant__ PassGeom g, TmaRows rin, TmaRows rout,
                       TmaRows rs, const __grid_constant__ DevKProgram P, FreqTabs ft, const cplx* __restrict__ tw) {
    typedef TmaCfg<N, TX> C;
    typedef LayoutPadded<N> Lay;
    constexpr int E = C::E, T = C::T, GT = C::GT, STAGES = C::STAGES;
    extern __shared__ __align__(1024) unsigned char gopf_smem_raw[];
    TmaKCtl* kctl = reinterpret_cast<TmaKCtl*>(gopf_smem_raw + STAGES * C::buf_bytes());
    TmaCtl* ctl = &kctl->ring;
    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT;
    const int l = gtid / T, t = gtid - l * T;
    uint64_t* sfull = reinterpret_cast<uint64_t*>(&kctl->sfull[grp]);
    const long long tilesB = g.bcount / TX, tiles = g.A * tilesB;
    const long long first = blockIdx.x, hop = gridDim.x;
    const long long mine = first < tiles ? (tiles - first + hop - 1) / hop : 0;

    auto buffer = [&](int s) -> cplx* { return reinterpret_cast<cplx*>(gopf_smem_raw + (size_t)s * C::buf_bytes()); };
    auto tile_col = [&](long long k, long long* a) -> long long {
        const long long tile = first + k * hop;
        *a = tile / tilesB;
        return window_col(g, tile - *a * tilesB, TX);
    };
    auto issue_load = [&](long long k) {  // one thread
        const int s = (int)(k % STAGES);
        long long a;
        const int c0 = (int)(2 * tile_col(k, &a));
        uint64_t* bar = reinterpret_cast<uint64_t*>(&ctl->full[s]);
        tma::mbar_arrive_expect_tx(bar, (unsigned)C::tile_bytes());
        for (int r = 0; r < N; r += rin.box_rows) tma_load_rows(buffer(s) + (size_t)r * TX, &tw_in, rin, c0, r, (int)a, bar);
        __threadfence_block();
        ctl->issued[s] = ctl->issued[s] + 1;
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tma::mbar_init(reinterpret_cast<uint64_t*>(&ctl->full[s]), 1);
            ctl->issued[s] = 0;
        }
        tma::mbar_init(reinterpret_cast<uint64_t*>(&kctl->sfull[0]), 1);
        tma::mbar_init(reinterpret_cast<uint64_t*>(&kctl->sfull[1]), 1);
        tma::fence_barrier_init();
        tma::prefetch_descriptor(&tw_in);
        tma::prefetch_descriptor(&tw_out);
        tma::prefetch_descriptor(&ts);
    }
    __syncthreads();
    if (tid == 0)
        for (long long k = 0; k < mine && k < STAGES; ++k) issue_load(k);

    unsigned use = 0;  // tiles this group has processed (phase of its sfull barrier)
    for (long long k = grp; k < mine; k += C::GROUPS, ++use) {
        const int s = (int)(k % STAGES);
        cplx* buf = buffer(s);
        long long a;
        const long long bcol = tile_col(k, &a);
        const long long b = bcol + l;
        const int c0 = (int)(2 * bcol);
        if (gtid == 0)  // this tile's spectrum rows into L2 while the forward FFT runs
            for (int r = 0; r < N; r += rs.box_rows) tma_prefetch_rows(&ts, rs, c0, r, (int)a);
        tma_wait_tile(ctl, s, (unsigned)(k / STAGES));
        cplx v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = buf[C::sw(t + T * m, l)];
        tma::group_sync(9 + grp, GT);
        line_fft<N, Lay, SyncLine<T> >(v, t, l, buf, tw);
        tma::group_sync(9 + grp, GT);  // every line has read its last exchange: the spectrum tile may land
        if (gtid == 0) {
            tma::mbar_arrive_expect_tx(sfull, (unsigned)C::tile_bytes());
            for (int r = 0; r < N; r += rs.box_rows) tma_load_rows(buf + (size_t)r * TX, &ts, rs, c0, r, (int)a, sfull);
        }
        // Reference Freq components [row, col, depth] = FFTW axes [1, 2, 0] (fftWrap.go:42-74).
        double fa, fb;
        const double* fline;
        if (g.axis == 0) {
            fa = ft.f1[ft.off1 + (int)(b / g.n2)];
            fb = ft.f2[(int)(b % g.n2)];
            fline = ft.f0;
        } else {  // axis 1
            fa = ft.f2[(int)b];
            fb = ft.rank > 2 ? ft.f0[(int)a] : 0.0;
            fline = ft.f1;
        }
        const double s2 = fa * fa + fb * fb;
        tma::mbar_wait(sfull, use & 1u);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int j = t + T * m;
            const double fl = fline[j];
            const int pos = C::sw(j, l);
            const cplx cur = fast_update(P, fma(fl, fl, s2), buf[pos], v[m]);
            buf[pos] = cur;  // each thread rewrites exactly the cells it read
            v[m] = cswap(cur);
        }
        tma::fence_proxy_async();
        tma::group_sync(9 + grp, GT);
        if (gtid == 0) {
            for (int r = 0; r < N; r += rs.box_rows) tma_store_rows(&ts, rs, c0, r, (int)a, buf + (size_t)r * TX);
            tma::store_commit();
            tma::store_wait_read();  // the new spectrum has left the buffer: it becomes the exchange tile again
        }
        tma::group_sync(9 + grp, GT);
        line_fft<N, Lay, SyncLine<T> >(v, t, l, buf, tw);
        tma::group_sync(9 + grp, GT);
#pragma unroll
        for (int m = 0; m < E; ++m) buf[C::sw(t + T * m, l)] = cswap(v[m]);
        tma::fence_proxy_async();
        tma::group_sync(9 + grp, GT);
        if (gtid == 0) {
            for (int r = 0; r < N; r += rout.box_rows) tma_store_rows(&tw_out, rout, c0, r, (int)a, buf + (size_t)r * TX);
            tma::store_commit();
            tma::store_wait_read();
            if (k + STAGES < mine) issue_load(k + STAGES);
        }
        __syncwarp();
    }
    if (gtid == 0) tma::store_wait_all();
}

// ---- fused real-space kernel on contiguous lines ---------------------------------------------------------
// k_fused_real (step_kernels.cuh) as independent line workers: a line is 16*N bytes of contiguous memory, moved
// by linear bulk copies (UBLKCP) straight to / from the base of its own padded exchange region, so a worker
// (the T threads of one line slot) never synchronises with another.  WORKERS lines in compute, the rest of the
// ring in flight.
template <int N>
struct TmaRealCfg {
    enum {
        E = PlanFor<N>::E,
        T = PlanFor<N>::T,
        THREADS = 512,
        WORKERS = THREADS / T,
        LS = N + N / 16,
        STAGES_RAW = (int)((220 * 1024) / (LS * 16)),
        STAGES = STAGES_RAW > GOPF_TMA_MAX_STAGES ? GOPF_TMA_MAX_STAGES : STAGES_RAW
    };
    static constexpr size_t buf_bytes() { return (size_t)LS * sizeof(cplx); }
    static constexpr size_t smem_bytes() { return STAGES * buf_bytes() + sizeof(TmaCtl) + 128; }
};

// inverse, /N, g(c), forward, in place on W (MODE 0 of k_fused_real)
template <int N>
__global__ void __launch_bounds__(TmaRealCfg<N>::THREADS, 1)
    k_fused_real_tma(cplx* __restrict__ W, long long lines, long long node0, const __grid_constant__ DevDerived D,
                     double inv_n, unsigned long long step, const cplx* __restrict__ tw) {
    typedef TmaRealCfg<N> C;
    constexpr int E = C::E, T = C::T, STAGES = C::STAGES, WORKERS = C::WORKERS;
    // the worker's line sits at the base of the buffer: LayoutPadded with line index 0
    struct Lay {
        static __device__ __forceinline__ int at(int pos, int) { return pos + (pos >> 4); }
    };
    extern __shared__ __align__(128) unsigned char gopf_smem_raw[];
    TmaCtl* ctl = reinterpret_cast<TmaCtl*>(gopf_smem_raw + STAGES * C::buf_bytes());
    const int tid = threadIdx.x;
    const int worker = tid / T, p = tid - worker * T;
    const long long first = blockIdx.x, hop = gridDim.x;
    const long long mine = first < lines ? (lines - first + hop - 1) / hop : 0;  // lines of this CTA
    constexpr unsigned LINE_BYTES = (unsigned)(N * sizeof(cplx));

    auto buffer = [&](int s) -> cplx* { return reinterpret_cast<cplx*>(gopf_smem_raw + (size_t)s * C::buf_bytes()); };
    auto issue_load = [&](long long i) {
        const int s = (int)(i % STAGES);
        uint64_t* bar = reinterpret_cast<uint64_t*>(&ctl->full[s]);
        tma::mbar_arrive_expect_tx(bar, LINE_BYTES);
        tma::load_1d(buffer(s), W + (size_t)(first + i * hop) * N, LINE_BYTES, bar);
        __threadfence_block();
        ctl->issued[s] = ctl->issued[s] + 1;
    };
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tma::mbar_init(reinterpret_cast<uint64_t*>(&ctl->full[s]), 1);
            ctl->issued[s] = 0;
        }
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0)
        for (long long i = 0; i < mine && i < STAGES; ++i) issue_load(i);

    for (long long i = worker; i < mine; i += WORKERS) {
        const int s = (int)(i % STAGES);
        cplx* buf = buffer(s);
        tma_wait_tile(ctl, s, (unsigned)(i / STAGES));
        cplx v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cswap(buf[p + T * m]);
        SyncLine<T>::run();  // the line is in registers: its buffer becomes the exchange region
        line_fft<N, Lay, SyncLine<T> >(v, p, 0, buf, tw);
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = mk(v[m].y * inv_n, v[m].x * inv_n);  // swap back, /N
        if (derived_is_fast(D)) {
            const int pw = D.ipower[0];
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = derived_fast(pw, v[m]);
        } else {
            const size_t base = (size_t)(first + i * hop) * N;
#pragma unroll
            for (int m = 0; m < E; ++m) buf[Lay::at(p + T * m, 0)] = v[m];
#pragma unroll 1
            for (int m = 0; m < E; ++m) {
                const int pos = Lay::at(p + T * m, 0);
                const cplx c = buf[pos];
                buf[pos] = eval_derived(D, [&](int) -> cplx { return c; }, step, (unsigned long long)node0 + base + p + T * m);
            }
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = buf[Lay::at(p + T * m, 0)];
            SyncLine<T>::run();
        }
        line_fft<N, Lay, SyncLine<T> >(v, p, 0, buf, tw);
#pragma unroll
        for (int m = 0; m < E; ++m) buf[p + T * m] = v[m];
        tma::fence_proxy_async();
        SyncLine<T>::run();
        if (p == 0) {
            tma::store_1d(W + (size_t)(first + i * hop) * N, buf, LINE_BYTES);
            tma::store_commit();
            tma::store_wait_read();
            if (i + STAGES < mine) issue_load(i + STAGES);
        }
        __syncwarp();
    }
    if (p == 0) tma::store_wait_all();
}

}  // namespace gopf
