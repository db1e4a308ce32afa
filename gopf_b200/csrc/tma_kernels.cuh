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
// maps are rank 5 over doubles: (2*column, row_low, row_high, slab_low, slab_high), outer dimensions in
// ascending stride order; row_high serves the split row addressing of the slab-sharded layouts, the slab split
// the blocked k-space layout of large grids (RowMap in fft_kernels.cuh).
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
    int log, mask;    // row j -> (j & mask, j >> log)
    int alog, amask;  // slab a -> (a & amask, a >> alog)
    int box_rows;     // rows per bulk copy (<= 256, divides the low row dimension)
    int p_lo, p_hi, p_a, p_ahi;
};

struct TmaCoord {
    int c[5];
};
__device__ __forceinline__ TmaCoord tma_coord(const TmaRows& R, int c0, int row, int a) {
    const int lo = row & R.mask, hi = row >> R.log;
    const int alo = a & R.amask, ahi = a >> R.alog;
    TmaCoord q;  // selects, not indexed stores: the coordinates stay in registers
    q.c[0] = c0;
#pragma unroll
    for (int i = 1; i < 5; ++i) q.c[i] = R.p_lo == i ? lo : (R.p_hi == i ? hi : (R.p_a == i ? alo : ahi));
    return q;
}
__device__ __forceinline__ void tma_load_rows(void* dst, const CUtensorMap* map, const TmaRows& R, int c0, int row, int a,
                                              uint64_t* bar) {
    const TmaCoord q = tma_coord(R, c0, row, a);
    tma::load_5d(dst, map, q.c[0], q.c[1], q.c[2], q.c[3], q.c[4], bar);
}
__device__ __forceinline__ void tma_store_rows(const CUtensorMap* map, const TmaRows& R, int c0, int row, int a, const void* src) {
    const TmaCoord q = tma_coord(R, c0, row, a);
    tma::store_5d(map, q.c[0], q.c[1], q.c[2], q.c[3], q.c[4], src);
}
__device__ __forceinline__ void tma_prefetch_rows(const CUtensorMap* map, const TmaRows& R, int c0, int row, int a) {
    const TmaCoord q = tma_coord(R, c0, row, a);
    tma::prefetch_5d(map, q.c[0], q.c[1], q.c[2], q.c[3], q.c[4]);
}

// control block behind the buffers
#define GOPF_TMA_MAX_STAGES 24
struct TmaCtl {
    unsigned long long full[GOPF_TMA_MAX_STAGES];  // per stage: the bytes have landed
    volatile unsigned issued[GOPF_TMA_MAX_STAGES];  // per stage: loads issued so far (guards the parity wait, see tma_wait_tile)
    volatile long long tile[GOPF_TMA_MAX_STAGES];   // per stage: the tile that was claimed for it (-1: no tiles left)
};

struct TmaKCtl {
    TmaCtl ring;
    unsigned long long sfull[4];  // per group: the spectrum tile has landed
};

// The workers interleave on the ring, so a worker can reach stage s for its use u before the worker that
// held the stage last has even issued that load; a bare parity wait would then pass on the phase of use u-2.
// The issue counter closes that window.
__device__ __forceinline__ void tma_wait_tile(TmaCtl* ctl, int stage, unsigned use) {
    if (ctl->issued[stage] <= use) {
        const unsigned long long t0 = tma::global_timer_ns();
        unsigned spins = 0;
        while (ctl->issued[stage] <= use) {
            if ((++spins & 1023u) == 0 && tma::global_timer_ns() - t0 > GOPF_TMA_WAIT_LIMIT_NS)
                tma::wait_timed_out(1, (unsigned)stage, use);
        }
    }
    tma::mbar_wait(reinterpret_cast<uint64_t*>(&ctl->full[stage]), use & 1u, 2);
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
    static constexpr size_t tw_bytes() { return (size_t)TwShared::elems(N) * sizeof(cplx); }
    static constexpr size_t smem_bytes() { return STAGES * buf_bytes() + tw_bytes() + sizeof(TmaKCtl) + 128; }
    // cell (row, line) of the landed / outgoing tile (hardware swizzle: 16-B chunk index ^= address bits 7..)
    static __device__ __forceinline__ int sw(int row, int l) { return row * TX + (l ^ ((row >> ROW_SHIFT) & (TX - 1))); }
};

// Tiles are handed out in order through a global counter (`next_tile`, zeroed before the launch), not by a
// static stride: persistent CTAs drift apart, and with a static partition the 64-B row segments in flight at
// any instant end up planes apart -- measured 5.6 TB/s at 55 tiles per CTA falling to 2.8 TB/s at 1771
// (1024^3).  With in-order claims the ~450 tiles in flight stay within two planes, like the hardware's own
// block scheduler does for a non-persistent grid.  A claim is made one tile ahead of its use, so the atomic's
// round trip is off the critical path.
template <int N, int TX>
__global__ void __launch_bounds__(TmaCfg<N, TX>::THREADS, 1)
    k_pass_strided_tma(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout,
                       const __grid_constant__ PassGeom g, const __grid_constant__ TmaRows rin, const __grid_constant__ TmaRows rout, int inv,
                       double scale,
                       const cplx* __restrict__ tw, unsigned* __restrict__ next_tile) {
    typedef TmaCfg<N, TX> C;
    typedef LayoutPadded<N> Lay;
    constexpr int E = C::E, T = C::T, GT = C::GT, STAGES = C::STAGES;
    extern __shared__ __align__(1024) unsigned char gopf_smem_raw[];
    cplx* twsm = reinterpret_cast<cplx*>(gopf_smem_raw + STAGES * C::buf_bytes());
    TmaCtl* ctl = reinterpret_cast<TmaCtl*>(gopf_smem_raw + STAGES * C::buf_bytes() + C::tw_bytes());
    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT;
    const int l = gtid / T, t = gtid - l * T;
    const long long tilesB = g.bcount / TX, tiles = g.A * tilesB;
    for (int j = tid; j < N; j += C::THREADS) twsm[TwShared::at(j)] = tw[j];

    auto buffer = [&](int s) -> cplx* { return reinterpret_cast<cplx*>(gopf_smem_raw + (size_t)s * C::buf_bytes()); };
    auto tile_coords = [&](long long tile, int* c0, int* c3) {
        const long long a = tile / tilesB;
        *c0 = (int)(2 * window_col(g, tile - a * tilesB, TX));
        *c3 = (int)a;
    };
    // one thread: stage k % STAGES receives `tile` (or the end marker when the claim ran past the last tile)
    auto issue_load = [&](long long k, long long tile) {
        const int s = (int)(k % STAGES);
        uint64_t* bar = reinterpret_cast<uint64_t*>(&ctl->full[s]);
        if (tile < tiles) {
            int c0, c3;
            tile_coords(tile, &c0, &c3);
            ctl->tile[s] = tile;
            tma::mbar_arrive_expect_tx(bar, (unsigned)C::tile_bytes());
            for (int r = 0; r < N; r += rin.box_rows) tma_load_rows(buffer(s) + (size_t)r * TX, &tin, rin, c0, r, c3, bar);
        } else {
            ctl->tile[s] = -1;
            tma::mbar_arrive(bar);
        }
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
        for (int k = 0; k < STAGES; ++k) issue_load(k, (long long)atomicAdd(next_tile, 1u));

    for (long long k = grp;; k += C::GROUPS) {
        const int s = (int)(k % STAGES);
        cplx* buf = buffer(s);
        tma_wait_tile(ctl, s, (unsigned)(k / STAGES));
        const long long tile = ctl->tile[s];
        if (tile < 0) {
            // No tiles left: pass the end marker on to the stage the other group will wait for, then leave.  The
            // marker re-arms THIS stage's barrier and completes its next phase at once, so every warp of the group
            // must have passed its wait on the current phase first: a warp arriving late would find the barrier
            // two phases on, its parity test would fail for ever (seen as a rare hang on small 2-D grids).
            tma::group_sync(9 + grp, GT);
            if (gtid == 0) issue_load(k + STAGES, tiles);
            break;
        }
        long long claimed = 0;
        if (gtid == 0) claimed = (long long)atomicAdd(next_tile, 1u);  // for stage reuse k + STAGES; consumed after the store
        cplx v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = buf[C::sw(t + T * m, l)];
        if (inv) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
        }
        // tile-wide, but only before the first exchange WRITE (the padded exchange regions of the lines overlay
        // the landed tile): the first butterflies run while the slower warps of the group still read
        line_fft_pre<N, Lay, SyncLine<T>, TwShared>(v, t, l, buf, twsm, [&]() { tma::group_sync(9 + grp, GT); });
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
            tile_coords(tile, &c0, &c3);
            for (int r = 0; r < N; r += rout.box_rows) tma_store_rows(&tout, rout, c0, r, c3, buf + (size_t)r * TX);
            tma::store_commit();
            tma::store_wait_read();  // the copy engine has read the buffer: refill it for the other group
            issue_load(k + STAGES, claimed);
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

template <int N, int TX>
__global__ void __launch_bounds__(TmaCfg<N, TX>::THREADS, 1)
    k_fused_kspace_tma(const __grid_constant__ CUtensorMap tw_in, const __grid_constant__ CUtensorMap tw_out,
                       const __grid_constant__ CUtensorMap ts, const __grid_constant__ PassGeom g,
                       const __grid_constant__ TmaRows rin, const __grid_constant__ TmaRows rout,
                       const __grid_constant__ TmaRows rs, const __grid_constant__ DevKProgram P, FreqTabs ft, const cplx* __restrict__ tw,
                       unsigned* __restrict__ next_tile) {
    typedef TmaCfg<N, TX> C;
    typedef LayoutPadded<N> Lay;
    constexpr int E = C::E, T = C::T, GT = C::GT, STAGES = C::STAGES;
    extern __shared__ __align__(1024) unsigned char gopf_smem_raw[];
    cplx* twsm = reinterpret_cast<cplx*>(gopf_smem_raw + STAGES * C::buf_bytes());
    TmaKCtl* kctl = reinterpret_cast<TmaKCtl*>(gopf_smem_raw + STAGES * C::buf_bytes() + C::tw_bytes());
    TmaCtl* ctl = &kctl->ring;
    for (int j = threadIdx.x; j < N; j += C::THREADS) twsm[TwShared::at(j)] = tw[j];
    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT;
    const int l = gtid / T, t = gtid - l * T;
    uint64_t* sfull = reinterpret_cast<uint64_t*>(&kctl->sfull[grp]);
    const long long tilesB = g.bcount / TX, tiles = g.A * tilesB;

    auto buffer = [&](int s) -> cplx* { return reinterpret_cast<cplx*>(gopf_smem_raw + (size_t)s * C::buf_bytes()); };
    auto tile_col = [&](long long tile, long long* a) -> long long {
        *a = tile / tilesB;
        return window_col(g, tile - *a * tilesB, TX);
    };
    auto issue_load = [&](long long k, long long tile) {  // one thread
        const int s = (int)(k % STAGES);
        uint64_t* bar = reinterpret_cast<uint64_t*>(&ctl->full[s]);
        if (tile < tiles) {
            long long a;
            const int c0 = (int)(2 * tile_col(tile, &a));
            ctl->tile[s] = tile;
            tma::mbar_arrive_expect_tx(bar, (unsigned)C::tile_bytes());
            for (int r = 0; r < N; r += rin.box_rows) tma_load_rows(buffer(s) + (size_t)r * TX, &tw_in, rin, c0, r, (int)a, bar);
            // the tile's spectrum rows into L2 while it waits its turn and runs its forward FFT
            if (g.pf_tiles == 0)
                for (int r = 0; r < N; r += rs.box_rows) tma_prefetch_rows(&ts, rs, c0, r, (int)a);
        } else {
            ctl->tile[s] = -1;
            tma::mbar_arrive(bar);
        }
        __threadfence_block();
        ctl->issued[s] = ctl->issued[s] + 1;
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tma::mbar_init(reinterpret_cast<uint64_t*>(&ctl->full[s]), 1);
            ctl->issued[s] = 0;
        }
        for (int q = 0; q < C::GROUPS; ++q) tma::mbar_init(reinterpret_cast<uint64_t*>(&kctl->sfull[q]), 1);
        tma::fence_barrier_init();
        tma::prefetch_descriptor(&tw_in);
        tma::prefetch_descriptor(&tw_out);
        tma::prefetch_descriptor(&ts);
    }
    __syncthreads();
    if (tid == 0)
        for (int k = 0; k < STAGES; ++k) issue_load(k, (long long)atomicAdd(next_tile, 1u));

    const LineFreq<N, T> lf(t);
    const int spf = (int)g.pf_tiles;  // when the spectrum tile is asked into L2: 0 at load issue, 1 at compute start, 2 never
    unsigned use = 0;  // tiles this group has processed (phase of its sfull barrier)
    for (long long k = grp;; k += C::GROUPS, ++use) {
        const int s = (int)(k % STAGES);
        cplx* buf = buffer(s);
        tma_wait_tile(ctl, s, (unsigned)(k / STAGES));
        const long long tile = ctl->tile[s];
        if (tile < 0) {
            tma::group_sync(9 + grp, GT);  // see k_pass_strided_tma
            if (gtid == 0) issue_load(k + STAGES, tiles);
            break;
        }
        long long claimed = 0;
        if (gtid == 0) claimed = (long long)atomicAdd(next_tile, 1u);
        long long a;
        const long long bcol = tile_col(tile, &a);
        const long long b = bcol + l;
        const int c0 = (int)(2 * bcol);
        if (spf == 1 && gtid == 0)
            for (int r = 0; r < N; r += rs.box_rows) tma_prefetch_rows(&ts, rs, c0, r, (int)a);
        cplx v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = buf[C::sw(t + T * m, l)];
        line_fft_head<N, Lay, SyncLine<T>, TwShared>(v, t, l, buf, twsm, [&]() { tma::group_sync(9 + grp, GT); });
        tma::group_sync(9 + grp, GT);  // every line has read its last exchange: the spectrum tile may land ...
        if (gtid == 0) {
            tma::mbar_arrive_expect_tx(sfull, (unsigned)C::tile_bytes());
            for (int r = 0; r < N; r += rs.box_rows) tma_load_rows(buf + (size_t)r * TX, &ts, rs, c0, r, (int)a, sfull);
        }
        line_fft_tail<N, Lay, SyncLine<T>, TwShared>(v, t, l, buf, twsm);  // ... under the register-only last stage
        // Reference Freq components [row, col, depth] = FFTW axes [1, 2, 0] (fftWrap.go:42-74).
        double fa, fb;
        if (g.axis == 0) {
            fa = ft.f1[ft.off1 + (int)(b / g.n2)];
            fb = ft.f2[(int)(b % g.n2)];
        } else if (g.axis == GOPF_AXIS0_BY_PLANE) {
            fa = ft.f1[ft.off1 + (int)a];
            fb = ft.f2[(int)b];
        } else {  // axis 1
            fa = ft.f2[(int)b];
            fb = ft.rank > 2 ? ft.f0[(int)a] : 0.0;
        }
        const double s2 = fa * fa + fb * fb;
        tma::mbar_wait(sfull, use & 1u, 3);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int j = t + T * m;
            const double fl = lf.at(m);
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
        }
        // the buffer becomes the exchange tile again once the copy engine has read the new spectrum out of it:
        // that wait sits behind the first butterflies of the inverse transform
        line_fft_pre<N, Lay, SyncLine<T>, TwShared>(v, t, l, buf, twsm, [&]() {
            if (gtid == 0) tma::store_wait_read();
            tma::group_sync(9 + grp, GT);
        });
        tma::group_sync(9 + grp, GT);
#pragma unroll
        for (int m = 0; m < E; ++m) buf[C::sw(t + T * m, l)] = cswap(v[m]);
        tma::fence_proxy_async();
        tma::group_sync(9 + grp, GT);
        if (gtid == 0) {
            for (int r = 0; r < N; r += rout.box_rows) tma_store_rows(&tw_out, rout, c0, r, (int)a, buf + (size_t)r * TX);
            tma::store_commit();
            tma::store_wait_read();
            issue_load(k + STAGES, claimed);
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
        STAGES_RAW = (int)((206 * 1024) / (LS * 16)),
        STAGES = STAGES_RAW > GOPF_TMA_MAX_STAGES ? GOPF_TMA_MAX_STAGES : STAGES_RAW
    };
    static constexpr size_t buf_bytes() { return (size_t)LS * sizeof(cplx); }
    static constexpr size_t tw_bytes() { return (size_t)TwShared::elems(N) * sizeof(cplx); }
    static constexpr size_t smem_bytes() { return STAGES * buf_bytes() + tw_bytes() + sizeof(TmaCtl) + 128; }
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
    cplx* twsm = reinterpret_cast<cplx*>(gopf_smem_raw + STAGES * C::buf_bytes());
    TmaCtl* ctl = reinterpret_cast<TmaCtl*>(gopf_smem_raw + STAGES * C::buf_bytes() + C::tw_bytes());
    const int tid = threadIdx.x;
    const int worker = tid / T, p = tid - worker * T;
    for (int j = tid; j < N; j += C::THREADS) twsm[TwShared::at(j)] = tw[j];
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
        line_fft<N, Lay, SyncLine<T>, TwShared>(v, p, 0, buf, twsm);
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = mk(v[m].y * inv_n, v[m].x * inv_n);  // swap back, /N
        if (derived_is_monomial_fast(D)) {
            const int pw = D.ipower[0];
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = derived_fast(pw, v[m]);
        } else if (derived_is_poly_fast(D)) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = derived_poly(D, v[m]);
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
        line_fft<N, Lay, SyncLine<T>, TwShared>(v, p, 0, buf, twsm);
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

// ---- the same for REAL fields: two lines per complex transform --------------------------------------------
// When the field is real in real space, a line of W here is the transform of a real line (Hermitian along this
// axis), and two such lines A, B ride through ONE complex transform each way:
//   z = IDFT(A + i B) = a + i b          (a, b the two real lines)
//   Z' = DFT(g(a) + i g(b))              (g real: a monomial or a real polynomial)
//   G_A[k] = (Z'[k] + conj Z'[N-k]) / 2,  G_B[k] = (Z'[k] - conj Z'[N-k]) / (2i)
// Half the butterflies and 30 % less shared-memory traffic per line: k_fused_real_tma<1024> is bound by exactly
// those two (fp64 pipe and shared-memory pipe both ~64 % busy at 0.73 of the HBM peak).  Against the
// complex-carrying form the result differs by rounding only (the imaginary residue of one line, ~1e-17 relative,
// lands in the other instead of being carried along); the reference tolerance is 1e-10.
// A worker (T threads) owns two line buffers for the whole kernel: both lines land, are combined into registers,
// buffer A serves as exchange region, the separated results leave from B (line A's) and A (line B's).
template <int N>
struct TmaRealPairCfg {
    enum {
        E = PlanFor<N>::E,
        T = PlanFor<N>::T,
        LS = N + N / 16,
        PER_WORKER = (LS + N) * 16,
        W_RAW = (int)((206 * 1024) / PER_WORKER),
        WORKERS = W_RAW * T > 512 ? 512 / T : W_RAW,
        THREADS = WORKERS * T
    };
    static constexpr size_t tw_bytes() { return (size_t)TwShared::elems(N) * sizeof(cplx); }
    static constexpr size_t smem_bytes() { return (size_t)WORKERS * PER_WORKER + tw_bytes() + WORKERS * sizeof(unsigned long long) + 128; }
};

__device__ __forceinline__ double real_ipow(int p, double x) {  // x^p by square-and-multiply, p <= 15
    double r = (p & 1) ? x : 1.0;
    double sq = x;
#pragma unroll
    for (int b = 1; b < 4; ++b) {
        if ((p >> b) == 0) break;
        sq = sq * sq;
        if ((p >> b) & 1) r = r * sq;
    }
    return r;
}

template <int N>
__global__ void __launch_bounds__(TmaRealPairCfg<N>::THREADS, 1)
    k_fused_real_pair_tma(cplx* __restrict__ W, long long pairs, const __grid_constant__ DevDerived D, double inv_n,
                          const cplx* __restrict__ tw) {
    typedef TmaRealPairCfg<N> C;
    constexpr int E = C::E, T = C::T, WORKERS = C::WORKERS;
    struct Lay {
        static __device__ __forceinline__ int at(int pos, int) { return pos + (pos >> 4); }
    };
    extern __shared__ __align__(128) unsigned char gopf_smem_raw[];
    cplx* twsm = reinterpret_cast<cplx*>(gopf_smem_raw + (size_t)WORKERS * C::PER_WORKER);
    uint64_t* bars = reinterpret_cast<uint64_t*>(gopf_smem_raw + (size_t)WORKERS * C::PER_WORKER + C::tw_bytes());
    const int tid = threadIdx.x;
    const int worker = tid / T, p = tid - worker * T;
    for (int j = tid; j < N; j += C::THREADS) twsm[TwShared::at(j)] = tw[j];
    cplx* bufA = reinterpret_cast<cplx*>(gopf_smem_raw + (size_t)worker * C::PER_WORKER);  // LS cells: landing + exchange
    cplx* bufB = bufA + C::LS;                                                             // N cells
    uint64_t* bar = bars + worker;
    constexpr unsigned LINE_BYTES = (unsigned)(N * sizeof(cplx));
    const long long hop = (long long)gridDim.x * WORKERS;
    long long pair = (long long)blockIdx.x + (long long)worker * gridDim.x;

    auto issue_load = [&](long long q) {  // one thread of the worker
        tma::mbar_arrive_expect_tx(bar, 2 * LINE_BYTES);
        tma::load_1d(bufA, W + (size_t)(2 * q) * N, LINE_BYTES, bar);
        tma::load_1d(bufB, W + (size_t)(2 * q + 1) * N, LINE_BYTES, bar);
    };
    if (tid == 0) {
        for (int w = 0; w < WORKERS; ++w) tma::mbar_init(bars + w, 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (p == 0 && pair < pairs) issue_load(pair);

    const bool mono = derived_is_monomial_fast(D);
    const int pw = D.ipower[0];
    for (unsigned it = 0; pair < pairs; pair += hop, ++it) {
        tma::mbar_wait(bar, it & 1u, 4);
        cplx v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const cplx a = bufA[p + T * m], b = bufB[p + T * m];
            v[m] = mk(a.y + b.x, a.x - b.y);  // cswap(A + i B): the inverse transform by the swap identity
        }
        SyncLine<T>::run();  // both lines are in registers: buffer A becomes the exchange region
        line_fft<N, Lay, SyncLine<T>, TwShared>(v, p, 0, bufA, twsm);
        // swap back and /N: a = Re z = v.y / N, b = Im z = v.x / N; then g on each
        if (mono) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = mk(real_ipow(pw, v[m].y * inv_n), real_ipow(pw, v[m].x * inv_n));
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m)
                v[m] = mk(derived_poly(D, mk(v[m].y * inv_n, 0.0)).x, derived_poly(D, mk(v[m].x * inv_n, 0.0)).x);
        }
        line_fft<N, Lay, SyncLine<T>, TwShared>(v, p, 0, bufA, twsm);
        // separate the two transforms: Z' in natural order in buffer A, partners read across the line
#pragma unroll
        for (int m = 0; m < E; ++m) bufA[p + T * m] = v[m];
        SyncLine<T>::run();
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int k = p + T * m;
            const cplx q = bufA[(N - k) & (N - 1)];
            bufB[k] = mk(0.5 * (v[m].x + q.x), 0.5 * (v[m].y - q.y));  // line A's transform
            v[m] = mk(0.5 * (v[m].y + q.y), 0.5 * (q.x - v[m].x));     // line B's
        }
        SyncLine<T>::run();  // every partner read of buffer A is done
#pragma unroll
        for (int m = 0; m < E; ++m) bufA[p + T * m] = v[m];
        tma::fence_proxy_async();
        SyncLine<T>::run();
        if (p == 0) {
            tma::store_1d(W + (size_t)(2 * pair) * N, bufB, LINE_BYTES);
            tma::store_1d(W + (size_t)(2 * pair + 1) * N, bufA, LINE_BYTES);
            tma::store_commit();
            tma::store_wait_read();
            if (pair + hop < pairs) issue_load(pair + hop);
        }
        __syncwarp();
    }
    if (p == 0) tma::store_wait_all();
}

}  // namespace gopf
