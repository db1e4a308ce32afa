// Warp-specialised, copy-engine-fed line kernels for long lines (N >= 512 cells).
//
// Why: a tile of TX lines of 1024 complex128 cells is 16 KB * TX.  With the register-resident kernels of
// fft_kernels.cuh / step_kernels.cuh an SM holds two 64-KB tiles (512 threads at 128 registers) and every CTA
// runs load -> FFT -> store in lock-step, so HBM idles while the tile computes and the SM idles while it
// loads (ncu, 1024^3: 55 % of DRAM peak, issue slots 32 % busy, 16 warps per SM;
// profiles/r1b_ncu_full_1024.md).  Here ONE persistent CTA per SM owns a ring of three 64-KB tile buffers
// and two consumer groups of 256 threads:
//
//   copy engine (TMA):   load tile k+1 / k+2 into a free buffer ............ UTMALDG, mbarrier complete_tx
//   group k % 2:         buffer -> registers, Stockham FFT (exchanges in the same buffer), registers -> buffer
//   copy engine:         store the buffer ................................. UTMASTG, bulk_group
//
// No thread ever issues a global load or store: the two groups compute out of phase while the third buffer
// is in flight, 12 288 cells per SM instead of 8 192, and the global traffic is decoupled from the warps'
// instruction streams (no lg_throttle / long_scoreboard stalls).  The FFT itself is the engine of
// fft_engine.cuh, unchanged: results are bitwise those of the register-resident kernels.
//
// The tensor maps are rank 4 over doubles: (2*column, row_low, row_high, slab); row_high serves the split row
// addressing of the slab-sharded layouts (RowMap in fft_kernels.cuh).  L2 promotion is set to 256 B so that
// the 64-B row segments of adjacent tiles share DRAM bursts.
#pragma once
#include "fft_kernels.cuh"
#include "step_kernels.cuh"
#include "tma.cuh"

namespace gopf {

template <int GT>
struct SyncGroup {
    static __device__ __forceinline__ void run() { tma::group_sync(1 + (int)(threadIdx.x / GT), GT); }
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

template <int N, int TX>
struct TmaCfg {
    enum {
        E = PlanFor<N>::E,
        T = PlanFor<N>::T,
        GT = T * TX,          // threads of one consumer group
        GROUPS = 2,
        STAGES = 3,
        CELLS = N * TX,       // cells of one tile
        THREADS = GROUPS * GT
    };
    static constexpr size_t tile_bytes() { return (size_t)CELLS * sizeof(cplx); }
    static constexpr size_t smem_bytes() { return STAGES * tile_bytes() + 128; }
};

// control block behind the tile buffers
struct TmaCtl {
    unsigned long long full[4];  // per stage: bytes of the tile have landed
    volatile unsigned issued[4];  // per stage: loads issued so far (guards the parity wait, see wait_tile)
};

// The two consumer groups interleave on the ring, so a group can reach stage s for its use u before the
// other group has even issued that load; a bare parity wait would then pass on the phase of use u-2.  The
// issue counter closes that window.
__device__ __forceinline__ void tma_wait_tile(TmaCtl* ctl, int stage, unsigned use) {
    while (ctl->issued[stage] <= use) {
    }
    tma::mbar_wait(reinterpret_cast<uint64_t*>(&ctl->full[stage]), use & 1u);
}

// ---- strided axis pass: out = FFT(in) along rows, TX adjacent lines per tile ---------------------------
template <int N, int TX>
__global__ void __launch_bounds__(TmaCfg<N, TX>::THREADS, 1)
    k_pass_strided_tma(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout,
                       const __grid_constant__ PassGeom g, TmaRows rin, TmaRows rout, int inv, double scale,
                       const cplx* __restrict__ tw) {
    typedef TmaCfg<N, TX> C;
    constexpr int E = C::E, T = C::T, GT = C::GT, CELLS = C::CELLS, STAGES = C::STAGES;
    extern __shared__ __align__(128) unsigned char gopf_smem_raw[];
    cplx* bufs = reinterpret_cast<cplx*>(gopf_smem_raw);
    TmaCtl* ctl = reinterpret_cast<TmaCtl*>(gopf_smem_raw + STAGES * C::tile_bytes());
    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT;
    const int l = gtid % TX, t = gtid / TX;
    const long long tilesB = g.bcount / TX, tiles = g.A * tilesB;
    const long long first = blockIdx.x, hop = gridDim.x;
    const long long mine = first < tiles ? (tiles - first + hop - 1) / hop : 0;  // tiles of this CTA

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
        for (int r = 0; r < N; r += rin.box_rows)
            tma_load_rows(bufs + (size_t)s * CELLS + (size_t)r * TX, &tin, rin, c0, r, c3, bar);
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
        cplx* buf = bufs + (size_t)s * CELLS;
        tma_wait_tile(ctl, s, (unsigned)(k / STAGES));
        cplx v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = buf[(t + T * m) * TX + l];
        if (inv) {
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cswap(v[m]);
        }
        SyncGroup<GT>::run();  // every cell of the tile is in registers: the buffer becomes the exchange tile
        line_fft<N, LayoutInterleaved<TX>, SyncGroup<GT> >(v, t, l, buf, tw);
        if (inv) {
#pragma unroll
            for (int m = 0; m < E; ++m) buf[(t + T * m) * TX + l] = mk(v[m].y * scale, v[m].x * scale);
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) buf[(t + T * m) * TX + l] = mk(v[m].x * scale, v[m].y * scale);
        }
        tma::fence_proxy_async();
        SyncGroup<GT>::run();
        if (gtid == 0) {
            int c0, c3;
            tile_coords(k, &c0, &c3);
            for (int r = 0; r < N; r += rout.box_rows)
                tma_store_rows(&tout, rout, c0, r, c3, buf + (size_t)r * TX);
            tma::store_commit();
            tma::store_wait_read();  // the copy engine has read the buffer: refill it for the other group
            if (k + STAGES < mine) issue_load(k + STAGES);
        }
        __syncwarp();
    }
    if (gtid == 0) tma::store_wait_all();
}

// ---- fused real-space kernel on contiguous lines ---------------------------------------------------------
// k_fused_real (step_kernels.cuh) with the same ring: a tile is LINES adjacent lines = one contiguous run of
// LINES*N cells, moved by linear bulk copies (UBLKCP).  In the buffer the lines land back to back; the FFT
// exchanges use the padded layout inside the same buffer (which is therefore LayoutPadded-sized).
template <int N>
struct TmaRealCfg {
    enum {
        E = PlanFor<N>::E,
        T = PlanFor<N>::T,
        LINES = (65536 / (N * 16)) < 1 ? 1 : (65536 / (N * 16)),
        GT = T * LINES,
        GROUPS = 2,
        STAGES = 3,
        CELLS = N * LINES,
        PADDED = (N + N / 16) * LINES,
        THREADS = GROUPS * GT
    };
    static constexpr size_t buf_bytes() { return (size_t)PADDED * sizeof(cplx); }
    static constexpr size_t smem_bytes() { return STAGES * buf_bytes() + 128; }
};

// MODE 0: inverse, /N, g(c), forward (in place on W).  `lines` must be a multiple of LINES.
template <int N>
__global__ void __launch_bounds__(TmaRealCfg<N>::THREADS, 1)
    k_fused_real_tma(cplx* __restrict__ W, long long lines, long long node0, const __grid_constant__ DevDerived D,
                     double inv_n, unsigned long long step, const cplx* __restrict__ tw) {
    typedef TmaRealCfg<N> C;
    typedef LayoutPadded<N> Lay;
    constexpr int E = C::E, T = C::T, GT = C::GT, STAGES = C::STAGES, LINES = C::LINES;
    extern __shared__ __align__(128) unsigned char gopf_smem_raw[];
    TmaCtl* ctl = reinterpret_cast<TmaCtl*>(gopf_smem_raw + STAGES * C::buf_bytes());
    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT;
    const int p = gtid % T, l = gtid / T;
    const long long tiles = lines / LINES;
    const long long first = blockIdx.x, hop = gridDim.x;
    const long long mine = first < tiles ? (tiles - first + hop - 1) / hop : 0;
    constexpr unsigned TILE_BYTES = (unsigned)(C::CELLS * sizeof(cplx));

    auto buffer = [&](int s) -> cplx* { return reinterpret_cast<cplx*>(gopf_smem_raw + (size_t)s * C::buf_bytes()); };
    auto issue_load = [&](long long k) {
        const int s = (int)(k % STAGES);
        uint64_t* bar = reinterpret_cast<uint64_t*>(&ctl->full[s]);
        tma::mbar_arrive_expect_tx(bar, TILE_BYTES);
        tma::load_1d(buffer(s), W + (size_t)(first + k * hop) * C::CELLS, TILE_BYTES, bar);
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
        for (long long k = 0; k < mine && k < STAGES; ++k) issue_load(k);

    for (long long k = grp; k < mine; k += C::GROUPS) {
        const int s = (int)(k % STAGES);
        cplx* buf = buffer(s);
        tma_wait_tile(ctl, s, (unsigned)(k / STAGES));
        cplx v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cswap(buf[l * N + p + T * m]);
        SyncGroup<GT>::run();
        line_fft<N, Lay, SyncGroup<GT> >(v, p, l, buf, tw);
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = mk(v[m].y * inv_n, v[m].x * inv_n);  // swap back, /N
        if (derived_is_fast(D)) {
            const int pw = D.ipower[0];
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = derived_fast(pw, v[m]);
        } else {
            const size_t base = (size_t)((first + k * hop) * LINES + l) * N;
#pragma unroll
            for (int m = 0; m < E; ++m) buf[Lay::at(p + T * m, l)] = v[m];
#pragma unroll 1
            for (int m = 0; m < E; ++m) {
                const int pos = Lay::at(p + T * m, l);
                const cplx c = buf[pos];
                buf[pos] = eval_derived(D, [&](int) -> cplx { return c; }, step, (unsigned long long)node0 + base + p + T * m);
            }
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = buf[Lay::at(p + T * m, l)];
            SyncGroup<GT>::run();
        }
        line_fft<N, Lay, SyncGroup<GT> >(v, p, l, buf, tw);
#pragma unroll
        for (int m = 0; m < E; ++m) buf[l * N + p + T * m] = v[m];
        tma::fence_proxy_async();
        SyncGroup<GT>::run();
        if (gtid == 0) {
            tma::store_1d(W + (size_t)(first + k * hop) * C::CELLS, buf, TILE_BYTES);
            tma::store_commit();
            tma::store_wait_read();
            if (k + STAGES < mine) issue_load(k + STAGES);
        }
        __syncwarp();
    }
    if (gtid == 0) tma::store_wait_all();
}

// ---- fused k-space kernel (slowest axis), fast-form programs ------------------------------------------------
// k_fused_kspace (step_kernels.cuh) with the copy engine: each consumer group owns one tile buffer (W in ->
// exchange -> W out) and the groups take turns on a third buffer for the spectrum tile (S in -> S out).
// Turn k belongs to CTA-local tile k; its group takes the buffer when turn k-1 has released it (mbarrier
// `sfree`, one phase per turn), loads S under its forward FFT, updates, stores the new S and releases.
// The next tile's W rows are prefetched into L2 by the copy engine while the current tile computes.
struct TmaKCtl {
    unsigned long long wfull[2];  // per group: W tile landed
    unsigned long long sfull;     // spectrum tile landed (one phase per turn)
    unsigned long long sfree;     // spectrum buffer released (one phase per turn)
};

template <int N, int TX>
__global__ void __launch_bounds__(TmaCfg<N, TX>::THREADS, 1)
    k_fused_kspace_tma(const __grid_constant__ CUtensorMap tw_in, const __grid_constant__ CUtensorMap tw_out,
                       const __grid_constant__ CUtensorMap ts, const __grid_constant__ PassGeom g, TmaRows rows, TmaRows srows,
                       const __grid_constant__ DevKProgram P, FreqTabs ft, const cplx* __restrict__ tw) {
    typedef TmaCfg<N, TX> C;
    typedef LayoutInterleaved<TX> Lay;
    constexpr int E = C::E, T = C::T, GT = C::GT, CELLS = C::CELLS;
    extern __shared__ __align__(128) unsigned char gopf_smem_raw[];
    cplx* bufs = reinterpret_cast<cplx*>(gopf_smem_raw);
    TmaKCtl* ctl = reinterpret_cast<TmaKCtl*>(gopf_smem_raw + 3 * C::tile_bytes());
    const int tid = threadIdx.x;
    const int grp = tid / GT, gtid = tid - grp * GT;
    const int l = gtid % TX, t = gtid / TX;
    cplx* wbuf = bufs + (size_t)grp * CELLS;
    cplx* sbuf = bufs + (size_t)2 * CELLS;
    uint64_t* wfull = reinterpret_cast<uint64_t*>(&ctl->wfull[grp]);
    uint64_t* sfull = reinterpret_cast<uint64_t*>(&ctl->sfull);
    uint64_t* sfree = reinterpret_cast<uint64_t*>(&ctl->sfree);
    const long long tilesB = g.bcount / TX, tiles = g.A * tilesB;
    const long long first = blockIdx.x, hop = gridDim.x;
    const long long mine = first < tiles ? (tiles - first + hop - 1) / hop : 0;

    auto tile_col = [&](long long k, long long* a) -> long long {
        const long long tile = first + k * hop;
        *a = tile / tilesB;
        return window_col(g, tile - *a * tilesB, TX);
    };
    auto issue_w = [&](long long k) {  // group leader
        long long a;
        const int c0 = (int)(2 * tile_col(k, &a));
        tma::mbar_arrive_expect_tx(wfull, (unsigned)C::tile_bytes());
        for (int r = 0; r < N; r += rows.box_rows)
            tma_load_rows(wbuf + (size_t)r * TX, &tw_in, rows, c0, r, (int)a, wfull);
    };
    auto prefetch_next = [&](long long k) {  // next tile of this group: W and S rows into L2
        long long a;
        const int c0 = (int)(2 * tile_col(k, &a));
        for (int r = 0; r < N; r += rows.box_rows) tma_prefetch_rows(&tw_in, rows, c0, r, (int)a);
        for (int r = 0; r < N; r += srows.box_rows) tma_prefetch_rows(&ts, srows, c0, r, (int)a);
    };

    if (tid == 0) {
        tma::mbar_init(reinterpret_cast<uint64_t*>(&ctl->wfull[0]), 1);
        tma::mbar_init(reinterpret_cast<uint64_t*>(&ctl->wfull[1]), 1);
        tma::mbar_init(sfull, 1);
        tma::mbar_init(sfree, 1);
        tma::fence_barrier_init();
        tma::prefetch_descriptor(&tw_in);
        tma::prefetch_descriptor(&tw_out);
        tma::prefetch_descriptor(&ts);
    }
    __syncthreads();
    if (gtid == 0 && grp < mine) issue_w(grp);

    unsigned use = 0;  // tiles this group has processed
    for (long long k = grp; k < mine; k += C::GROUPS, ++use) {
        long long a;
        const long long b = tile_col(k, &a) + l;
        const int c0 = (int)(2 * (b - l));
        tma::mbar_wait(wfull, use & 1u);
        cplx v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = wbuf[(t + T * m) * TX + l];
        SyncGroup<GT>::run();
        fft_stage<N, 0, Lay, SyncGroup<GT> >(v, t, l, wbuf, tw);
        // the spectrum buffer: wait for turn k-1 to release it (every thread: the parity wait on `sfull` below is
        // only sound once turn k-1 is over), then fetch this tile's spectrum under the rest of the forward FFT
        if (k > 0) tma::mbar_wait(sfree, (unsigned)((k - 1) & 1));
        if (gtid == 0) {
            tma::mbar_arrive_expect_tx(sfull, (unsigned)C::tile_bytes());
            for (int r = 0; r < N; r += srows.box_rows) tma_load_rows(sbuf + (size_t)r * TX, &ts, srows, c0, r, (int)a, sfull);
            if (k + C::GROUPS < mine) prefetch_next(k + C::GROUPS);
        }
        line_fft_from<N, 1, Lay, SyncGroup<GT> >(v, t, l, wbuf, tw);

        // Reference Freq components [row, col, depth] = FFTW axes [1, 2, 0] (fftWrap.go:42-74)
        double fa, fb;
        const double* fline;
        if (g.axis == 0) {
            fa = ft.f1[ft.off1 + (int)(b / g.n2)];
            fb = ft.f2[(int)(b % g.n2)];
            fline = ft.f0;
        } else {
            fa = ft.f2[(int)b];
            fb = ft.rank > 2 ? ft.f0[(int)a] : 0.0;
            fline = ft.f1;
        }
        const double s2 = fa * fa + fb * fb;
        tma::mbar_wait(sfull, (unsigned)(k & 1));
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int j = t + T * m;
            const double fl = fline[j];
            const cplx cur = fast_update(P, fma(fl, fl, s2), sbuf[j * TX + l], v[m]);
            sbuf[j * TX + l] = cur;  // each thread rewrites exactly the cells it read
            v[m] = cswap(cur);
        }
        tma::fence_proxy_async();
        SyncGroup<GT>::run();
        if (gtid == 0) {
            for (int r = 0; r < N; r += srows.box_rows) tma_store_rows(&ts, srows, c0, r, (int)a, sbuf + (size_t)r * TX);
            tma::store_commit();
            tma::store_wait_read();   // the copy engine has read the new spectrum: the other group's turn
            tma::mbar_arrive(sfree);
        }
        line_fft<N, Lay, SyncGroup<GT> >(v, t, l, wbuf, tw);
#pragma unroll
        for (int m = 0; m < E; ++m) wbuf[(t + T * m) * TX + l] = cswap(v[m]);
        tma::fence_proxy_async();
        SyncGroup<GT>::run();
        if (gtid == 0) {
            for (int r = 0; r < N; r += rows.box_rows)
                tma_store_rows(&tw_out, rows, c0, r, (int)a, wbuf + (size_t)r * TX);
            tma::store_commit();
            tma::store_wait_read();
            if (k + C::GROUPS < mine) issue_w(k + C::GROUPS);
        }
        __syncwarp();
    }
    if (gtid == 0) tma::store_wait_all();
}

}  // namespace gopf
