// Launch side of the copy-engine-fed line kernels: tensor-map construction (cached), shape checks,
// persistent grids of one CTA per SM.
#include "tma_launch.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "fused_launch.h"
#include "tma_kernels.cuh"

namespace gopf {

namespace {

std::atomic<long long> g_tma_launches{0};

int env_flag(const char* name, int fallback) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : fallback;
}

bool enabled(const char* which) { return env_flag("GOPF_TMA", 1) != 0 && env_flag(which, 1) != 0; }
int min_n() { return env_flag("GOPF_TMA_MIN_N", 1024); }

int sm_count_raw() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev] > 0 ? sms[dev] : 148;
}

long long grid_for(const PassGeom& g, long long tiles) {
    long long cap = sm_count_raw();
    if (g.grid_cap > 0 && g.grid_cap < cap) cap = g.grid_cap;
    return std::min<long long>(tiles, cap);
}

// One tensor map + the row addressing the kernel needs for it.
struct MapKey {
    const void* base;
    long long B, a_stride, row_stride, split_stride, a_split_stride, A;
    int N, split_log, a_split_log, tx, promo, dev, swizzle;
    bool operator<(const MapKey& o) const { return std::memcmp(this, &o, sizeof(MapKey)) < 0; }
};
struct MapVal {
    CUtensorMap map;
    TmaRows rows;
};

// rows j of a tile: address a*a_stride + (j >> split_log)*split_stride + (j & mask)*row_stride + b
cudaError_t tensor_map_for(const void* base, long long B, const RowMap& rm, long long A, int N, int tx, int swizzle,
                           MapVal* out) {
    static std::map<MapKey, MapVal> cache;
    static std::mutex mu;
    MapKey key;
    std::memset(&key, 0, sizeof(key));
    key.base = base;
    key.B = B;
    key.a_stride = rm.a_stride;
    key.row_stride = rm.row_stride;
    key.split_stride = rm.split_stride;
    key.A = A;
    key.N = N;
    key.split_log = rm.split_log;
    key.a_split_stride = rm.a_split_stride;
    key.a_split_log = rm.a_split_log;
    key.tx = tx;
    key.promo = env_flag("GOPF_TMA_L2", 3);
    key.swizzle = swizzle;
    cudaGetDevice(&key.dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return cudaSuccess;
    }
    const long long rows_lo = rm.split_log >= 30 ? N : std::min<long long>(N, 1LL << rm.split_log);
    const long long rows_hi = N / rows_lo;
    if (rows_lo * rows_hi != N) return cudaErrorNotSupported;
    const long long a_lo = rm.a_split_log >= 30 ? A : std::min<long long>(A, 1LL << rm.a_split_log);
    const long long a_hi = A / a_lo;
    if (a_lo * a_hi != A) return cudaErrorNotSupported;
    // outer dimensions in ascending stride order; extent-1 dimensions go last with a stride that keeps the
    // sequence monotonic (their stride is never used)
    struct Dim {
        long long extent, stride;
        int who;  // 0 row_low, 1 row_high, 2 slab_low, 3 slab_high
    };
    std::vector<Dim> real_dims, unit_dims;
    const Dim cand[4] = {{rows_lo, rm.row_stride, 0}, {rows_hi, rm.split_stride, 1}, {a_lo, rm.a_stride, 2},
                         {a_hi, rm.a_split_stride, 3}};
    for (const Dim& d : cand) (d.extent > 1 ? real_dims : unit_dims).push_back(d);
    std::sort(real_dims.begin(), real_dims.end(), [](const Dim& x, const Dim& y) { return x.stride < y.stride; });
    std::vector<Dim> dims = real_dims;
    long long top = real_dims.empty() ? B : real_dims.back().stride * real_dims.back().extent;
    for (Dim d : unit_dims) {
        d.stride = top;
        dims.push_back(d);
    }
    MapVal val;
    std::memset(&val, 0, sizeof(val));
    val.rows.log = rm.split_log >= 30 ? 31 : rm.split_log;
    val.rows.mask = rm.split_log >= 30 ? 0x7fffffff : (int)(rows_lo - 1);
    val.rows.alog = rm.a_split_log >= 30 ? 31 : rm.a_split_log;
    val.rows.amask = rm.a_split_log >= 30 ? 0x7fffffff : (int)(a_lo - 1);
    val.rows.box_rows = (int)std::min<long long>(256, rows_lo);
    unsigned long long gd[5] = {(unsigned long long)B, 1, 1, 1, 1}, gs[4] = {0, 0, 0, 0};
    unsigned box[5] = {(unsigned)tx, 1, 1, 1, 1};
    for (int i = 0; i < 4; ++i) {
        gd[1 + i] = (unsigned long long)dims[i].extent;
        gs[i] = (unsigned long long)dims[i].stride;
        if (gs[i] == 0 || gs[i] >= (1ULL << 36)) return cudaErrorNotSupported;  // bytes < 2^40
        if (dims[i].who == 0) {
            val.rows.p_lo = 1 + i;
            box[1 + i] = (unsigned)val.rows.box_rows;
        } else if (dims[i].who == 1) {
            val.rows.p_hi = 1 + i;
        } else if (dims[i].who == 2) {
            val.rows.p_a = 1 + i;
        } else {
            val.rows.p_ahi = 1 + i;
        }
    }
    cudaError_t e = tma::encode_c128(&val.map, base, 5, gd, gs, box, key.promo, swizzle);
    if (e != cudaSuccess) return e;
    if (cache.size() > 256) cache.clear();
    cache[key] = val;
    *out = val;
    return cudaSuccess;
}

// Tile counter of one launch of a dynamically scheduled kernel: a slot of a per-device pool, zeroed on the
// launch's stream just before the kernel.  Slots rotate, so kernels running concurrently on different streams
// never share one (a slot comes round again 4096 launches later).
cudaError_t fresh_counter(cudaStream_t s, unsigned** out) {
    constexpr int SLOTS = 4096;
    static unsigned* pool[64] = {nullptr};
    static std::atomic<unsigned> next{0};
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!pool[dev]) {
        std::lock_guard<std::mutex> lock(mu);
        if (!pool[dev]) {
            unsigned* p = nullptr;
            cudaError_t e = cudaMalloc(&p, SLOTS * sizeof(unsigned));
            if (e != cudaSuccess) return e;
            pool[dev] = p;
        }
    }
    unsigned* c = pool[dev] + (next++ % SLOTS);
    cudaError_t e = cudaMemsetAsync(c, 0, sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    *out = c;
    return cudaSuccess;
}

template <class Kern>
cudaError_t opt_in_smem(Kern kern, size_t smem) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <int N, int TX>
cudaError_t pass_tma_n(const PassGeom& g, const PassIO& io, const cplx* tw, cudaStream_t s) {
    typedef TmaCfg<N, TX> C;
    if (g.bw % TX != 0 || g.bcount % TX != 0) return cudaErrorNotSupported;
    MapVal in, out;
    cudaError_t e = tensor_map_for(io.in, g.B, g.in, g.A, N, TX, C::SWIZZLE, &in);
    if (e != cudaSuccess) return cudaErrorNotSupported;
    e = tensor_map_for(io.out, g.B, g.out, g.A, N, TX, C::SWIZZLE, &out);
    if (e != cudaSuccess) return cudaErrorNotSupported;
    auto kern = k_pass_strided_tma<N, TX>;
    e = opt_in_smem(kern, C::smem_bytes());
    if (e != cudaSuccess) return e;
    const long long tiles = g.A * (g.bcount / TX);
    const unsigned grid = (unsigned)grid_for(g, tiles);
    unsigned* ctr = nullptr;
    e = fresh_counter(s, &ctr);
    if (e != cudaSuccess) return e;
    kern<<<grid, C::THREADS, C::smem_bytes(), s>>>(in.map, out.map, g, in.rows, out.rows, io.inv, io.scale, tw, ctr);
    g_tma_launches++;
    return cudaGetLastError();
}

template <int N>
cudaError_t real_tma_n(const PassGeom& g, cplx* W, const DevDerived& D, double inv_n, unsigned long long step, const cplx* tw,
                       cudaStream_t s) {
    typedef TmaRealCfg<N> C;
    auto kern = k_fused_real_tma<N>;
    cudaError_t e = opt_in_smem(kern, C::smem_bytes());
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)grid_for(g, (g.A + C::WORKERS - 1) / C::WORKERS);
    kern<<<grid, C::THREADS, C::smem_bytes(), s>>>(W, g.A, g.node0, D, inv_n, step, tw);
    g_tma_launches++;
    return cudaGetLastError();
}

// real fields: two lines per complex transform (k_fused_real_pair_tma)
template <int N>
cudaError_t real_pair_tma_n(const PassGeom& g, cplx* W, const DevDerived& D, double inv_n, const cplx* tw, cudaStream_t s) {
    typedef TmaRealPairCfg<N> C;
    auto kern = k_fused_real_pair_tma<N>;
    cudaError_t e = opt_in_smem(kern, C::smem_bytes());
    if (e != cudaSuccess) return e;
    const long long pairs = g.A / 2;
    const unsigned grid = (unsigned)grid_for(g, (pairs + C::WORKERS - 1) / C::WORKERS);
    kern<<<grid, C::THREADS, C::smem_bytes(), s>>>(W, pairs, D, inv_n, tw);
    g_tma_launches++;
    return cudaGetLastError();
}

template <int N, int TX>
cudaError_t kspace_tma_n(const PassGeom& g, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P, const FreqTabs& ft,
                         const cplx* tw, cudaStream_t s) {
    typedef TmaCfg<N, TX> C;
    if (g.bw % TX != 0 || g.bcount % TX != 0) return cudaErrorNotSupported;
    MapVal win, wout, sp;
    if (tensor_map_for(W, g.B, g.in, g.A, N, TX, C::SWIZZLE, &win) != cudaSuccess) return cudaErrorNotSupported;
    if (tensor_map_for(Wout, g.B, g.out, g.A, N, TX, C::SWIZZLE, &wout) != cudaSuccess) return cudaErrorNotSupported;
    if (tensor_map_for(S, g.B, g.in, g.A, N, TX, C::SWIZZLE, &sp) != cudaSuccess) return cudaErrorNotSupported;
    auto kern = k_fused_kspace_tma<N, TX>;
    const size_t smem = C::smem_bytes();
    cudaError_t e = opt_in_smem(kern, smem);
    if (e != cudaSuccess) return e;
    const long long tiles = g.A * (g.bcount / TX);
    const unsigned grid = (unsigned)grid_for(g, tiles);
    unsigned* ctr = nullptr;
    e = fresh_counter(s, &ctr);
    if (e != cudaSuccess) return e;
    PassGeom gk = g;
    gk.pf_tiles = env_flag("GOPF_TMA_SPF", 2);  // spectrum-tile L2 prefetch: 0 at load issue, 1 at compute start, 2 never (1024^3: 16.15 / 15.03 / 14.94 ms)
    kern<<<grid, C::THREADS, smem, s>>>(win.map, wout.map, sp.map, gk, win.rows, wout.rows, sp.rows, P, ft, tw, ctr);
    g_tma_launches++;
    return cudaGetLastError();
}

}  // namespace

long long tma_launch_count(bool reset) {
    const long long v = g_tma_launches.load();
    if (reset) g_tma_launches = 0;
    return v;
}

cudaError_t launch_pass_tma(const PassGeom& g, const PassIO& io, const cplx* tw, cudaStream_t s) {
    if (!enabled("GOPF_TMA_PASS") || g.B == 1 || g.N < min_n() || io.load_kind != LK_PLAIN || g.peer.n > 0)
        return cudaErrorNotSupported;
    switch (g.N) {
        case 512: return pass_tma_n<512, 8>(g, io, tw, s);
        case 1024: return pass_tma_n<1024, 4>(g, io, tw, s);
        default: return cudaErrorNotSupported;
    }
}

cudaError_t launch_fused_real_tma(const PassGeom& g, cplx* W, const DevDerived& D, double inv_n, unsigned long long step,
                                  const cplx* tw, cudaStream_t s) {
    // From GOPF_TMA_MIN_N (1024) on.  At 512-cell lines the paired kernel alone beats the register kernel (512^3:
    // 0.89 -> 0.71 ms) but the step does not gain (6.29 -> 6.26 ms, cfg 5): between register kernels with 64-KB tiles
    // it is the only launch with a 221-KB carve-out, and the two reconfigurations per step cost what it saves.
    if (!enabled("GOPF_TMA_REAL") || g.B != 1 || g.N < min_n()) return cudaErrorNotSupported;
    const bool host_fast = (D.kind == DK_MONOMIAL && D.n_factors == 1 && D.ipower[0] >= 0 && D.ipower[0] <= 15) ||
                           (D.kind == DK_RPN && D.poly_deg >= 0);
    if (g.real_pairs && (g.A % 2) == 0 && host_fast && env_flag("GOPF_REAL_PAIRS", 1) != 0) {
        switch (g.N) {
            case 512: return real_pair_tma_n<512>(g, W, D, inv_n, tw, s);
            case 1024: return real_pair_tma_n<1024>(g, W, D, inv_n, tw, s);
            default: break;
        }
    }
    switch (g.N) {
        case 512: return real_tma_n<512>(g, W, D, inv_n, step, tw, s);
        case 1024: return real_tma_n<1024>(g, W, D, inv_n, step, tw, s);
        default: return cudaErrorNotSupported;
    }
}

cudaError_t launch_fused_kspace_tma(const PassGeom& g, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P,
                                    const FreqTabs& ft, const cplx* tw, cudaStream_t s) {
    if (!enabled("GOPF_TMA_KSPACE") || g.N < min_n() || P.fast != 1 || g.peer.n > 0) return cudaErrorNotSupported;
    switch (g.N) {
        case 512: return kspace_tma_n<512, 8>(g, W, Wout, S, P, ft, tw, s);
        case 1024: return kspace_tma_n<1024, 4>(g, W, Wout, S, P, ft, tw, s);
        default: return cudaErrorNotSupported;
    }
}

}  // namespace gopf
