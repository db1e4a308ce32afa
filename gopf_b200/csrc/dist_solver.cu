// DistSolver implementation (see dist_solver.h).
#include "dist_solver.h"

#include <cstring>

namespace gopf {

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

DistSolver::DistSolver(Model* m, int n, int world, int rank, double dt, int device)
    : m_model_(m), n_(n), world_(world), rank_(rank), dt_(dt) {
    if (!m) throw Error("dist solver: model is NULL");
    if (world < 1 || rank < 0 || rank >= world) throw Error("dist solver: bad world / rank");
    if (!is_pow2(n) || !is_pow2(world) || n % world != 0 || n / world < 2)
        throw Error(strf("dist solver: need a power-of-two cubic edge and world with n/world >= 2 (n=%d, world=%d)", n, world));
    if (!fused_length_supported(n)) throw Error(strf("dist solver: edge %d is not supported by the fused kernels", n));
    m_ = n / world;
    m->init();
    for (const HostField& f : m->fields)
        if (f.n != (size_t)m_ * n * n)
            throw Error("dist solver: every field must hold this rank's slab (n/world * n * n cells)");
    derived_ = single_field_derived_index(*m);
    if (derived_ < 0)
        throw Error("dist solver: the sharded path covers the single-field fused step (one field, one equation, "
                    "one nonlinear derived field); this model needs the general path, which is single-GPU");
    const int dims[3] = {n, n, n};
    plan_.reset(new FftPlan(3, dims, device));
    m->fill_program(&prog_, dt, 3);
    if (program_has_knoise(prog_)) throw Error("dist solver: k-space noise is not wired into the sharded path");
    prog_.filter = nullptr;
    prog_.filter_n = 0;
    finalize_single_field_program(&prog_, 1);
    for (int i = 0; i < GOPF_MAX_PEERS; ++i) X_[i] = Y_[i] = nullptr;
}

// max |Im x| over an array (upload: is this rank's slab real?)
static __global__ void k_slab_max_abs_imag(const cplx* __restrict__ x, long long n, double* out) {
    double m = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmax(m, fabs(x[i].y));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0)  // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
}

void DistSolver::check_real(const cplx* W) {
    if (!d_imag_max_) GOPF_CUDA(cudaMalloc(&d_imag_max_, sizeof(double)));
    if (!h_imag_max_) GOPF_CUDA(cudaMallocHost(&h_imag_max_, sizeof(double)));
    cudaStream_t s = stream();
    GOPF_CUDA(cudaMemsetAsync(d_imag_max_, 0, sizeof(double), s));
    const long long n = (long long)local_cells();
    k_slab_max_abs_imag<<<148 * 8, 256, 0, s>>>(W, n, d_imag_max_);
    GOPF_CUDA(cudaGetLastError());
    GOPF_CUDA(cudaMemcpyAsync(h_imag_max_, d_imag_max_, sizeof(double), cudaMemcpyDeviceToHost, s));
    real_check_pending_ = true;
    launches_++;
}

// Each rank decides for its own slab: both kernels are valid for real data, so the ranks need not agree.
int DistSolver::real_pairs() {
    if (real_check_pending_) {
        GOPF_CUDA(cudaStreamSynchronize(stream()));
        field_real_ = *h_imag_max_ == 0.0;
        real_check_pending_ = false;
    }
    return (field_real_ && prog_.fast == 1) ? 1 : 0;
}

DistSolver::~DistSolver() {
    cudaSetDevice(plan_->device);
    if (d_imag_max_) cudaFree(d_imag_max_);
    if (h_imag_max_) cudaFreeHost(h_imag_max_);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    if (ev_compute_) cudaEventDestroy(ev_compute_);
    if (ev_copy_) cudaEventDestroy(ev_copy_);
    for (int q = 0; q < GOPF_MAX_PEERS; ++q) {
        if (q == rank_) {
            if (X_[q]) cudaFree(X_[q]);
            if (Y_[q]) cudaFree(Y_[q]);
        } else {
            if (X_[q]) cudaIpcCloseMemHandle(X_[q]);
            if (Y_[q]) cudaIpcCloseMemHandle(Y_[q]);
        }
    }
}

// ---- peer-store exchange ---------------------------------------------------------------
void DistSolver::peer_alloc() {
    if (world_ > GOPF_MAX_PEERS) throw Error(strf("dist solver: peer-store exchange supports up to %d ranks", GOPF_MAX_PEERS));
    plan_->use_device();
    const size_t bytes = sizeof(cplx) * local_cells();
    if (!X_[rank_]) GOPF_CUDA(cudaMalloc(&X_[rank_], bytes));
    if (!Y_[rank_]) GOPF_CUDA(cudaMalloc(&Y_[rank_], bytes));
}

void DistSolver::peer_export(int which, void* handle64) const {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cplx* p = which == 0 ? X_[rank_] : Y_[rank_];
    if (!p) throw Error("dist solver: peer_alloc() first");
    cudaIpcMemHandle_t h;
    GOPF_CUDA(cudaIpcGetMemHandle(&h, p));
    std::memcpy(handle64, &h, sizeof(h));
}

void DistSolver::peer_import(int which, int rank, const void* handle64) {
    if (rank < 0 || rank >= world_) throw Error("dist solver: peer rank out of range");
    if (rank == rank_) return;  // own buffers are used directly
    plan_->use_device();
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    GOPF_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    (which == 0 ? X_ : Y_)[rank] = reinterpret_cast<cplx*>(p);
}

void DistSolver::peer_unmap() {
    plan_->use_device();
    for (int q = 0; q < GOPF_MAX_PEERS; ++q) {
        if (q == rank_) continue;
        if (X_[q]) cudaIpcCloseMemHandle(X_[q]);
        if (Y_[q]) cudaIpcCloseMemHandle(Y_[q]);
        X_[q] = Y_[q] = nullptr;
    }
}

bool DistSolver::peer_ready() const {
    for (int q = 0; q < world_; ++q)
        if (!X_[q] || !Y_[q]) return false;
    return true;
}

// kspace_rows: the producing pass runs along k0 (geom (n, m, n) axis 0, rows of m*n cells) and
// fills X = [p][i0l][k1l][k2]; otherwise along k1 (geom (m, n, n) axis 1, rows of n cells, one
// tile row index a = i0l) and fills Y = [k0 = p*m + i0l][k1l][k2].
PeerOut DistSolver::peer_out(cplx* const* bufs, bool kspace_rows) const {
    if (!peer_ready()) throw Error("dist solver: peer buffers are not mapped (peer_alloc / peer_import)");
    PeerOut po{};
    for (int q = 0; q < world_; ++q) po.base[q] = bufs[q];
    po.block_off = (long long)rank_ * m_ * m_ * n_;
    po.log = ilog2(m_);
    po.mask = m_ - 1;
    po.n = world_;
    if (kspace_rows) {
        po.a_stride = 0;
        po.row_stride = (long long)m_ * n_;
    } else {
        po.a_stride = (long long)m_ * n_;
        po.row_stride = n_;
    }
    return po;
}

void DistSolver::inverse_start_peer(const cplx* S) {
    plan_->use_device();
    PassGeom g = make_geom(n_, m_, n_, 0);
    g.peer = peer_out(X_, true);
    check(launch_pass(g, plan_->tx_want, plain_io(S, nullptr, true, 1.0), plan_->twiddle(0), stream()),
          "inverse axis 0 -> peers");
}

void DistSolver::forward_mid_peer(const cplx* W) {
    plan_->use_device();
    PassGeom g = make_geom(m_, n_, n_, 1);
    g.peer = peer_out(Y_, false);
    check(launch_pass(g, plan_->tx_want, plain_io(W, nullptr, false, 1.0), plan_->twiddle(1), stream()),
          "forward axis 1 -> peers");
}

void DistSolver::forward_local_peer(cplx* W) {
    plan_->use_device();
    check_real(W);
    check(launch_pass(make_geom(m_, n_, n_, 2), plan_->tx_want, plain_io(W, W, false, 1.0), plan_->twiddle(2), stream()),
          "forward axis 2");
    forward_mid_peer(W);
}

void DistSolver::forward_finish_peer(cplx* S) {
    plan_->use_device();
    check(launch_pass(make_geom(n_, m_, n_, 0), plan_->tx_want, plain_io(Y_[rank_], S, false, 1.0), plan_->twiddle(0),
                      stream()),
          "forward axis 0");
}

void DistSolver::kspace_step_peer(cplx* S) {
    plan_->use_device();
    FreqTabs ft;
    ft.f0 = plan_->freq_axis(0);
    ft.f1 = plan_->freq_axis(1);
    ft.f2 = plan_->freq_axis(2);
    ft.rank = 3;
    ft.off1 = rank_ * m_;
    PassGeom g = make_geom(n_, m_, n_, 0);
    g.peer = peer_out(X_, true);
    check(launch_fused_kspace(g, plan_->tx_want, Y_[rank_], nullptr, S, prog_, ft, plan_->twiddle(0), stream()),
          "fused k-space kernel -> peers");
}

void DistSolver::check(cudaError_t e, const char* what) {
    launches_++;
    if (e != cudaSuccess) throw Error(strf("dist solver: %s: %s", what, cudaGetErrorString(e)));
}

RowMap DistSolver::split_map() const {
    RowMap r;
    r.a_stride = (long long)m_ * n_;
    r.row_stride = n_;
    r.split_stride = (long long)m_ * m_ * n_;
    r.split_log = ilog2(m_);
    r.split_mask = m_ - 1;
    r.a_split_stride = 0;
    r.a_split_log = 31;
    r.a_split_mask = 0x7fffffff;
    return r;
}

PassGeom DistSolver::slab_axis1(bool split_in, bool split_out) const {
    PassGeom g = make_geom(m_, n_, n_, 1);
    if (split_in) g.in = split_map();
    if (split_out) g.out = split_map();
    return g;
}

// ---- chunked phases + copy-engine exchange ------------------------------------------------
void DistSolver::inverse_mid_planes(const cplx* recv, cplx* W, int begin, int count) {
    plan_->use_device();
    PassGeom g = slab_axis1(true, false);
    g.A = count;
    g.grid_cap = grid_cap_;
    check(launch_pass(g, plan_->tx_want,
                      plain_io(recv + (size_t)begin * m_ * n_, W + (size_t)begin * n_ * n_, true, 1.0), plan_->twiddle(1),
                      stream()),
          "inverse axis 1 (planes)");
}

void DistSolver::real_step_planes(cplx* W, int begin, int count) {
    plan_->use_device();
    PassGeom g = make_geom(count, n_, n_, 2);
    g.grid_cap = grid_cap_;
    g.real_pairs = real_pairs();
    g.node0 = ((long long)rank_ * m_ + begin) * n_ * n_;
    const double inv_n = 1.0 / ((double)n_ * n_ * n_);
    check(launch_fused_real(g, 0, W + (size_t)begin * n_ * n_, nullptr, m_model_->derived[derived_].dev, inv_n,
                            (unsigned long long)steps_taken_, plan_->twiddle(2), stream()),
          "fused real-space kernel (planes)");
}

void DistSolver::forward_mid_planes(const cplx* W, cplx* send, int begin, int count) {
    plan_->use_device();
    PassGeom g = slab_axis1(false, true);
    g.A = count;
    check(launch_pass(g, plan_->tx_want,
                      plain_io(W + (size_t)begin * n_ * n_, send + (size_t)begin * m_ * n_, false, 1.0), plan_->twiddle(1),
                      stream()),
          "forward axis 1 (planes)");
}

void DistSolver::kspace_step_cols(const cplx* Tin, cplx* S, cplx* Tout, int k1_begin, int k1_count) {
    plan_->use_device();
    FreqTabs ft;
    ft.f0 = plan_->freq_axis(0);
    ft.f1 = plan_->freq_axis(1);
    ft.f2 = plan_->freq_axis(2);
    ft.rank = 3;
    ft.off1 = rank_ * m_;
    PassGeom g = make_geom(n_, m_, n_, 0);
    g.b0 = (long long)k1_begin * n_;
    g.bcount = g.bw = (long long)k1_count * n_;
    check(launch_fused_kspace(g, plan_->tx_want, Tin, Tout, S, prog_, ft, plan_->twiddle(0), stream()),
          "fused k-space kernel (columns)");
}

void DistSolver::copy_after_compute() {
    if (!peer_ready()) throw Error("dist solver: peer buffers are not mapped (peer_alloc / peer_import)");
    plan_->use_device();
    if (!copy_stream_) {
        int lo = 0, hi = 0;
        GOPF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GOPF_CUDA(cudaStreamCreateWithPriority(&copy_stream_, cudaStreamNonBlocking, hi));
        GOPF_CUDA(cudaEventCreateWithFlags(&ev_compute_, cudaEventDisableTiming));
        GOPF_CUDA(cudaEventCreateWithFlags(&ev_copy_, cudaEventDisableTiming));
    }
    GOPF_CUDA(cudaEventRecord(ev_compute_, stream()));
    GOPF_CUDA(cudaStreamWaitEvent(copy_stream_, ev_compute_, 0));
}

// send = [q][i0l][k1l(q)][k2]; planes [begin, begin+count) of block q -> Y_q[k0 = p*m + i0l][k1l][k2]
void DistSolver::exchange_forward(const cplx* send, int begin, int count) {
    copy_after_compute();
    const size_t plane = (size_t)m_ * n_, block = (size_t)m_ * plane;
    for (int d = 0; d < world_; ++d) {
        const int q = (rank_ + 1 + d) % world_;  // stagger the destinations over the ranks
        GOPF_CUDA(cudaMemcpyAsync(Y_[q] + ((size_t)rank_ * m_ + begin) * plane, send + q * block + begin * plane,
                                  sizeof(cplx) * count * plane, cudaMemcpyDefault, copy_stream_));
    }
}

// T = [q][k0l][k1l][k2]; columns k1l in [k1_begin, k1_begin+k1_count) of block q -> X_q[p][i0l = k0l][k1l][k2]
void DistSolver::exchange_inverse(const cplx* T, int k1_begin, int k1_count) {
    copy_after_compute();
    const size_t plane = (size_t)m_ * n_, block = (size_t)m_ * plane;
    for (int d = 0; d < world_; ++d) {
        const int q = (rank_ + 1 + d) % world_;
        GOPF_CUDA(cudaMemcpy2DAsync(X_[q] + rank_ * block + (size_t)k1_begin * n_, sizeof(cplx) * plane,
                                    T + q * block + (size_t)k1_begin * n_, sizeof(cplx) * plane,
                                    sizeof(cplx) * k1_count * n_, m_, cudaMemcpyDefault, copy_stream_));
    }
}

void DistSolver::forward_mid_peer_planes(const cplx* W, int begin, int count, int max_ctas) {
    copy_after_compute();
    PassGeom g = make_geom(count, n_, n_, 1);
    g.peer = peer_out(Y_, false);
    g.peer.block_off += (long long)begin * g.peer.a_stride;
    g.peer.max_ctas = max_ctas;
    check(launch_pass(g, plan_->tx_want, plain_io(W + (size_t)begin * n_ * n_, nullptr, false, 1.0), plan_->twiddle(1),
                      copy_stream_),
          "forward axis 1 -> peers (planes)");
}

void DistSolver::exchange_join() {
    if (!copy_stream_) return;
    plan_->use_device();
    GOPF_CUDA(cudaEventRecord(ev_copy_, copy_stream_));
    GOPF_CUDA(cudaStreamWaitEvent(stream(), ev_copy_, 0));
}

void DistSolver::forward_local(cplx* W, cplx* send) {
    plan_->use_device();
    check_real(W);
    check(launch_pass(make_geom(m_, n_, n_, 2), plan_->tx_want, plain_io(W, W, false, 1.0), plan_->twiddle(2), stream()),
          "forward axis 2");
    forward_mid(W, send);
}

void DistSolver::forward_mid(const cplx* W, cplx* send) {
    plan_->use_device();
    check(launch_pass(slab_axis1(false, true), plan_->tx_want, plain_io(W, send, false, 1.0), plan_->twiddle(1), stream()),
          "forward axis 1");
}

void DistSolver::forward_finish(cplx* T) {
    plan_->use_device();
    check(launch_pass(make_geom(n_, m_, n_, 0), plan_->tx_want, plain_io(T, T, false, 1.0), plan_->twiddle(0), stream()),
          "forward axis 0");
}

void DistSolver::inverse_start(const cplx* S, cplx* T) {
    plan_->use_device();
    check(launch_pass(make_geom(n_, m_, n_, 0), plan_->tx_want, plain_io(S, T, true, 1.0), plan_->twiddle(0), stream()),
          "inverse axis 0");
}

void DistSolver::inverse_mid(const cplx* recv, cplx* W) {
    plan_->use_device();
    check(launch_pass(slab_axis1(true, false), plan_->tx_want, plain_io(recv, W, true, 1.0), plan_->twiddle(1), stream()),
          "inverse axis 1");
}

void DistSolver::real_step(cplx* W) {
    plan_->use_device();
    PassGeom g = make_geom(m_, n_, n_, 2);
    g.node0 = (long long)rank_ * m_ * n_ * n_;
    g.real_pairs = real_pairs();
    const double inv_n = 1.0 / ((double)n_ * n_ * n_);
    check(launch_fused_real(g, 0, W, nullptr, m_model_->derived[derived_].dev, inv_n, (unsigned long long)steps_taken_,
                            plan_->twiddle(2), stream()),
          "fused real-space kernel");
}

void DistSolver::inverse_finish(cplx* W, cplx* real_out) {
    plan_->use_device();
    const double inv_n = 1.0 / ((double)n_ * n_ * n_);
    check(launch_pass(make_geom(m_, n_, n_, 2), plan_->tx_want, plain_io(W, real_out, true, inv_n), plan_->twiddle(2),
                      stream()),
          "inverse axis 2");
}

void DistSolver::kspace_step(cplx* T, cplx* S) {
    plan_->use_device();
    FreqTabs ft;
    ft.f0 = plan_->freq_axis(0);
    ft.f1 = plan_->freq_axis(1);
    ft.f2 = plan_->freq_axis(2);
    ft.rank = 3;
    ft.off1 = rank_ * m_;
    check(launch_fused_kspace(make_geom(n_, m_, n_, 0), plan_->tx_want, T, T, S, prog_, ft, plan_->twiddle(0), stream()),
          "fused k-space kernel");
}

}  // namespace gopf
