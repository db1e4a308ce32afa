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
    prog_.filter = nullptr;
    prog_.filter_n = 0;
    finalize_single_field_program(&prog_, 1);
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
    return r;
}

PassGeom DistSolver::slab_axis1(bool split_in, bool split_out) const {
    PassGeom g = make_geom(m_, n_, n_, 1);
    if (split_in) g.in = split_map();
    if (split_out) g.out = split_map();
    return g;
}

void DistSolver::forward_local(cplx* W, cplx* send) {
    plan_->use_device();
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
    check(launch_fused_kspace(make_geom(n_, m_, n_, 0), plan_->tx_want, T, S, prog_, ft, plan_->twiddle(0), stream()),
          "fused k-space kernel");
}

}  // namespace gopf
