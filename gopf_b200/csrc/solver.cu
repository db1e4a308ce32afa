// Solver implementation (see solver.h).  Citations: /root/reference.
#include "solver.h"

#include <cmath>
#include <cstring>

#include "fused_launch.h"

namespace gopf {

// ---- small kernels -------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_update_generic(const __grid_constant__ DevKProgram P, SpectraPtrs sp, ImplicitTab tab, FreqGeom fg, long long n) {
    update_all(P, sp, tab, fg, n);
}

__global__ void __launch_bounds__(256)
    k_implicit_table(const __grid_constant__ DevKProgram P, int i, cplx* out, FreqGeom fg, long long n) {
    implicit_table_all(P, i, out, fg, n);
}

// filter / (1 - dt*den) of the single-field program, real part (finalize_single_field_program admits only real,
// time-independent implicit terms to the tabulated form)
__global__ void __launch_bounds__(256)
    k_fused_dtab(const __grid_constant__ DevKProgram P, double* out, FreqGeom fg, long long n) {
    const bool small = n < (1LL << 31);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq_fast(fg, idx, small, f);
        const KPoint kp = make_kpoint(f[0], f[1], f[2]);
        const DevEquation& q = P.eq[0];
        cplx den = mk(0.0, 0.0);
        for (int j = 0; j < q.n_den; ++j) den += eval_term(P, q.den[j], kp, [&](int) -> cplx { return mk(1.0, 0.0); });
        cplx r = cdiv(mk(1.0, 0.0), mk(1.0 - P.dt * den.x, -P.dt * den.y));
        if (P.filter) r = mk(r.x * filter_eval(P.filter, P.filter_n, kp.frad * 2.0 / GOPF_PI), 0.0);
        out[idx] = r.x;
    }
}

// real part of every cell; big-endian byte order on request (encoding/binary.BigEndian in
// pf/fileIO.go:85-95 SaveFloat64)
__global__ void k_real_part(const cplx* __restrict__ in, double* __restrict__ out, int big_endian, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = in[i].x;
        if (big_endian) {
            const unsigned long long u = (unsigned long long)__double_as_longlong(v);
            const unsigned lo = (unsigned)u, hi = (unsigned)(u >> 32);
            const unsigned long long sw =
                ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | (unsigned long long)__byte_perm(hi, 0, 0x0123);
            v = __longlong_as_double((long long)sw);
        }
        out[i] = v;
    }
}

// row-major [n0][n1][n2] <-> blocked [n0 >> s][n1][1 << s][n2] (solver.h): whole n2-lines move, 16 B per thread
__global__ void __launch_bounds__(256)
    k_relayout_blocked(const cplx* __restrict__ in, cplx* __restrict__ out, int n1, int n2, int s, int to_blocked, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long line = i / n2;
        const int i2 = (int)(i - line * n2);
        const long long i0 = line / n1;
        const int i1 = (int)(line - i0 * n1);
        const long long j = ((((i0 >> s) * n1 + i1) << s) + (i0 & ((1 << s) - 1))) * n2 + i2;
        if (to_blocked) out[j] = in[i];
        else out[i] = in[j];
    }
}

__global__ void k_scale(cplx* a, double s, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        a[i] = mk(a[i].x * s, a[i].y * s);
}

__global__ void __launch_bounds__(256)
    k_eval_derived(const __grid_constant__ DevDerived D, RealPtrs R, cplx* out, unsigned long long step, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = eval_derived(D, [&](int f) -> cplx { return R.r[f][i]; }, step, i);
}

__global__ void k_volume_lp_update(double* state, const cplx* field_spec, const cplx* indicator_spec, double dt) {
    volume_lp_update(state, field_spec, indicator_spec, dt);
}

// M(k) of the elastic term for every node (elastic.cuh); time-independent, tabulated once.
__global__ void __launch_bounds__(256)
    k_elastic_table(const __grid_constant__ ElastParams E, FreqGeom fg, double* out, long long n) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};  // homoLinElast.go:103-111: padded to 3
        ref_freq(fg, idx, f);
        out[idx] = elastic_multiplier(E, f[0], f[1], f[2]);
    }
}

__global__ void __launch_bounds__(256)
    k_rk4_rhs(const __grid_constant__ DevKProgram P, SpectraPtrs sp, SpectraPtrs kout, FreqGeom fg, long long n) {
    rk4_rhs_all(P, sp, kout, fg, n);
}

__global__ void __launch_bounds__(256)
    k_rk4_point(const __grid_constant__ DevKProgram P, int mode, double fdt, SpectraPtrs field, SpectraPtrs initial,
                SpectraPtrs final_, SpectraPtrs kf, FreqGeom fg, long long n) {
    rk4_point_all(P, mode, fdt, field, initial, final_, kf, fg, n);
}

// Go's cmplx.Pow leaves an O(1e-16) imaginary residue on negative real scalars (m1 = -1
// becomes -1 + 1.2e-16i).  The fused fast form works with real polynomial
// coefficients and drops a residue below 4 ulp of the real part; the general path keeps it.
static bool negligible_imag(const DevTerm& t) { return std::fabs(t.cim) <= 1e-15 * std::fabs(t.cre); }

static unsigned grid_for(long long n) {
    long long blocks = (n + 255) / 256;
    const long long cap = 148LL * 16;
    return (unsigned)(blocks < cap ? blocks : cap);
}

// ---- construction --------------------------------------------------------------------
Solver::Solver(Model* m, int rank, const int* n, double dt, int device) : m_(m), dt_(dt) {
    if (!m) throw Error("solver: model is NULL");
    if (rank != 2 && rank != 3)
        throw Error(strf("solver: rank must be 2 or 3 (got %d); FFTWWrapper.Freq indexes res[1] (fftWrap.go:61)", rank));
    m_->init();  // NewSolver calls m.Init() (solver.go:42)
    jit_on_ = jit::enabled();
    jit_inpass_ = jit::inpass_enabled();
    plan_.reset(new FftPlan(rank, n, device));
    for (const HostField& f : m_->fields)  // solver.go:55-60
        if (f.n != plan_->N) throw Error("solver: Inconsistent domain size and number of grid points");
    std::memset(&S_, 0, sizeof(S_));
    std::memset(&R_, 0, sizeof(R_));
    for (int i = 0; i < GOPF_MAX_FIELDS; ++i)
        Rw_[i] = rk_initial_[i] = rk_final_[i] = rk_k_[i] = implicit_tab_[i] = ie_orig_[i] = ie_rhs_prev_[i] = ie_res_[i] = nullptr;
    for (int i = 0; i < 3; ++i) sg_tmp_[i] = nullptr;
    for (int i = 0; i < GOPF_MAX_SPECIAL; ++i) {
        elast_mtab_[i] = nullptr;
        elast_phi_[i] = nullptr;
    }
    for (int i = 0; i < GOPF_MAX_SPECTRA; ++i) d_table_[i] = nullptr;
    decide_path();
    m_->attached_solvers++;
}

Solver::~Solver() {
    m_->attached_solvers--;
    cudaSetDevice(plan_->device);
    drop_graph();
    for (int i = 0; i < GOPF_MAX_SPECTRA; ++i)
        if (S_.s[i]) cudaFree(S_.s[i]);
    if (W2_) cudaFree(W2_);
    if (fused_dtab_) cudaFree(fused_dtab_);
    for (int i = 0; i < GOPF_MAX_FIELDS; ++i) {
        if (Rw_[i]) cudaFree(Rw_[i]);
        if (rk_initial_[i]) cudaFree(rk_initial_[i]);
        if (rk_final_[i]) cudaFree(rk_final_[i]);
        if (rk_k_[i]) cudaFree(rk_k_[i]);
        if (implicit_tab_[i]) cudaFree(implicit_tab_[i]);
        if (ie_orig_[i]) cudaFree(ie_orig_[i]);
        if (ie_rhs_prev_[i]) cudaFree(ie_rhs_prev_[i]);
        if (ie_res_[i]) cudaFree(ie_res_[i]);
    }
    for (double* v : ie_vec_)
        if (v) cudaFree(v);
    if (ie_partial_) cudaFree(ie_partial_);
    for (int i = 0; i < 3; ++i)
        if (sg_tmp_[i]) cudaFree(sg_tmp_[i]);
    for (int i = 0; i < GOPF_MAX_SPECIAL; ++i) {
        if (elast_mtab_[i]) cudaFree(elast_mtab_[i]);
        if (elast_phi_[i]) cudaFree(elast_phi_[i]);
    }
    for (int i = 0; i < GOPF_MAX_SPECTRA; ++i)
        if (d_table_[i]) cudaFree(d_table_[i]);
    for (jit::Kernel* k : jit_derived_) jit::unload(k);
    for (jit::Kernel* k : jit_pass_) jit::unload(k);
    jit::unload(jit_kupdate_);
    jit::unload(jit_rk4_rhs_);
    jit::unload(jit_rk4_point_);
    free_catalog_buffers();
    sdd_free_buffers();
    if (W_) cudaFree(W_);
    if (d_real_out_) cudaFree(d_real_out_);
    if (d_filter_) cudaFree(d_filter_);
    if (d_lp_state_) cudaFree(d_lp_state_);
    if (d_imag_max_) cudaFree(d_imag_max_);
    if (h_imag_max_) cudaFreeHost(h_imag_max_);
    for (KernelTimer& t : timers_)
        for (auto& ev : t.events) {
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
}

// ---- CUDA graph replay -------------------------------------------------------------------
// Below ~1 M cells the fused step is a handful of kernels of a few microseconds each and the
// host launch rate bounds it (128^2: 16.6 us per 2-kernel step).  One graph of GRAPH_STEPS steps
// is captured from the stream and replayed; anything that changes a kernel argument drops it.
static const int GRAPH_STEPS = 8;

void Solver::drop_graph() {
    if (graph_exec_) {
        cudaGraphExecDestroy(graph_exec_);
        graph_exec_ = nullptr;
    }
    graph_steps_ = 0;
}

// true when some term draws white noise at the k-point (TK_WHITE_NOISE_K, step_program.h)
bool program_has_knoise(const DevKProgram& P) {
    for (int i = 0; i < P.n_fields; ++i)
        for (int j = 0; j < P.eq[i].n_rhs; ++j)
            if (P.eq[i].rhs[j].kind == TK_WHITE_NOISE_K) return true;
    return false;
}

// the step counter of the k-space noise stream lives in the program (TensorHessianParams::K[2])
static void stamp_noise_step(DevKProgram* P, unsigned long long step) {
    for (int i = 0; i < P->n_fields; ++i)
        for (int j = 0; j < P->eq[i].n_rhs; ++j)
            if (P->eq[i].rhs[j].kind == TK_WHITE_NOISE_K) P->th[P->eq[i].rhs[j].param].K[2] = gopf_double_of(step);
}

bool Solver::graph_applicable() const {
    if (graph_disabled_ || !fused_ || profiling_ || stepper_ != StepperKind::Euler || !w_valid_ || blocked_) return false;
    if (has_knoise_) return false;  // the step counter is part of the kernel arguments
    if (plan_->N > (size_t)1 << 20) return false;
    const int k = m_->derived[fused_derived_].dev.kind;
    return k == DK_MONOMIAL || k == DK_RPN;  // noise / table fields take the step number as a kernel argument
}

bool Solver::build_graph(int steps) {
    cudaStream_t s = stream();
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        graph_disabled_ = true;
        return false;
    }
    bool ok = true;
    const long long launches_before = launches_;
    try {
        for (int i = 0; i < steps; ++i) euler_step_fused();
    } catch (...) {
        ok = false;
    }
    launches_ = launches_before;  // capturing launches nothing
    if (cudaStreamEndCapture(s, &graph) != cudaSuccess || !graph) ok = false;
    if (ok && cudaGraphInstantiate(&graph_exec_, graph, 0) != cudaSuccess) ok = false;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
        cudaGetLastError();
        graph_exec_ = nullptr;
        graph_disabled_ = true;
        return false;
    }
    graph_steps_ = steps;
    return true;
}

void Solver::synchronize() {
    plan_->use_device();
    GOPF_CUDA(cudaStreamSynchronize(stream()));
}

void Solver::set_stepper(const std::string& name) {
    leave_blocked();
    if (name == "euler") stepper_ = StepperKind::Euler;
    else if (name == "rk4") stepper_ = StepperKind::RK4;
    // not a SetStepper name in the reference: there the user assigns solver.Stepper = &pf.ImplicitEuler{...}
    // (pf/implicitEuler_test.go:193-198); the C ABI has no struct to assign, so the name selects it
    else if (name == "implicit_euler") stepper_ = StepperKind::ImplicitEuler;
    // likewise solver.Stepper = &sdd (pf/sdd_test.go:86); settings through sdd_set / sdd_set_orientation
    else if (name == "sdd") {
        stepper_ = StepperKind::SDD;
        sdd_ = SddState();  // NewSDD (sdd.go:120-129)
    }
    else throw Error("Unknown stepper scheme");  // solver.go:101
    current_step_ = 0;  // SetStepper builds a fresh stepper struct (solver.go:90-99)
    decide_path();
}

void Solver::force_generic(bool on) {
    leave_blocked();
    allow_fused_ = !on;
    decide_path();
}

void Solver::set_filter(const double* table, int n) {
    if (stepper_ == StepperKind::SDD && table && n > 1) throw Error("SDD: Does not support modal filters");  // sdd.go:431-433
    plan_->use_device();
    if (d_filter_) {
        cudaFree(d_filter_);
        d_filter_ = nullptr;
    }
    filter_n_ = 0;
    if (table && n > 1) {
        GOPF_CUDA(cudaMalloc(&d_filter_, sizeof(double) * n));
        GOPF_CUDA(cudaMemcpy(d_filter_, table, sizeof(double) * n, cudaMemcpyHostToDevice));
        filter_n_ = n;
    }
    prog_dirty_ = true;
}

// which derived fields are transformed, and whether the single-field fused path applies
void Solver::decide_path() {
    leave_blocked();
    block_log_ = -1;
    fused_ = false;
    fused_derived_ = -1;
    w_valid_ = false;
    const int F = (int)m_->fields.size();
    int n_used = 0, used = -1;
    for (size_t d = 0; d < m_->derived.size(); ++d)
        if (m_->derived[d].used) { n_used++; used = (int)d; }
    bool ok = allow_fused_ && stepper_ == StepperKind::Euler && F == 1 && n_used == 1 && m_->n_work_spectra == 0 &&
              plan_->freq_axis_consistent() && m_->compiled.size() == 1;
    for (int ax = 0; ax < 3 && ok; ++ax) {
        if (plan_->extent(ax) <= 1) continue;
        if (!plan_->axis_fast(ax) || !fused_length_supported(plan_->extent(ax))) ok = false;
    }
    if (ok && plan_->rank == 2 && (plan_->n1 < 2 || plan_->n2 < 2)) ok = false;
    if (ok) {
        // VolumeConservingLP reads the indicator spectrum after the step; keep that on the generic path
        for (const auto& kv : m_->user_terms)
            if (kv.second.kind == UserTermKind::VolumeConservingLP || kv.second.kind == UserTermKind::ConservativeNoise)
                ok = false;
    }
    fused_ = ok;
    fused_derived_ = ok ? used : -1;
    prog_dirty_ = true;
}

void Solver::ensure_buffers() {
    plan_->use_device();
    const int F = (int)m_->fields.size();
    const size_t bytes = sizeof(cplx) * plan_->N;
    for (int i = 0; i < F; ++i)
        if (!S_.s[i]) GOPF_CUDA(cudaMalloc(&S_.s[i], bytes));
    if (fused_) {
        if (!W_) GOPF_CUDA(cudaMalloc(&W_, bytes));
    } else {
        bool need_real = m_->n_work_spectra > 0 || stepper_ == StepperKind::ImplicitEuler;
        for (size_t d = 0; d < m_->derived.size(); ++d) {
            if (!m_->derived[d].used) continue;
            need_real = true;
            if (!S_.s[F + d]) GOPF_CUDA(cudaMalloc(&S_.s[F + d], bytes));
        }
        for (int w = 0; w < m_->n_work_spectra; ++w) {
            const int si = F + (int)m_->derived.size() + w;
            if (!S_.s[si]) GOPF_CUDA(cudaMalloc(&S_.s[si], bytes));
        }
        bool any_sg = false;
        for (const auto& kv : m_->user_terms) any_sg |= kv.second.kind == UserTermKind::SquaredGradient;
        if (m_->n_work_spectra > 0)
            for (int d = 0; d < (any_sg ? plan_->rank : 1); ++d)
                if (!sg_tmp_[d]) GOPF_CUDA(cudaMalloc(&sg_tmp_[d], bytes));
        for (const auto& kv : m_->user_terms) {
            const UserTerm& u = kv.second;
            if (u.kind != UserTermKind::HomogeneousModulusLinElast) continue;
            if (!elast_mtab_[u.slot]) {
                GOPF_CUDA(cudaMalloc(&elast_mtab_[u.slot], sizeof(double) * plan_->N));
                ElastParams E;
                make_elast_params(&E, u.stiffness, u.misfit, plan_->rank);
                k_elastic_table<<<grid_for((long long)plan_->N), 256, 0, stream()>>>(E, plan_->freq_geom(),
                                                                                     elast_mtab_[u.slot], (long long)plan_->N);
                GOPF_CUDA(cudaGetLastError());
                launches_++;
            }
            if (stepper_ != StepperKind::Euler && !elast_phi_[u.slot]) {
                GOPF_CUDA(cudaMalloc(&elast_phi_[u.slot], bytes));
                GOPF_CUDA(cudaMemsetAsync(elast_phi_[u.slot], 0, bytes, stream()));
            }
        }
        if (need_real)
            for (int i = 0; i < F; ++i)
                if (!Rw_[i]) {
                    GOPF_CUDA(cudaMalloc(&Rw_[i], bytes));
                    R_.r[i] = Rw_[i];
                }
    }
    if (stepper_ == StepperKind::RK4)
        for (int i = 0; i < F; ++i) {
            if (!rk_initial_[i]) GOPF_CUDA(cudaMalloc(&rk_initial_[i], bytes));
            if (!rk_final_[i]) GOPF_CUDA(cudaMalloc(&rk_final_[i], bytes));
            if (!rk_k_[i]) GOPF_CUDA(cudaMalloc(&rk_k_[i], bytes));
        }
    // prescribed (table) derived fields: upload once, patch the device pointer
    for (size_t d = 0; d < m_->derived.size(); ++d) {
        DerivedSpec& ds = m_->derived[d];
        if (ds.origin != DerivedOrigin::Table || d_table_[d]) continue;
        GOPF_CUDA(cudaMalloc(&d_table_[d], sizeof(double) * ds.table.size()));
        GOPF_CUDA(cudaMemcpy(d_table_[d], ds.table.data(), sizeof(double) * ds.table.size(), cudaMemcpyHostToDevice));
        ds.dev.table = d_table_[d];
        ds.dev.table_n = (long long)plan_->N;
    }
    // download target
    for (int i = 0; i < F; ++i)
        if (!Rw_[i]) {
            GOPF_CUDA(cudaMalloc(&Rw_[i], bytes));
            R_.r[i] = Rw_[i];
        }
    if (!d_lp_state_) {
        GOPF_CUDA(cudaMalloc(&d_lp_state_, sizeof(double) * 3 * GOPF_MAX_SPECIAL));
        double init[3 * GOPF_MAX_SPECIAL];
        for (int i = 0; i < GOPF_MAX_SPECIAL; ++i) {
            init[3 * i + 0] = 0.0;  // Multiplier
            init[3 * i + 1] = 0.0;  // CurrentIntegral
            init[3 * i + 2] = 1.0;  // IsFirstUpdate
        }
        GOPF_CUDA(cudaMemcpy(d_lp_state_, init, sizeof(init), cudaMemcpyHostToDevice));
    }
}

void Solver::rebuild_program() {
    if (!prog_dirty_) return;
    drop_graph();
    m_->fill_program(&prog_, dt_, plan_->rank);
    prog_.filter = d_filter_;
    prog_.filter_n = filter_n_;
    for (int i = 0; i < GOPF_MAX_SPECIAL; ++i) prog_.lp_multiplier[i] = d_lp_state_ ? d_lp_state_ + 3 * i : nullptr;
    has_knoise_ = program_has_knoise(prog_);
#ifndef GOPF_KNOISE
    if (has_knoise_)
        throw Error("solver: the model asks for k-space noise (gopf_model_set_kspace_noise) but this library's device "
                    "code was built without -DGOPF_KNOISE");
#endif
    // the pairing of k and -k follows the Freq components, which match the transform's layout only
    // for 2-D and cubic 3-D grids (fft_plan.h freq_axis_consistent)
    if (has_knoise_ && !plan_->freq_axis_consistent())
        throw Error("solver: k-space noise needs a 2-D or cubic 3-D grid");
    fused_prog_ = prog_;
    if (fused_) {
        finalize_single_field_program(&fused_prog_, (int)m_->fields.size(), env_int("GOPF_FUSED_TABLE", 1) != 0);
        if (fused_prog_.fast == 2) {
            // tabulated form: filter / (1 - dt*den) for every k-point, in the spectrum's (row-major) order
            const long long n = (long long)plan_->N;
            if (!fused_dtab_) GOPF_CUDA(cudaMalloc(&fused_dtab_, sizeof(double) * n));
            k_fused_dtab<<<grid_for(n), 256, 0, stream()>>>(fused_prog_, fused_dtab_, plan_->freq_geom(), n);
            GOPF_CUDA(cudaGetLastError());
            launches_++;
            fused_prog_.dtab = fused_dtab_;
        }
    }
    prog_dirty_ = false;
    implicit_tab_dirty_ = true;
}

// Program as seen by the fused single-field kernels: spectrum 0 = the field, 1 = the derived
// field; plus the real-polynomial fast form when every term is a monomial with a real
// coefficient, degree <= 4, no self term on the explicit side and no filter.
void finalize_single_field_program(DevKProgram* prog, int n_fields, bool allow_table) {
    DevKProgram& P = *prog;
    DevEquation& q = P.eq[0];
    for (int j = 0; j < q.n_rhs; ++j)
        if (q.rhs[j].brick >= n_fields) q.rhs[j].brick = 1;
    for (int j = 0; j < q.n_den; ++j)
        if (q.den[j].brick >= n_fields) q.den[j].brick = 1;
    bool rhs_poly = true, den_poly = true, den_real = true;
    double p_noise[GOPF_MAX_POLY + 1];
    for (int i = 0; i <= GOPF_MAX_POLY; ++i) P.p_nl[i] = P.p_self[i] = P.q[i] = p_noise[i] = 0.0;
    P.deg_nl = 0;
    P.deg_self = -1;
    P.deg_q = 0;
    P.noise_param = -1;
    P.dtab = nullptr;
    int n_noise = 0, deg_noise = 0;
    for (int j = 0; j < q.n_rhs && rhs_poly; ++j) {
        const DevTerm& t = q.rhs[j];
        if (t.kind == TK_WHITE_NOISE_K && negligible_imag(t) && t.lap <= GOPF_MAX_POLY) {
            n_noise++;
            P.noise_param = t.param;
            p_noise[t.lap] += t.cre;
            if (t.lap > deg_noise) deg_noise = t.lap;
            continue;
        }
        if (t.kind != TK_MONOMIAL || !negligible_imag(t) || t.lap > GOPF_MAX_POLY || t.brick < 0) { rhs_poly = false; break; }
        if (t.brick == 1) {
            P.p_nl[t.lap] += t.cre;
            if (t.lap > P.deg_nl) P.deg_nl = t.lap;
        } else {
            P.p_self[t.lap] += t.cre;
            if (t.lap > P.deg_self) P.deg_self = t.lap;
        }
    }
    for (int j = 0; j < q.n_den; ++j) {
        const DevTerm& t = q.den[j];
        // real-valued, time-independent multipliers (step_program.h eval_term)
        const bool real_kind = (t.kind == TK_MONOMIAL || t.kind == TK_PAIR_CORR || t.kind == TK_SPECTRAL_VISC ||
                                t.kind == TK_TENSOR_HESSIAN) && t.brick < 0 && negligible_imag(t);
        if (!real_kind) den_real = false;
        if (t.kind != TK_MONOMIAL || !negligible_imag(t) || t.lap > GOPF_MAX_POLY || t.brick >= 0) { den_poly = false; continue; }
        P.q[t.lap] += t.cre;
        if (t.lap > P.deg_q) P.deg_q = t.lap;
    }
    const bool fast1 = rhs_poly && den_poly && n_noise == 0 && P.deg_nl <= 4 && P.deg_q <= 4 && P.deg_self < 0 &&
                       P.filter == nullptr;
    const bool fast2 = !fast1 && allow_table && rhs_poly && den_real && n_noise <= 1 && P.deg_nl <= 4 && P.deg_self <= 4 &&
                       deg_noise <= 4;
    P.fast = fast1 ? 1 : (fast2 ? 2 : 0);
    if (!fast2) P.noise_param = -1;
    for (int i = 0; i < 5; ++i) {
        P.fa[i] = P.dt * P.p_nl[i];
        P.fq[i] = (i == 0 ? 1.0 : 0.0) - P.dt * P.q[i];
        P.fself[i] = (i == 0 ? 1.0 : 0.0) + P.dt * P.p_self[i];
        P.fnz[i] = P.dt * p_noise[i];
    }
}

// The single-field fused path applies to: one field, one equation, exactly one derived field
// in use, no work-spectrum terms, no term with a per-step device hook.  Returns the derived
// field's index or -1.
int single_field_derived_index(const Model& m) {
    if (m.fields.size() != 1 || m.compiled.size() != 1 || m.n_work_spectra != 0) return -1;
    int n_used = 0, used = -1;
    for (size_t d = 0; d < m.derived.size(); ++d)
        if (m.derived[d].used) { n_used++; used = (int)d; }
    if (n_used != 1) return -1;
    for (const auto& kv : m.user_terms)
        if (kv.second.kind == UserTermKind::VolumeConservingLP || kv.second.kind == UserTermKind::ConservativeNoise)
            return -1;
    return used;
}

void Solver::fused_form(int* form, int* derived_form) {
    *form = *derived_form = 0;
    if (!fused_) return;
    plan_->use_device();
    rebuild_program();
    *form = fused_prog_.fast;
    const DevDerived& D = m_->derived[fused_derived_].dev;
    if (D.kind == DK_MONOMIAL && D.n_factors == 1 && D.ipower[0] >= 0 && D.ipower[0] <= 15) *derived_form = 1;
    else if (D.kind == DK_RPN && D.poly_deg >= 0) *derived_form = 2;
}

FreqTabs Solver::freq_tabs() const {
    FreqTabs ft;
    ft.f0 = plan_->freq_axis(0);
    ft.f1 = plan_->freq_axis(1);
    ft.f2 = plan_->freq_axis(2);
    ft.rank = plan_->rank;
    ft.off1 = 0;
    return ft;
}

// ---- profiling -----------------------------------------------------------------------
void Solver::set_profiling(bool on) {
    profiling_ = on;
    if (on) {
        for (KernelTimer& t : timers_) {
            for (auto& ev : t.events) {
                cudaEventDestroy(ev.first);
                cudaEventDestroy(ev.second);
            }
            t.events.clear();
            t.total_ms = 0.0;
            t.launches = 0;
        }
    }
}

int Solver::tick(const char* name, double bytes) {
    launches_++;
    if (!profiling_) return -1;
    int id = -1;
    for (size_t i = 0; i < timers_.size(); ++i)
        if (timers_[i].name == name) id = (int)i;
    if (id < 0) {
        KernelTimer t;
        t.name = name;
        timers_.push_back(t);
        id = (int)timers_.size() - 1;
    }
    timers_[id].bytes_per_launch = bytes;
    cudaEvent_t a, b;
    GOPF_CUDA(cudaEventCreate(&a));
    GOPF_CUDA(cudaEventCreate(&b));
    timers_[id].events.push_back({a, b});
    GOPF_CUDA(cudaEventRecord(a, stream()));
    return id;
}

void Solver::tock(int id) {
    if (id < 0) return;
    GOPF_CUDA(cudaEventRecord(timers_[id].events.back().second, stream()));
}

const std::vector<KernelTimer>& Solver::collect_profile() {
    synchronize();
    for (KernelTimer& t : timers_) {
        for (auto& ev : t.events) {
            float ms = 0.f;
            GOPF_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
            t.total_ms += ms;
            t.launches++;
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
        t.events.clear();
    }
    return timers_;
}

// ---- transforms ----------------------------------------------------------------------
void Solver::inverse_to_real(const cplx* spec, cplx* out) {
    cudaStream_t s = stream();
    const double inv_n = 1.0 / (double)plan_->N;
    bool all_fast = true;
    int n_active = 0, last_axis = -1;
    for (int ax = 0; ax < 3; ++ax)
        if (plan_->extent(ax) > 1) {
            n_active++;
            last_axis = ax;
            if (!plan_->axis_fast(ax)) all_fast = false;
        }
    const double cell = 32.0 * (double)plan_->N;
    if (!all_fast || n_active == 0) {
        if (out != spec) GOPF_CUDA(cudaMemcpyAsync(out, spec, sizeof(cplx) * plan_->N, cudaMemcpyDeviceToDevice, s));
        plan_->exec_device(out, +1, s);
        k_scale<<<grid_for((long long)plan_->N), 256, 0, s>>>(out, inv_n, (long long)plan_->N);
        GOPF_CUDA(cudaGetLastError());
        launches_ += 2;
        return;
    }
    bool first = true;
    for (int ax = 0; ax < 3; ++ax) {
        if (plan_->extent(ax) <= 1) continue;
        const PassGeom g = plan_->geom(ax);
        const PassIO io = plain_io(first ? spec : out, out, true, ax == last_axis ? inv_n : 1.0);
        const int id = tick("pass_inverse", cell);
        cudaError_t e = launch_pass(g, plan_->tx_want, io, plan_->twiddle(ax), s);
        tock(id);
        if (e != cudaSuccess) throw Error(strf("inverse pass axis %d: %s", ax, cudaGetErrorString(e)));
        first = false;
    }
}

void Solver::forward_in_place(cplx* data) { plan_->exec_device(data, -1, stream()); }

// A registered function (or noise / table) evaluated at every node into `out`.  Registered
// functions run as NVRTC-compiled straight-line kernels when the specialisation is on and
// compiles; otherwise (and for every other kind) the RPN interpreter kernel does the work.
void Solver::derived_pointwise(int d, cplx* out, unsigned long long step_no, cudaStream_t s) {
    const DevDerived& dd = m_->derived[d].dev;
    const long long n = (long long)plan_->N;
    if (jit_on_ && dd.kind == DK_RPN) {
        if (jit_derived_.size() != m_->derived.size()) {
            for (jit::Kernel* k : jit_derived_) jit::unload(k);
            jit_derived_.assign(m_->derived.size(), nullptr);
            jit_tried_.assign(m_->derived.size(), 0);
        }
        if (!jit_tried_[d]) {
            jit_tried_[d] = 1;
            std::string log;
            std::vector<char> cubin;
            try {
                if (jit::compile_cubin(jit::derived_kernel_source(dd, nullptr), &cubin, &log))
                    jit_derived_[d] = jit::load(cubin, "gopf_jit_derived", &log);
            } catch (const std::exception& e) {
                log = e.what();
            }
            if (!jit_derived_[d]) jit_log_ += "derived '" + m_->derived[d].name + "': " + log + "\n";
        }
        if (jit_derived_[d]) {
            const cplx* f[GOPF_MAX_FIELDS];
            for (int i = 0; i < GOPF_MAX_FIELDS; ++i) f[i] = R_.r[i];
            long long nn = n;
            void* args[] = {&f[0], &f[1], &f[2], &f[3], &out, &nn};
            std::string log;
            if (!jit::launch(jit_derived_[d], grid_for((n + 1) / 2), 256, args, s, &log))
                throw Error("jit launch of derived '" + m_->derived[d].name + "': " + log);
            return;
        }
    }
    k_eval_derived<<<grid_for(n), 256, 0, s>>>(dd, R_, out, step_no, n);
    GOPF_CUDA(cudaGetLastError());
}

int Solver::jit_kernels() const {
    int c = jit_kupdate_ ? 1 : 0;
    for (const jit::Kernel* k : jit_derived_)
        if (k) c++;
    for (const jit::Kernel* k : jit_pass_)
        if (k) c++;
    return c;
}

void Solver::forward_derived(int d) {
    cudaStream_t s = stream();
    const int F = (int)m_->fields.size();
    cplx* out = S_.s[F + d];
    bool all_fast = true;
    int first_axis = -1;
    for (int ax = 2; ax >= 0; --ax)
        if (plan_->extent(ax) > 1) {
            if (first_axis < 0) first_axis = ax;
            if (!plan_->axis_fast(ax)) all_fast = false;
        }
    const double cell = 32.0 * (double)plan_->N;
    const unsigned long long step_no = (unsigned long long)steps_taken_;
    if (!all_fast || first_axis < 0) {
        derived_pointwise(d, out, step_no, s);
        launches_++;
        plan_->exec_device(out, -1, s);
        return;
    }
    // Interpreted derived fields (registered functions, noise) are instruction-bound inside a pass
    // (one CTA holds few warps; measured 0.76 TB/s at 512^3): evaluate them in a pointwise kernel
    // at full occupancy and transform the result with plain passes (measured: see DESIGN.md 4.4).
    // With GOPF_JIT_INPASS a registered function gets its own copy of the contiguous pass with the
    // function compiled into the load (jit.h), which needs neither.
    const DevDerived& dd = m_->derived[d].dev;
    const bool in_pass = dd.kind == DK_MONOMIAL || dd.kind == DK_TABLE;
    jit::Kernel* jit_pass = nullptr;
    if (!in_pass && jit_on_ && jit_inpass_ && dd.kind == DK_RPN && plan_->geom(first_axis).B == 1)
        jit_pass = jit_pass_kernel(d, plan_->geom(first_axis).N);
    if (!in_pass && !jit_pass) {
        const int F0 = (int)m_->fields.size();
        const int id = tick("derived_pointwise", 16.0 * (double)plan_->N * (F0 + 1));
        derived_pointwise(d, out, step_no, s);
        tock(id);
    }
    for (int ax = 2; ax >= 0; --ax) {
        if (plan_->extent(ax) <= 1) continue;
        const PassGeom g = plan_->geom(ax);
        PassIO io = plain_io(out, out, false, 1.0);
        const bool fused_load = ax == first_axis && (in_pass || jit_pass);
        if (fused_load) {
            io.load_kind = LK_DERIVED;
            io.D = m_->derived[d].dev;
            io.R = R_;
            io.step = step_no;
        }
        const int id = tick(fused_load ? "pass_forward_derived" : "pass_forward", cell);
        if (fused_load && jit_pass) {
            unsigned grid = 0, block = 0;
            size_t smem = 0;
            if (!contig_launch_config(g.N, g.A, &grid, &block, &smem)) throw Error("jit pass: unsupported line length");
            PassGeom gl = g;
            gl.pf_tiles = 0;
            const cplx* tw = plan_->twiddle(ax);
            void* args[] = {&gl, &io, &tw};
            std::string log;
            if (!jit::launch(jit_pass, grid, block, args, s, &log, smem))
                throw Error("jit launch of the forward pass of derived '" + m_->derived[d].name + "': " + log);
        } else {
            cudaError_t e = launch_pass(g, plan_->tx_want, io, plan_->twiddle(ax), s);
            if (e != cudaSuccess) throw Error(strf("derived forward pass axis %d: %s", ax, cudaGetErrorString(e)));
        }
        tock(id);
    }
}

// k_pass_contig<N> with registered function d compiled into its load; NULL: not available
jit::Kernel* Solver::jit_pass_kernel(int d, int N) {
    if (jit_pass_.size() != m_->derived.size()) {
        for (jit::Kernel* k : jit_pass_) jit::unload(k);
        jit_pass_.assign(m_->derived.size(), nullptr);
        jit_pass_tried_.assign(m_->derived.size(), 0);
    }
    if (!jit_pass_tried_[d]) {
        jit_pass_tried_[d] = 1;
        std::string log, name_expr, lowered;
        std::vector<char> cubin;
        try {
            const std::string src = jit::derived_pass_source(m_->derived[d].dev, N, &name_expr);
            if (jit::compile_cubin(src, &cubin, &log, &name_expr, &lowered)) jit_pass_[d] = jit::load(cubin, lowered.c_str(), &log);
        } catch (const std::exception& e) {
            log = e.what();
        }
        if (!jit_pass_[d]) jit_log_ += "forward pass of derived '" + m_->derived[d].name + "': " + log + "\n";
    }
    return jit_pass_[d];
}

void Solver::eval_real_fields() {
    const int F = (int)m_->fields.size();
    for (int i = 0; i < F; ++i) inverse_to_real(S_.s[i], Rw_[i]);
}

// SquaredGradient (pf/squareGradientTerm.go:38-65): dim inverse transforms of the gradient
// components, squares summed in real space, ONE forward transform.
void Solver::squared_gradient_terms() {
    if (m_->n_work_spectra == 0) return;
    cudaStream_t s = stream();
    const double cell = 32.0 * (double)plan_->N;
    for (const auto& kv : m_->user_terms) {
        const UserTerm& u = kv.second;
        if (u.kind != UserTermKind::SquaredGradient) continue;
        const int fi = m_->field_index(u.field);
        if (fi < 0) throw Error("SquaredGradient: unknown field " + u.field);
        for (size_t e = 0; e < m_->compiled.size(); ++e)
            for (const DevTerm& t : m_->compiled[e].rhs)
                if (t.brick == u.work_spectrum && (int)e > fi)
                    throw Error("SquaredGradient of field '" + u.field +
                                "' used in a later equation than its own: the reference would read the already "
                                "updated spectrum (euler.go:27-39); this ordering is not supported on the device");
        bool all_fast = true;
        for (int ax = 0; ax < 3; ++ax)
            if (plan_->extent(ax) > 1 && !plan_->axis_fast(ax)) all_fast = false;
        if (!all_fast) throw Error("SquaredGradient needs power-of-two extents on the device path");
        const double inv_n = 1.0 / (double)plan_->N;
        int last_axis = -1;
        for (int ax = 0; ax < 3; ++ax)
            if (plan_->extent(ax) > 1) last_axis = ax;
        for (int d = 0; d < plan_->rank; ++d) {
            bool first = true;
            for (int ax = 0; ax < 3; ++ax) {
                if (plan_->extent(ax) <= 1) continue;
                const PassGeom g = plan_->geom(ax);
                PassIO io = plain_io(first ? S_.s[fi] : sg_tmp_[d], sg_tmp_[d], true, ax == last_axis ? inv_n : 1.0);
                if (plan_->freq_axis_consistent()) {
                    // the multiplier of component d sits in the pass along that component's axis (LK_GRADIENT_LINE)
                    if (ax == plan_->axis_of_component(d)) {
                        io.load_kind = LK_GRADIENT_LINE;
                        io.rtab = plan_->freq_axis(ax);
                    }
                } else if (first) {
                    io.load_kind = LK_GRADIENT;
                    io.fg = plan_->freq_geom();
                    io.comp = d;
                }
                const int id = tick("pass_inverse_gradient", cell);
                cudaError_t e = launch_pass(g, plan_->tx_want, io, plan_->twiddle(ax), s);
                tock(id);
                if (e != cudaSuccess) throw Error(strf("gradient pass: %s", cudaGetErrorString(e)));
                first = false;
            }
        }
        cplx* out = S_.s[u.work_spectrum];
        bool first = true;
        for (int ax = 2; ax >= 0; --ax) {
            if (plan_->extent(ax) <= 1) continue;
            const PassGeom g = plan_->geom(ax);
            PassIO io = plain_io(out, out, false, 1.0);
            if (first) {
                io.load_kind = LK_SUM_SQUARES;
                io.dim = plan_->rank;
                io.g[0] = sg_tmp_[0];
                io.g[1] = sg_tmp_[1];
                io.g[2] = plan_->rank > 2 ? sg_tmp_[2] : sg_tmp_[1];
            }
            const int id = tick("pass_forward_gradsq", cell);
            cudaError_t e = launch_pass(g, plan_->tx_want, io, plan_->twiddle(ax), s);
            tock(id);
            if (e != cudaSuccess) throw Error(strf("gradient-square forward pass: %s", cudaGetErrorString(e)));
            first = false;
        }
    }
}

bool Solver::has_elastic() const {
    for (const auto& kv : m_->user_terms)
        if (kv.second.kind == UserTermKind::HomogeneousModulusLinElast) return true;
    return false;
}

// HomogeneousModulusLinElast (pf/homoLinElast.go:47-99) with the dim(dim+1)/2 strain round trips
// collapsed by linearity (elastic.cuh): forward of H(phi), inverse of M(k) H^, forward of
// H'(phi) e - 2 E H(phi) H'(phi).  phi is the term's own real-space copy of the field: zeros
// before the first OnStepFinished (:145), the field at the end of the previous step afterwards.
void Solver::elastic_terms() {
    cudaStream_t s = stream();
    const double cell = 32.0 * (double)plan_->N;
    for (const auto& kv : m_->user_terms) {
        const UserTerm& u = kv.second;
        if (u.kind != UserTermKind::HomogeneousModulusLinElast) continue;
        const int fi = m_->field_index(u.field);
        if (fi < 0) throw Error("HomogeneousModulusLinElast: unknown field " + u.field);
        cplx* out = S_.s[u.work_spectrum];
        if (!elast_valid_) {  // H(0) = H'(0) = 0: the whole term is zero
            GOPF_CUDA(cudaMemsetAsync(out, 0, sizeof(cplx) * plan_->N, s));
            continue;
        }
        for (int ax = 0; ax < 3; ++ax)
            if (plan_->extent(ax) > 1 && !plan_->axis_fast(ax))
                throw Error("HomogeneousModulusLinElast needs power-of-two extents on the device path");
        const cplx* phi = stepper_ != StepperKind::Euler ? elast_phi_[u.slot] : Rw_[fi];
        ElastParams E;
        const double e_density = make_elast_params(&E, u.stiffness, u.misfit, plan_->rank);
        int first_fwd = -1, last_inv = -1;
        for (int ax = 0; ax < 3; ++ax)
            if (plan_->extent(ax) > 1) {
                if (first_fwd < 0 || ax > first_fwd) first_fwd = ax;
                last_inv = ax;
            }
        const double inv_n = 1.0 / (double)plan_->N;
        // FFT(H(phi)) -> out
        for (int ax = 2; ax >= 0; --ax) {
            if (plan_->extent(ax) <= 1) continue;
            PassIO io = plain_io(out, out, false, 1.0);
            if (ax == first_fwd) {
                io.load_kind = LK_ELAST_H;
                io.g[0] = io.g[1] = io.g[2] = phi;
            }
            const int id = tick("pass_forward_elast_h", cell);
            cudaError_t e = launch_pass(plan_->geom(ax), plan_->tx_want, io, plan_->twiddle(ax), s);
            tock(id);
            if (e != cudaSuccess) throw Error(strf("elastic indicator pass: %s", cudaGetErrorString(e)));
        }
        // IFFT(M(k) H^)/N -> sg_tmp_[0]
        bool first = true;
        for (int ax = 0; ax < 3; ++ax) {
            if (plan_->extent(ax) <= 1) continue;
            PassIO io = plain_io(first ? out : sg_tmp_[0], sg_tmp_[0], true, ax == last_inv ? inv_n : 1.0);
            if (first) {
                io.load_kind = LK_MUL_TABLE;
                io.rtab = elast_mtab_[u.slot];
            }
            const int id = tick("pass_inverse_elast_strain", cell);
            cudaError_t e = launch_pass(plan_->geom(ax), plan_->tx_want, io, plan_->twiddle(ax), s);
            tock(id);
            if (e != cudaSuccess) throw Error(strf("elastic strain pass: %s", cudaGetErrorString(e)));
            first = false;
        }
        // FFT(H'(phi) e - 2 E H(phi) H'(phi)) -> out
        for (int ax = 2; ax >= 0; --ax) {
            if (plan_->extent(ax) <= 1) continue;
            PassIO io = plain_io(ax == first_fwd ? sg_tmp_[0] : out, out, false, 1.0);
            if (ax == first_fwd) {
                io.load_kind = LK_ELAST_R;
                io.g[0] = io.g[1] = io.g[2] = phi;
                io.aux = 2.0 * e_density;
            }
            const int id = tick("pass_forward_elast_force", cell);
            cudaError_t e = launch_pass(plan_->geom(ax), plan_->tx_want, io, plan_->twiddle(ax), s);
            tock(id);
            if (e != cudaSuccess) throw Error(strf("elastic driving-force pass: %s", cudaGetErrorString(e)));
        }
    }
}

// HomogeneousModulusLinElast.OnStepFinished (pf/homoLinElast.go:130-134).  Under Euler the
// copy is the real-space field the next step computes anyway; RK4 evaluates the term at
// intermediate stages against the end-of-step snapshot, so it is taken here.
void Solver::elastic_hooks() {
    if (!has_elastic()) return;
    if (stepper_ != StepperKind::Euler)
        for (const auto& kv : m_->user_terms) {
            const UserTerm& u = kv.second;
            if (u.kind != UserTermKind::HomogeneousModulusLinElast) continue;
            inverse_to_real(S_.s[m_->field_index(u.field)], elast_phi_[u.slot]);
        }
    elast_valid_ = true;
}

void Solver::volume_lp_hooks() {
    for (const auto& kv : m_->user_terms) {
        const UserTerm& u = kv.second;
        if (u.kind != UserTermKind::VolumeConservingLP) continue;
        const int fi = m_->field_index(u.field);
        const int ii = m_->spectrum_index(u.indicator);
        if (fi < 0 || ii < 0) throw Error("VolumeConservingLP: unknown field or indicator");
        k_volume_lp_update<<<1, 32, 0, stream()>>>(d_lp_state_ + 3 * u.slot, S_.s[fi], S_.s[ii], u.dt);
        GOPF_CUDA(cudaGetLastError());
        launches_++;
    }
}

double Solver::lp_multiplier(int slot) {
    if (slot < 0 || slot >= GOPF_MAX_SPECIAL || !d_lp_state_) throw Error("lp_multiplier: bad slot");
    synchronize();
    double v = 0.0;
    GOPF_CUDA(cudaMemcpy(&v, d_lp_state_ + 3 * slot, sizeof(double), cudaMemcpyDeviceToHost));
    return v;
}

void Solver::launch_update(const DevKProgram& P) {
    const long long n = (long long)plan_->N;
    // tabulate the implicit factor of equations whose denominator is expensive per k
    ImplicitTab tab{};
    for (int i = 0; i < P.n_fields; ++i) {
        const DevEquation& q = P.eq[i];
        bool expensive = P.filter != nullptr;
        bool k_only = true;
        for (int j = 0; j < q.n_den; ++j) {
            if (q.den[j].kind == TK_SPECTRAL_VISC || q.den[j].kind == TK_PAIR_CORR) expensive = true;
            if (q.den[j].brick >= 0 || q.den[j].kind == TK_VOLUME_LP || q.den[j].kind == TK_CONS_NOISE) k_only = false;
        }
        if (!(expensive && k_only)) continue;
        if (!implicit_tab_[i]) GOPF_CUDA(cudaMalloc(&implicit_tab_[i], sizeof(cplx) * n));
        if (implicit_tab_dirty_) {
            k_implicit_table<<<grid_for(n), 256, 0, stream()>>>(P, i, implicit_tab_[i], plan_->freq_geom(), n);
            GOPF_CUDA(cudaGetLastError());
            launches_++;
        }
        tab.t[i] = implicit_tab_[i];
    }
    implicit_tab_dirty_ = false;
    // algorithmic bytes: every spectrum the program reads + the fields it writes (+ tables)
    int n_read = 0;
    for (int b = 0; b < GOPF_MAX_SPECTRA; ++b)
        if (S_.s[b]) n_read++;
    int n_tab = 0;
    for (int i = 0; i < P.n_fields; ++i) n_tab += tab.t[i] ? 1 : 0;
    const int id = tick("k_update", 16.0 * (double)n * (n_read + P.n_fields + n_tab));
    if (!launch_update_jit(P, tab)) {
        k_update_generic<<<grid_for(n), 256, 0, stream()>>>(P, S_, tab, plan_->freq_geom(), n);
        GOPF_CUDA(cudaGetLastError());
    }
    tock(id);
}

// The k-space kernels compiled for exactly this program (jit.h).  The images are keyed on the
// program bytes (they hold the filter / multiplier addresses too) and on which fields have a
// tabulated implicit factor; any change recompiles.  false: not specialised.
bool Solver::ensure_jit_program(const DevKProgram& P, unsigned tab_mask) {
    if (!jit_on_ || has_knoise_) return false;  // the noise step counter would change the image every step
    std::string key(reinterpret_cast<const char*>(&P), sizeof(P));
    key.push_back((char)tab_mask);
    if (key != jit_kupdate_key_) {
        jit::unload(jit_kupdate_);
        jit::unload(jit_rk4_rhs_);
        jit::unload(jit_rk4_point_);
        jit_kupdate_ = jit_rk4_rhs_ = jit_rk4_point_ = nullptr;
        jit_kupdate_key_ = key;
        std::string log;
        std::vector<char> cubin;
        try {
            if (jit::compile_cubin(jit::kupdate_kernel_source(P, plan_->freq_geom(), (long long)plan_->N, tab_mask), &cubin, &log)) {
                jit_kupdate_ = jit::load(cubin, "gopf_jit_kupdate", &log);
                jit_rk4_rhs_ = jit::load(cubin, "gopf_jit_rk4_rhs", &log);
                jit_rk4_point_ = jit::load(cubin, "gopf_jit_rk4_point", &log);
            }
        } catch (const std::exception& e) {
            log = e.what();
        }
        if (!jit_kupdate_ || !jit_rk4_rhs_ || !jit_rk4_point_) jit_log_ += "k-space kernels: " + log + "\n";
    }
    return jit_kupdate_ && jit_rk4_rhs_ && jit_rk4_point_;
}

bool Solver::launch_update_jit(const DevKProgram& P, const ImplicitTab& tab) {
    unsigned mask = 0;
    for (int i = 0; i < P.n_fields; ++i)
        if (tab.t[i]) mask |= 1u << i;
    if (!ensure_jit_program(P, mask)) return false;
    SpectraPtrs sp = S_;
    ImplicitTab t = tab;
    void* args[] = {&sp, &t};
    std::string log;
    if (!jit::launch(jit_kupdate_, grid_for((long long)plan_->N), 256, args, stream(), &log))
        throw Error("jit launch of k_update: " + log);
    return true;
}

// ---- host synchronisation ------------------------------------------------------------
// max |Im x| over an array (upload: is the field real?)
__global__ void k_max_abs_imag(const cplx* __restrict__ x, long long n, double* out) {
    double m = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmax(m, fabs(x[i].y));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
}

void Solver::resolve_real_check() {
    if (!real_check_pending_) return;
    GOPF_CUDA(cudaStreamSynchronize(stream()));
    const bool now_real = *h_imag_max_ == 0.0;
    if (now_real != field_real_) drop_graph();  // a captured step holds the kernel choice
    field_real_ = now_real;
    real_check_pending_ = false;
}

void Solver::upload() {
    blocked_ = false;  // the spectra are rewritten, row-major
    ensure_buffers();
    rebuild_program();
    cudaStream_t s = stream();
    const int F = (int)m_->fields.size();
    for (int i = 0; i < F; ++i) {
        if (!m_->fields[i].host) throw Error("solver: field '" + m_->fields[i].name + "' has no host array");
        GOPF_CUDA(cudaMemcpyAsync(S_.s[i], m_->fields[i].host, sizeof(cplx) * plan_->N, cudaMemcpyHostToDevice, s));
        if (fused_ && i == 0) {
            // is the field real?  One read of the array on the device (a host scan would cost more than the copy);
            // the answer is collected at the next step (resolve_real_check)
            if (!d_imag_max_) GOPF_CUDA(cudaMalloc(&d_imag_max_, sizeof(double)));
            if (!h_imag_max_) GOPF_CUDA(cudaMallocHost(&h_imag_max_, sizeof(double)));
            GOPF_CUDA(cudaMemsetAsync(d_imag_max_, 0, sizeof(double), s));
            k_max_abs_imag<<<grid_for((long long)plan_->N), 256, 0, s>>>(S_.s[i], (long long)plan_->N, d_imag_max_);
            GOPF_CUDA(cudaGetLastError());
            GOPF_CUDA(cudaMemcpyAsync(h_imag_max_, d_imag_max_, sizeof(double), cudaMemcpyDeviceToHost, s));
            real_check_pending_ = true;
            launches_++;
        }
        plan_->exec_device(S_.s[i], -1, s);  // euler.go:19-21
    }
    on_device_ = true;
    w_valid_ = false;
}

void Solver::download() {
    if (!on_device_) throw Error("solver: nothing on the device to download");
    plan_->use_device();
    leave_blocked();
    cudaStream_t s = stream();
    const int F = (int)m_->fields.size();
    for (int i = 0; i < F; ++i) {
        inverse_to_real(S_.s[i], Rw_[i]);  // euler.go:42-45
        GOPF_CUDA(cudaMemcpyAsync(m_->fields[i].host, Rw_[i], sizeof(cplx) * plan_->N, cudaMemcpyDeviceToHost, s));
    }
    GOPF_CUDA(cudaStreamSynchronize(s));
}

void Solver::download_real(int field, double* host_out, bool big_endian) {
    if (!on_device_) throw Error("solver: nothing on the device to download");
    if (field < 0 || field >= (int)m_->fields.size()) throw Error("download_real: field index out of range");
    if (!host_out) throw Error("download_real: host_out is NULL");
    plan_->use_device();
    leave_blocked();
    ensure_buffers();
    cudaStream_t s = stream();
    const long long n = (long long)plan_->N;
    if (!d_real_out_) GOPF_CUDA(cudaMalloc(&d_real_out_, sizeof(double) * n));
    inverse_to_real(S_.s[field], Rw_[field]);  // euler.go:42-45
    k_real_part<<<grid_for(n), 256, 0, s>>>(Rw_[field], d_real_out_, big_endian ? 1 : 0, n);
    GOPF_CUDA(cudaGetLastError());
    launches_++;
    GOPF_CUDA(cudaMemcpyAsync(host_out, d_real_out_, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    GOPF_CUDA(cudaStreamSynchronize(s));
}

// ---- steppers ------------------------------------------------------------------------
void Solver::euler_step_generic() {
    euler_update_generic();
    volume_lp_hooks();                                       // solver.go:74-82
    elastic_hooks();
}

void Solver::euler_update_generic() {
    bool any_derived = false;
    for (const DerivedSpec& d : m_->derived) any_derived |= d.used;
    const bool elastic = has_elastic();
    if (any_derived || (elastic && elast_valid_)) {
        eval_real_fields();                                  // real-space fields for SyncDerivedFields (euler.go:18)
        for (size_t d = 0; d < m_->derived.size(); ++d)
            if (m_->derived[d].used) forward_derived((int)d);  // euler.go:22-24
    }
    squared_gradient_terms();
    catalog_terms();
    if (elastic) elastic_terms();
    launch_update(prog_);                                    // euler.go:27-39
}

// ---- blocked k-space layout (solver.h) -------------------------------------------------
// GOPF_BLOCKED: 0 never, 1 whenever the shape allows, unset: 3-D grids whose axis-0 row stride reaches 8 MB and
// whose axis-0 / axis-1 lines the copy-engine kernels cover (measured on B200, scripts/tune_blocked.py, 1024^3:
// axis-0 pass 2973 GB/s row-major -> 6060 GB/s blocked s = 7; middle-axis passes converting on the fly 6064 /
// 6217 GB/s against 6206 GB/s row-major to row-major).  GOPF_BLOCK_LOG overrides s.
int Solver::blocked_log() const {
    if (!fused_ || stepper_ != StepperKind::Euler || plan_->rank != 3) return 0;
    if (fused_prog_.fast == 2) return 0;  // the table of the tabulated form is kept row-major
    const int n0 = plan_->n0, n1 = plan_->n1, n2 = plan_->n2;
    if ((n0 & (n0 - 1)) != 0 || n0 < 16) return 0;
    const int mode = env_int("GOPF_BLOCKED", -1);
    if (mode == 0) return 0;
    const bool tma_ok = env_int("GOPF_TMA", 1) != 0 && n0 >= env_int("GOPF_TMA_MIN_N", 1024) &&
                        n1 >= env_int("GOPF_TMA_MIN_N", 1024) && (n0 == 512 || n0 == 1024) && (n1 == 512 || n1 == 1024);
    if (mode < 0 && (!tma_ok || (long long)n1 * n2 * (long long)sizeof(cplx) < (8LL << 20))) return 0;
    int s = env_int("GOPF_BLOCK_LOG", 7);
    while (s > 1 && (n0 >> s) < 2) --s;
    return s < 1 ? 0 : s;
}

PassGeom Solver::blocked_axis0_geom() const {
    const int s = block_log_;
    PassGeom g = plan_->geom(0);
    g.axis = GOPF_AXIS0_BY_PLANE;
    g.A = plan_->n1;
    g.B = plan_->n2;
    g.bcount = g.bw = g.B;
    RowMap r = uniform_rows((long long)plan_->n2 << s, plan_->n2);
    r.split_stride = ((long long)plan_->n1 * plan_->n2) << s;
    r.split_log = s;
    r.split_mask = (1 << s) - 1;
    g.in = g.out = r;
    return g;
}

PassGeom Solver::blocked_axis1_geom(bool in_blocked, bool out_blocked) const {
    const int s = block_log_;
    PassGeom g = plan_->geom(1);
    RowMap r = uniform_rows(plan_->n2, (long long)plan_->n2 << s);  // slab = axis-0 index (split), row = axis-1 index
    r.a_split_stride = ((long long)plan_->n1 * plan_->n2) << s;
    r.a_split_log = s;
    r.a_split_mask = (1 << s) - 1;
    if (in_blocked) g.in = r;
    if (out_blocked) g.out = r;
    return g;
}

void Solver::enter_blocked() {
    if (blocked_) return;
    cudaStream_t s = stream();
    const long long n = (long long)plan_->N;
    if (!W2_) GOPF_CUDA(cudaMalloc(&W2_, sizeof(cplx) * n));
    k_relayout_blocked<<<grid_for(n), 256, 0, s>>>(S_.s[0], W2_, plan_->n1, plan_->n2, block_log_, 1, n);
    GOPF_CUDA(cudaGetLastError());
    launches_++;
    std::swap(S_.s[0], W2_);
    blocked_ = true;
    w_valid_ = false;
}

void Solver::leave_blocked() {
    if (!blocked_) return;
    plan_->use_device();
    cudaStream_t s = stream();
    const long long n = (long long)plan_->N;
    // W2_ holds the first inverse pass of the current spectrum at this point (w_valid_) or nothing of value:
    // either way it is recomputed by the next fused step
    k_relayout_blocked<<<grid_for(n), 256, 0, s>>>(S_.s[0], W2_, plan_->n1, plan_->n2, block_log_, 0, n);
    GOPF_CUDA(cudaGetLastError());
    launches_++;
    std::swap(S_.s[0], W2_);
    blocked_ = false;
    w_valid_ = false;
}

void Solver::euler_step_fused() {
    cudaStream_t s = stream();
    const double n = (double)plan_->N;
    const int slow = plan_->rank == 3 ? 0 : 1;  // slowest active axis
    if (block_log_ < 0) block_log_ = blocked_log();
    const bool blk = block_log_ > 0;
    if (blk) enter_blocked();
    const PassGeom gs = blk ? blocked_axis0_geom() : plan_->geom(slow);
    cplx* Wk = blk ? W2_ : W_;  // the work array on the k-space side of the middle-axis passes
    if (!w_valid_) {
        // first inverse pass of the current spectrum: S -> W
        const int id = tick("pass_inverse", 32.0 * n);
        cudaError_t e = launch_pass(gs, plan_->tx_want, plain_io(S_.s[0], Wk, true, 1.0), plan_->twiddle(slow), s);
        tock(id);
        if (e != cudaSuccess) throw Error(strf("fused: first inverse pass: %s", cudaGetErrorString(e)));
        w_valid_ = true;
    }
    if (plan_->rank == 3) {
        const PassGeom g1 = blk ? blocked_axis1_geom(true, false) : plan_->geom(1);
        const int id = tick("pass_inverse_mid", 32.0 * n);
        cudaError_t e = launch_pass(g1, plan_->tx_want, plain_io(Wk, W_, true, 1.0), plan_->twiddle(1), s);
        tock(id);
        if (e != cudaSuccess) throw Error(strf("fused: middle inverse pass: %s", cudaGetErrorString(e)));
    }
    {
        PassGeom g2 = plan_->geom(2);
        // real field + a fast-form program -- real polynomial multipliers (1) or a real tabulated factor with an
        // optional Hermitian noise spectrum (2): the spectrum stays Hermitian, so the real-space kernel may carry two
        // lines per complex transform (tma_kernels.cuh)
        g2.real_pairs = (field_real_ && (fused_prog_.fast == 1 || fused_prog_.fast == 2)) ? 1 : 0;
        const int id = tick("fused_real", 32.0 * n);
        cudaError_t e = launch_fused_real(g2, 0, W_, nullptr, m_->derived[fused_derived_].dev, 1.0 / n,
                                          (unsigned long long)steps_taken_, plan_->twiddle(2), s);
        tock(id);
        if (e != cudaSuccess) throw Error(strf("fused: real-space kernel: %s", cudaGetErrorString(e)));
    }
    if (plan_->rank == 3) {
        const PassGeom g1 = blk ? blocked_axis1_geom(false, true) : plan_->geom(1);
        const int id = tick("pass_forward_mid", 32.0 * n);
        cudaError_t e = launch_pass(g1, plan_->tx_want, plain_io(W_, Wk, false, 1.0), plan_->twiddle(1), s);
        tock(id);
        if (e != cudaSuccess) throw Error(strf("fused: middle forward pass: %s", cudaGetErrorString(e)));
    }
    {
        const int id = tick("fused_kspace", 64.0 * n);
        cudaError_t e = launch_fused_kspace(gs, plan_->tx_want, Wk, Wk, S_.s[0], fused_prog_, freq_tabs(),
                                            plan_->twiddle(slow), s);
        tock(id);
        if (e != cudaSuccess) throw Error(strf("fused: k-space kernel: %s", cudaGetErrorString(e)));
    }
}

// pf/rk4.go:29-74
void Solver::rk4_step() {
    cudaStream_t s = stream();
    const int F = (int)m_->fields.size();
    const long long n = (long long)plan_->N;
    const FreqGeom fg = plan_->freq_geom();
    SpectraPtrs init{}, fin{}, kf{};
    for (int i = 0; i < F; ++i) {
        init.s[i] = rk_initial_[i];
        fin.s[i] = rk_final_[i];
        kf.s[i] = rk_k_[i];
    }
    bool any_derived = false;
    for (const DerivedSpec& d : m_->derived) any_derived |= d.used;
    auto sync_and_rhs = [&]() {
        if (any_derived) {
            eval_real_fields();
            for (size_t d = 0; d < m_->derived.size(); ++d)
                if (m_->derived[d].used) forward_derived((int)d);
        }
        squared_gradient_terms();
        catalog_terms();
        if (has_elastic()) elastic_terms();
        if (ensure_jit_program(prog_, 0)) {
            SpectraPtrs sp = S_, ko = kf;
            void* args[] = {&sp, &ko};
            std::string log;
            if (!jit::launch(jit_rk4_rhs_, grid_for(n), 256, args, s, &log)) throw Error("jit launch of rk4_rhs: " + log);
        } else {
            k_rk4_rhs<<<grid_for(n), 256, 0, s>>>(prog_, S_, kf, fg, n);
            GOPF_CUDA(cudaGetLastError());
        }
        launches_++;
    };
    auto point = [&](int mode, double fdt) {
        if (ensure_jit_program(prog_, 0)) {
            SpectraPtrs f = S_, a = init, b = fin, c = kf;
            void* args[] = {&mode, &fdt, &f, &a, &b, &c};
            std::string log;
            if (!jit::launch(jit_rk4_point_, grid_for(n), 256, args, s, &log)) throw Error("jit launch of rk4_point: " + log);
        } else {
            k_rk4_point<<<grid_for(n), 256, 0, s>>>(prog_, mode, fdt, S_, init, fin, kf, fg, n);
            GOPF_CUDA(cudaGetLastError());
        }
        launches_++;
    };
    for (int i = 0; i < F; ++i) {  // rk4.go:36-43
        GOPF_CUDA(cudaMemcpyAsync(rk_initial_[i], S_.s[i], sizeof(cplx) * n, cudaMemcpyDeviceToDevice, s));
        GOPF_CUDA(cudaMemcpyAsync(rk_final_[i], S_.s[i], sizeof(cplx) * n, cudaMemcpyDeviceToDevice, s));
    }
    sync_and_rhs();                 // firstCorrection
    point(0, dt_ / 6.0);
    point(1, 0.5 * dt_);            // correction(0.5)
    sync_and_rhs();
    point(0, dt_ / 3.0);
    point(1, 0.5 * dt_);
    sync_and_rhs();
    point(0, dt_ / 3.0);
    point(1, 1.0 * dt_);            // correction(1.0)
    sync_and_rhs();
    point(0, dt_ / 6.0);
    point(2, dt_);                  // rk4.go:57-68
    // solver.go:74-82: OnStepFinished of every term after Stepper.Step, whatever the stepper.  The
    // indicator spectrum of the 4th stage is still in S_, which is what the reference's bricks hold.
    volume_lp_hooks();
    elastic_hooks();
}

void Solver::step(int nsteps) {
    if (!on_device_) throw Error("solver: upload() must run before step()");
    if (nsteps < 0) throw Error("solver: negative step count");
    plan_->use_device();
    ensure_buffers();
    rebuild_program();
    if (!(fused_ && stepper_ == StepperKind::Euler)) leave_blocked();
    {
        // the layout choice follows the environment from one step() call to the next (tuning scripts, tests)
        const int want = blocked_log();
        if (want != block_log_) {
            leave_blocked();
            block_log_ = want;
        }
    }
    resolve_real_check();
    int done = 0;
    if (fused_ && stepper_ == StepperKind::Euler && nsteps > GRAPH_STEPS) {
        // the first step runs eagerly (it also produces W when it is not valid yet and sets every
        // kernel's attributes); whole blocks of GRAPH_STEPS steps are then replayed from the graph
        if (has_knoise_) stamp_noise_step(&fused_prog_, (unsigned long long)steps_taken_);
        euler_step_fused();
        current_step_++;
        steps_taken_++;
        done = 1;
        if (graph_applicable() && (graph_exec_ || build_graph(GRAPH_STEPS))) {
            const long long per_block = (plan_->rank == 3 ? 4 : 2) * (long long)graph_steps_;
            while (nsteps - done >= graph_steps_) {
                GOPF_CUDA(cudaGraphLaunch(graph_exec_, stream()));
                launches_ += per_block;
                current_step_ += graph_steps_;
                steps_taken_ += graph_steps_;
                done += graph_steps_;
            }
        }
    }
    for (int i = done; i < nsteps; ++i) {
        if (has_knoise_) {
            stamp_noise_step(&prog_, (unsigned long long)steps_taken_);
            stamp_noise_step(&fused_prog_, (unsigned long long)steps_taken_);
        }
        if (stepper_ == StepperKind::RK4) {
            rk4_step();  // Step does not advance CurrentStep (rk4.go:130-135)
        } else if (stepper_ == StepperKind::SDD) {
            sdd_step();  // advances its own CurrentStep (sdd.go:299)
        } else if (stepper_ == StepperKind::ImplicitEuler) {
            implicit_euler_step();
            current_step_++;  // implicitEuler.go:206
        } else {
            if (fused_) euler_step_fused();
            else euler_step_generic();
            current_step_++;  // euler.go:46
        }
        steps_taken_++;
    }
}

}  // namespace gopf
