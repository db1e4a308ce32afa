// ImplicitEuler on the device (pf/implicitEuler.go:20-229).
//
// One step (implicitEuler.go:165-207): snapshot the spectra and the RHS at the start of the
// step, one semi-implicit Euler step as the initial guess (:181-187), then solve F(x) = 0 for
// the real-space fields x, where (updateEquation, :68-95)
//   F(x) = Re IFFT( x^ - [ c0^ e^{den dt} + I(den, rhs(x), rhs0) ] ) / N,
//   I    = a (f - 1)/den + b (f - den dt - 1)/den^2,  a = rhs0, b = (rhs - rhs0)/dt, f = e^{den dt}
//          (0.5 dt (rhs + rhs0 f) where |den| < 1e-5; nonlinearIntegral, :151-162).
// The residual is the hot-path round FFT -> RHS -> IFFT and runs on the same kernels as the
// Euler step; the vectors of the Newton-Krylov iteration stay in HBM, only scalars (dot
// products, the small Hessenberg system) visit the host.
//
// PARITY UNPINNED for the nonlinear solve: the reference calls gononlin v0.2.2 NewtonKrylov +
// gonum/exp linsolve.GMRES, neither under /root/reference.  The algorithm is the one stated in
// oracle/pf.py (NewtonKrylov), mirrored choice for choice; two converged runs agree to the solver
// tolerance (1e-7 on max|F|), not to 1e-10.
#include <cmath>
#include <vector>

#include "solver.h"

namespace gopf {

namespace {

unsigned ie_grid(long long n) {
    long long blocks = (n + 255) / 256;
    const long long cap = 148LL * 16;
    return (unsigned)(blocks < cap ? blocks : cap);
}

constexpr int IE_PARTIALS = 1024;

__global__ void k_real_to_cplx(const double* __restrict__ x, cplx* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = mk(x[i], 0.0);
}

__global__ void k_cplx_to_real(const cplx* __restrict__ in, double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = in[i].x;
}

// out = a + alpha * b  (out may alias a)
__global__ void k_lincomb(const double* a, double alpha, const double* __restrict__ b, double* out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fma(alpha, b[i], a[i]);
}

__global__ void k_scale_real(double* a, double s, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] *= s;
}

// mode 0: sum a*b ; mode 1: max |a|.  One partial per block, finished on the host.
__global__ void __launch_bounds__(256) k_reduce(const double* __restrict__ a, const double* __restrict__ b, int mode,
                                                double* __restrict__ partial, long long n) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mode == 0) acc = fma(a[i], b[i], acc);
        else acc = fmax(acc, fabs(a[i]));
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = mode == 0 ? sh[threadIdx.x] + sh[threadIdx.x + s] : fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__device__ __forceinline__ cplx cexp_(cplx z) {
    double s, c;
    sincos(z.y, &s, &c);
    const double e = exp(z.x);
    return mk(e * c, e * s);
}

// RHS of every equation at every k from the current spectra (GetRHS without the in-place field
// update of Euler.Step: implicitEuler.go:175-177, 78-79)
__global__ void __launch_bounds__(256)
    k_ie_rhs(const __grid_constant__ DevKProgram P, SpectraPtrs sp, SpectraPtrs out, FreqGeom fg, long long n) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq_fast(fg, idx, n <= 0x7fffffffLL, f);  // 32-bit index arithmetic when the grid allows (same IEEE divides)
        const KPoint kp = make_kpoint(f[0], f[1], f[2]);
        auto get = [&](int b) -> cplx { return sp.s[b][idx]; };
        for (int i = 0; i < P.n_fields; ++i) {
            const DevEquation& q = P.eq[i];
            cplx rhs = mk(0.0, 0.0);
            for (int j = 0; j < q.n_rhs; ++j) rhs += eval_term(P, q.rhs[j], kp, get);
            out.s[i][idx] = rhs;
        }
    }
}

// res_i = x^_i - (orig_i e^{den dt} + I(den, rhs, rhs_prev_i))      (implicitEuler.go:81-90)
__global__ void __launch_bounds__(256)
    k_ie_residual(const __grid_constant__ DevKProgram P, SpectraPtrs sp, SpectraPtrs orig, SpectraPtrs rhs_prev, SpectraPtrs res,
                  FreqGeom fg, long long n) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq_fast(fg, idx, n <= 0x7fffffffLL, f);  // 32-bit index arithmetic when the grid allows (same IEEE divides)
        const KPoint kp = make_kpoint(f[0], f[1], f[2]);
        auto get = [&](int b) -> cplx { return sp.s[b][idx]; };
        for (int i = 0; i < P.n_fields; ++i) {
            const DevEquation& q = P.eq[i];
            cplx rhs = mk(0.0, 0.0), den = mk(0.0, 0.0);
            for (int j = 0; j < q.n_rhs; ++j) rhs += eval_term(P, q.rhs[j], kp, get);
            for (int j = 0; j < q.n_den; ++j) den += eval_term(P, q.den[j], kp, get);
            const cplx rp = rhs_prev.s[i][idx];
            const cplx fac = cexp_(mk(den.x * P.dt, den.y * P.dt));
            cplx integral;
            if (hypot(den.x, den.y) < 1e-5) {
                const cplx t = rhs + rp * fac;
                integral = mk(0.5 * P.dt * t.x, 0.5 * P.dt * t.y);
            } else {
                const cplx a = rp;
                const cplx b = mk((rhs.x - rp.x) / P.dt, (rhs.y - rp.y) / P.dt);
                const cplx t1 = cdiv(a * mk(fac.x - 1.0, fac.y), den);
                const cplx t2 = cdiv(b * mk(fac.x - den.x * P.dt - 1.0, fac.y - den.y * P.dt), den * den);
                integral = t1 + t2;
            }
            const cplx update = orig.s[i][idx] * fac + integral;
            res.s[i][idx] = sp.s[i][idx] - update;
        }
    }
}

const double* fd_weights(int stencil, int* n) {
    static const double w2[] = {0.5}, w4[] = {2.0 / 3.0, -1.0 / 12.0}, w6[] = {0.75, -0.15, 1.0 / 60.0};
    switch (stencil) {
        case 2: *n = 1; return w2;
        case 4: *n = 2; return w4;
        case 6: *n = 3; return w6;
        default: throw Error("ImplicitEuler: Stencil must be 2, 4 or 6");
    }
}

}  // namespace

// vector slots
enum { IE_X = 0, IE_FX, IE_B, IE_S, IE_W, IE_TP, IE_TM, IE_R, IE_V0 };

void Solver::ie_ensure_buffers() {
    const int F = (int)m_->fields.size();
    const size_t bytes = sizeof(cplx) * plan_->N;
    for (int i = 0; i < F; ++i) {
        if (!ie_orig_[i]) GOPF_CUDA(cudaMalloc(&ie_orig_[i], bytes));
        if (!ie_rhs_prev_[i]) GOPF_CUDA(cudaMalloc(&ie_rhs_prev_[i], bytes));
        if (!ie_res_[i]) GOPF_CUDA(cudaMalloc(&ie_res_[i], bytes));
    }
    const size_t want = IE_V0 + (size_t)nk_.restart + 1;
    if (ie_vec_restart_ != nk_.restart || ie_vec_.size() != want) {
        for (double* v : ie_vec_)
            if (v) cudaFree(v);
        ie_vec_.assign(want, nullptr);
        for (double*& v : ie_vec_) GOPF_CUDA(cudaMalloc(&v, sizeof(double) * plan_->N * F));
        ie_vec_restart_ = nk_.restart;
    }
    if (!ie_partial_) GOPF_CUDA(cudaMalloc(&ie_partial_, sizeof(double) * IE_PARTIALS));
}

double Solver::ie_dot(const double* a, const double* b) {
    const long long M = (long long)plan_->N * (long long)m_->fields.size();
    unsigned blocks = ie_grid(M);
    if (blocks > IE_PARTIALS) blocks = IE_PARTIALS;
    k_reduce<<<blocks, 256, 0, stream()>>>(a, b, 0, ie_partial_, M);
    GOPF_CUDA(cudaGetLastError());
    launches_++;
    std::vector<double> h(blocks);
    GOPF_CUDA(cudaMemcpyAsync(h.data(), ie_partial_, sizeof(double) * blocks, cudaMemcpyDeviceToHost, stream()));
    GOPF_CUDA(cudaStreamSynchronize(stream()));
    double s = 0.0;
    for (double v : h) s += v;
    return s;
}

double Solver::ie_max_abs(const double* a) {
    const long long M = (long long)plan_->N * (long long)m_->fields.size();
    unsigned blocks = ie_grid(M);
    if (blocks > IE_PARTIALS) blocks = IE_PARTIALS;
    k_reduce<<<blocks, 256, 0, stream()>>>(a, a, 1, ie_partial_, M);
    GOPF_CUDA(cudaGetLastError());
    launches_++;
    std::vector<double> h(blocks);
    GOPF_CUDA(cudaMemcpyAsync(h.data(), ie_partial_, sizeof(double) * blocks, cudaMemcpyDeviceToHost, stream()));
    GOPF_CUDA(cudaStreamSynchronize(stream()));
    double s = 0.0;
    for (double v : h) s = std::fmax(s, v);
    return s;
}

// F(x): updateEquation (implicitEuler.go:68-95).  Leaves S_ = FFT(x).
void Solver::ie_residual(const double* x, double* out) {
    cudaStream_t s = stream();
    const int F = (int)m_->fields.size();
    const long long n = (long long)plan_->N;
    for (int i = 0; i < F; ++i) {  // vec2fields + fft of the fields
        k_real_to_cplx<<<ie_grid(n), 256, 0, s>>>(x + (size_t)i * n, Rw_[i], n);
        GOPF_CUDA(cudaGetLastError());
        GOPF_CUDA(cudaMemcpyAsync(S_.s[i], Rw_[i], sizeof(cplx) * n, cudaMemcpyDeviceToDevice, s));
        plan_->exec_device(S_.s[i], -1, s);
        launches_ += 2;
    }
    for (size_t d = 0; d < m_->derived.size(); ++d)  // SyncDerivedFields + fft of the derived fields
        if (m_->derived[d].used) forward_derived((int)d);
    squared_gradient_terms();
    catalog_terms();
    if (has_elastic()) elastic_terms();
    SpectraPtrs orig{}, rp{}, res{};
    for (int i = 0; i < F; ++i) {
        orig.s[i] = ie_orig_[i];
        rp.s[i] = ie_rhs_prev_[i];
        res.s[i] = ie_res_[i];
    }
    {
        // algorithmic bytes: every spectrum read once, origFields and rhsPrev read, the residual written
        int n_read = 0;
        for (int b = 0; b < GOPF_MAX_SPECTRA; ++b)
            if (S_.s[b]) n_read++;
        const int id = tick("ie_residual_kernel", 16.0 * (double)n * (n_read + 3 * F));
        k_ie_residual<<<ie_grid(n), 256, 0, s>>>(prog_, S_, orig, rp, res, plan_->freq_geom(), n);
        tock(id);
    }
    GOPF_CUDA(cudaGetLastError());
    for (int i = 0; i < F; ++i) {
        inverse_to_real(ie_res_[i], ie_res_[i]);  // IFFT and /N in place
        k_cplx_to_real<<<ie_grid(n), 256, 0, s>>>(ie_res_[i], out + (size_t)i * n, n);
        GOPF_CUDA(cudaGetLastError());
        launches_++;
    }
    ie_residual_evals_++;
}

// J(x) v by central differences of the residual (oracle/pf.py NewtonKrylov.jac_vec)
void Solver::ie_jac_vec(const double* x, const double* v, double* out) {
    cudaStream_t s = stream();
    const long long M = (long long)plan_->N * (long long)m_->fields.size();
    const double nv = std::sqrt(ie_dot(v, v));
    GOPF_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * M, s));
    if (nv == 0.0) return;
    const double eps = nk_.step_size * std::sqrt((double)M) / nv;
    int nw = 0;
    const double* w = fd_weights(nk_.stencil, &nw);
    double* tp = ie_vec_[IE_TP];
    double* tm = ie_vec_[IE_TM];
    double* xp = ie_vec_[IE_R];  // perturbed point (free while a J v product is formed)
    for (int k = 1; k <= nw; ++k) {
        k_lincomb<<<ie_grid(M), 256, 0, s>>>(x, k * eps, v, xp, M);
        ie_residual(xp, tp);
        k_lincomb<<<ie_grid(M), 256, 0, s>>>(x, -(k * eps), v, xp, M);
        ie_residual(xp, tm);
        k_lincomb<<<ie_grid(M), 256, 0, s>>>(out, w[k - 1], tp, out, M);
        k_lincomb<<<ie_grid(M), 256, 0, s>>>(out, -w[k - 1], tm, out, M);
        GOPF_CUDA(cudaGetLastError());
        launches_ += 4;
    }
    k_scale_real<<<ie_grid(M), 256, 0, s>>>(out, 1.0 / eps, M);
    GOPF_CUDA(cudaGetLastError());
    launches_++;
}

// Restarted GMRES for J(x) s = b (oracle/pf.py NewtonKrylov.gmres): modified Gram-Schmidt
// Arnoldi on device vectors, Givens rotations of the small Hessenberg matrix on the host.
void Solver::ie_gmres(const double* x, const double* b, double* sol) {
    cudaStream_t st = stream();
    const long long M = (long long)plan_->N * (long long)m_->fields.size();
    const int m = nk_.restart;
    GOPF_CUDA(cudaMemsetAsync(sol, 0, sizeof(double) * M, st));
    const double bnorm = std::sqrt(ie_dot(b, b));
    if (bnorm == 0.0) return;
    const double target = nk_.inner_tol * bnorm;
    double* w = ie_vec_[IE_W];
    bool sol_nonzero = false;
    std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), y(m);
    auto h = [&](int i, int j) -> double& { return H[(size_t)i * m + j]; };
    for (int cycle = 0; cycle < nk_.max_restarts; ++cycle) {
        double* v0 = ie_vec_[IE_V0];
        if (sol_nonzero) {  // r = b - J s
            ie_jac_vec(x, sol, w);
            k_lincomb<<<ie_grid(M), 256, 0, st>>>(b, -1.0, w, v0, M);
        } else {
            GOPF_CUDA(cudaMemcpyAsync(v0, b, sizeof(double) * M, cudaMemcpyDeviceToDevice, st));
        }
        const double beta = std::sqrt(ie_dot(v0, v0));
        if (beta <= target) break;
        k_scale_real<<<ie_grid(M), 256, 0, st>>>(v0, 1.0 / beta, M);
        std::fill(H.begin(), H.end(), 0.0);
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        int k_used = 0;
        for (int j = 0; j < m; ++j) {
            ie_jac_vec(x, ie_vec_[IE_V0 + j], w);
            for (int i = 0; i <= j; ++i) {
                h(i, j) = ie_dot(w, ie_vec_[IE_V0 + i]);
                k_lincomb<<<ie_grid(M), 256, 0, st>>>(w, -h(i, j), ie_vec_[IE_V0 + i], w, M);
            }
            h(j + 1, j) = std::sqrt(ie_dot(w, w));
            if (h(j + 1, j) > 0.0) {
                GOPF_CUDA(cudaMemcpyAsync(ie_vec_[IE_V0 + j + 1], w, sizeof(double) * M, cudaMemcpyDeviceToDevice, st));
                k_scale_real<<<ie_grid(M), 256, 0, st>>>(ie_vec_[IE_V0 + j + 1], 1.0 / h(j + 1, j), M);
            }
            for (int i = 0; i < j; ++i) {
                const double t = cs[i] * h(i, j) + sn[i] * h(i + 1, j);
                h(i + 1, j) = -sn[i] * h(i, j) + cs[i] * h(i + 1, j);
                h(i, j) = t;
            }
            const double d = std::hypot(h(j, j), h(j + 1, j));
            if (d > 0.0) { cs[j] = h(j, j) / d; sn[j] = h(j + 1, j) / d; }
            else { cs[j] = 1.0; sn[j] = 0.0; }
            h(j, j) = cs[j] * h(j, j) + sn[j] * h(j + 1, j);
            h(j + 1, j) = 0.0;
            g[j + 1] = -sn[j] * g[j];
            g[j] = cs[j] * g[j];
            k_used = j + 1;
            if (std::fabs(g[j + 1]) <= target || h(j, j) == 0.0) break;
        }
        for (int i = k_used - 1; i >= 0; --i) {
            double acc = g[i];
            for (int l = i + 1; l < k_used; ++l) acc -= h(i, l) * y[l];
            y[i] = h(i, i) != 0.0 ? acc / h(i, i) : 0.0;
        }
        for (int i = 0; i < k_used; ++i) k_lincomb<<<ie_grid(M), 256, 0, st>>>(sol, y[i], ie_vec_[IE_V0 + i], sol, M);
        GOPF_CUDA(cudaGetLastError());
        sol_nonzero = true;
        if (std::fabs(g[k_used]) <= target) break;
    }
}

// ImplicitEuler.Step (implicitEuler.go:165-207)
void Solver::implicit_euler_step() {
    cudaStream_t s = stream();
    const int F = (int)m_->fields.size();
    const long long n = (long long)plan_->N;
    const long long M = n * F;
    if (fused_) throw Error("ImplicitEuler runs on the general path");
    ie_ensure_buffers();
    // ie.fft(m): S_ already holds FFT(fields); derived fields from the real-space fields
    eval_real_fields();
    for (size_t d = 0; d < m_->derived.size(); ++d)
        if (m_->derived[d].used) forward_derived((int)d);
    squared_gradient_terms();
    catalog_terms();
    if (has_elastic()) elastic_terms();
    SpectraPtrs rp{};
    for (int i = 0; i < F; ++i) {
        GOPF_CUDA(cudaMemcpyAsync(ie_orig_[i], S_.s[i], sizeof(cplx) * n, cudaMemcpyDeviceToDevice, s));  // origFields
        rp.s[i] = ie_rhs_prev_[i];
    }
    k_ie_rhs<<<ie_grid(n), 256, 0, s>>>(prog_, S_, rp, plan_->freq_geom(), n);  // rhsPrev
    GOPF_CUDA(cudaGetLastError());
    launches_++;
    // explicitEuler.Step(m) as the initial guess; the OnStepFinished hooks belong to Propagate
    euler_update_generic();
    double* x = ie_vec_[IE_X];
    double* fx = ie_vec_[IE_FX];
    eval_real_fields();
    for (int i = 0; i < F; ++i) {  // fields2vec
        k_cplx_to_real<<<ie_grid(n), 256, 0, s>>>(Rw_[i], x + (size_t)i * n, n);
        GOPF_CUDA(cudaGetLastError());
        launches_++;
    }
    // Newton-Krylov (oracle/pf.py NewtonKrylov.Solve)
    ie_converged_ = false;
    for (int it = 0; it < nk_.maxiter; ++it) {
        ie_residual(x, fx);
        if (ie_max_abs(fx) < nk_.tol) {
            ie_converged_ = true;
            break;
        }
        double* b = ie_vec_[IE_B];
        double* step = ie_vec_[IE_S];
        GOPF_CUDA(cudaMemcpyAsync(b, fx, sizeof(double) * M, cudaMemcpyDeviceToDevice, s));
        k_scale_real<<<ie_grid(M), 256, 0, s>>>(b, -1.0, M);
        ie_gmres(x, b, step);
        k_lincomb<<<ie_grid(M), 256, 0, s>>>(x, 1.0, step, x, M);
        GOPF_CUDA(cudaGetLastError());
        launches_ += 2;
    }
    if (!ie_converged_) {  // the update of the last iteration may have converged: look once more before reporting
        ie_residual(x, fx);
        ie_converged_ = ie_max_abs(fx) < nk_.tol;
    }
    // vec2fields(res.X): the persistent spectrum of the new fields
    for (int i = 0; i < F; ++i) {
        k_real_to_cplx<<<ie_grid(n), 256, 0, s>>>(x + (size_t)i * n, S_.s[i], n);
        GOPF_CUDA(cudaGetLastError());
        plan_->exec_device(S_.s[i], -1, s);
        launches_++;
    }
    volume_lp_hooks();  // solver.go:74-82
    elastic_hooks();
}

}  // namespace gopf
