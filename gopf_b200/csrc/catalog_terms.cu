// Catalog terms evaluated outside the k-space program:
//   ChargeTransport  pf/chargeTransport.go:56-146
//   point Sources    pf/sourceTerm.go:25-30, pf/model.go:291-294
// Each fills one work spectrum that the compiled equation reads as a brick, exactly like
// SquaredGradient and HomogeneousModulusLinElast (solver.cu).  The per-cell arithmetic is the
// __host__ __device__ code of catalog_terms.cuh.  Citations: /root/reference.
#include <cmath>
#include <vector>

#include "catalog_terms.cuh"
#include "solver.h"

namespace gopf {

struct FieldPtrs3 {
    const cplx* e[3];
};

static unsigned ct_grid(long long n) {
    long long blocks = (n + 255) / 256;
    const long long cap = 148LL * 16;
    return (unsigned)(blocks < cap ? blocks : cap);
}

// chargeTransport.go:64-73: k-space factor of the comp-th component of the induced field
__global__ void __launch_bounds__(256)
    k_ct_field(const cplx* __restrict__ rho, cplx* __restrict__ out, FreqGeom fg, int comp, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(fg, i, f);
        const double w = ct_field_multiplier(f, fg.rank, comp);
        const cplx r = rho[i];
        out[i] = mk(-r.y * w, r.x * w);  // rho * complex(0, w)
    }
}

// chargeTransport.go:76-86: J_d2 = sum_d sigma[voigt(d, d2)] * (E_d - ExternalField[d]), d ascending
__global__ void __launch_bounds__(256)
    k_ct_current(FieldPtrs3 E, const double* __restrict__ sigma, ChargeParams p, int d2, cplx* __restrict__ out,
                 long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double re = 0.0, im = 0.0;
        for (int d = 0; d < p.dim; ++d) {
            const double s = sigma[(long long)ct_voigt(d, d2, p.dim) * n + i];
            const cplx e = E.e[d][i];
            re += s * (e.x - p.ext[d]);
            im += s * e.y;
        }
        out[i] = mk(re, im);
    }
}

// chargeTransport.go:103-113: field (+)= complex(0, 2 pi f_comp) * FFT(J_comp), Nyquist plane skipped
__global__ void __launch_bounds__(256)
    k_ct_divergence(const cplx* __restrict__ jhat, cplx* __restrict__ out, FreqGeom fg, int comp, int first, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(fg, i, f);
        const double w = ct_divergence_multiplier(f, comp);
        const cplx j = jhat[i];
        cplx acc = first ? mk(0.0, 0.0) : out[i];
        acc = mk(acc.x - j.y * w, acc.y + j.x * w);
        out[i] = acc;
    }
}

// ChargeTransport.Current (:140-145): res[d][i] = -real(current[d*N+i])
__global__ void __launch_bounds__(256) k_ct_minus_real(const cplx* __restrict__ in, double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = -in[i].x;
}

// model.go:291-294 + sourceTerm.go:25-30: sum of the sources of one equation
__global__ void __launch_bounds__(256) k_sources(cplx* __restrict__ out, FreqGeom fg, SourceParams sp, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(fg, i, f);
        double re = 0.0, im = 0.0;
        for (int s = 0; s < sp.n; ++s) {
            double a, b;
            source_value(f, sp.pos[s], sp.rank, sp.amp[s], &a, &b);
            re += a;
            im += b;
        }
        out[i] = mk(re, im);
    }
}

void Solver::free_catalog_buffers() {
    if (ct_tmp_) cudaFree(ct_tmp_);
    ct_tmp_ = nullptr;
    if (obs_partial_) cudaFree(obs_partial_);
    obs_partial_ = nullptr;
    for (int i = 0; i < GOPF_MAX_SPECIAL; ++i) {
        if (ct_sigma_[i]) cudaFree(ct_sigma_[i]);
        ct_sigma_[i] = nullptr;
    }
}

// Real-space current components J_d2 of one ChargeTransport term from the spectrum `rho`
// (ChargeTransport.current, :56-89).  Leaves E_d in sg_tmp_[d]; `emit(d2)` consumes ct_tmp_ = J_d2.
template <class Emit>
void Solver::charge_current_components(const UserTerm& u, const cplx* rho, Emit emit) {
    cudaStream_t s = stream();
    const long long n = (long long)plan_->N;
    const size_t bytes = sizeof(cplx) * plan_->N;
    const int dim = plan_->rank;
    if (u.n_voigt != (dim == 2 ? 3 : 6))
        throw Error(strf("ChargeTransport: %d conductivity components given, a %d-D grid needs %d", u.n_voigt, dim,
                         dim == 2 ? 3 : 6));
    if (u.slot < 0 || u.slot >= GOPF_MAX_SPECIAL) throw Error("ChargeTransport: bad slot");
    if (u.conductivity.size() != (size_t)u.n_voigt * plan_->N)
        throw Error("ChargeTransport: conductivity table does not match the grid");
    if (!ct_sigma_[u.slot]) {
        GOPF_CUDA(cudaMalloc(&ct_sigma_[u.slot], sizeof(double) * u.conductivity.size()));
        GOPF_CUDA(cudaMemcpyAsync(ct_sigma_[u.slot], u.conductivity.data(), sizeof(double) * u.conductivity.size(),
                                  cudaMemcpyHostToDevice, s));
    }
    if (!ct_tmp_) GOPF_CUDA(cudaMalloc(&ct_tmp_, bytes));
    for (int d = 0; d < dim; ++d)
        if (!sg_tmp_[d]) GOPF_CUDA(cudaMalloc(&sg_tmp_[d], bytes));
    const FreqGeom fg = plan_->freq_geom();
    FieldPtrs3 E;
    E.e[0] = E.e[1] = E.e[2] = nullptr;
    for (int d = 0; d < dim; ++d) {
        const int id = tick("charge_field", 32.0 * (double)n);
        k_ct_field<<<ct_grid(n), 256, 0, s>>>(rho, sg_tmp_[d], fg, d, n);
        tock(id);
        GOPF_CUDA(cudaGetLastError());
        inverse_to_real(sg_tmp_[d], sg_tmp_[d]);  // IFFT and /N (:74-75)
        E.e[d] = sg_tmp_[d];
    }
    ChargeParams p;
    p.dim = dim;
    p.n_voigt = u.n_voigt;
    for (int d = 0; d < 3; ++d) p.ext[d] = u.external_field[d];
    for (int d2 = 0; d2 < dim; ++d2) {
        const int id = tick("charge_current", (16.0 * dim + 8.0 * dim + 16.0) * (double)n);
        k_ct_current<<<ct_grid(n), 256, 0, s>>>(E, ct_sigma_[u.slot], p, d2, ct_tmp_, n);
        tock(id);
        GOPF_CUDA(cudaGetLastError());
        emit(d2);
    }
}

// ChargeTransport.Construct (:94-119) for every registered term -> its work spectrum
void Solver::charge_transport_terms() {
    cudaStream_t s = stream();
    const long long n = (long long)plan_->N;
    for (const auto& kv : m_->user_terms) {
        const UserTerm& u = kv.second;
        if (u.kind != UserTermKind::ChargeTransport) continue;
        const int fi = m_->field_index(u.field);
        if (fi < 0) throw Error("ChargeTransport: unknown field " + u.field);
        for (size_t e = 0; e < m_->compiled.size(); ++e)
            for (const DevTerm& t : m_->compiled[e].rhs)
                if (t.brick == u.work_spectrum && (int)e > fi)
                    throw Error("ChargeTransport of field '" + u.field +
                                "' used in a later equation than its own: the reference would read the already "
                                "updated spectrum (euler.go:27-39); this ordering is not supported on the device");
        cplx* out = S_.s[u.work_spectrum];
        const FreqGeom fg = plan_->freq_geom();
        charge_current_components(u, S_.s[fi], [&](int d2) {
            forward_in_place(ct_tmp_);  // :104
            const int id = tick("charge_divergence", (d2 == 0 ? 32.0 : 48.0) * (double)n);
            k_ct_divergence<<<ct_grid(n), 256, 0, s>>>(ct_tmp_, out, fg, d2, d2 == 0 ? 1 : 0, n);
            tock(id);
            GOPF_CUDA(cudaGetLastError());
        });
    }
}

// ChargeTransport.Current(density, N, false) on the device-resident spectrum of the term's field
void Solver::charge_current(const std::string& name, double* host_out) {
    leave_blocked();
    if (!on_device_) throw Error("solver: nothing on the device (upload first)");
    if (!host_out) throw Error("charge_current: host_out is NULL");
    auto it = m_->user_terms.find(name);
    if (it == m_->user_terms.end() || it->second.kind != UserTermKind::ChargeTransport)
        throw Error("charge_current: '" + name + "' is not a registered ChargeTransport term");
    const UserTerm& u = it->second;
    const int fi = m_->field_index(u.field);
    if (fi < 0) throw Error("ChargeTransport: unknown field " + u.field);
    plan_->use_device();
    ensure_buffers();
    cudaStream_t s = stream();
    const long long n = (long long)plan_->N;
    if (!d_real_out_) GOPF_CUDA(cudaMalloc(&d_real_out_, sizeof(double) * n));
    charge_current_components(u, S_.s[fi], [&](int d2) {
        k_ct_minus_real<<<ct_grid(n), 256, 0, s>>>(ct_tmp_, d_real_out_, n);
        GOPF_CUDA(cudaGetLastError());
        launches_++;
        GOPF_CUDA(cudaMemcpyAsync(host_out + (size_t)d2 * n, d_real_out_, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
        GOPF_CUDA(cudaStreamSynchronize(s));  // d_real_out_ is reused by the next component
    });
}

// sum of the point sources of every equation at t = GetTime() -> the equation's source spectrum
void Solver::source_terms() {
    cudaStream_t s = stream();
    const long long n = (long long)plan_->N;
    const double t = get_time();
    for (size_t e = 0; e < m_->sources.size() && e < m_->source_spectrum.size(); ++e) {
        const int si = m_->source_spectrum[e];
        if (si < 0) continue;
        SourceParams sp;
        sp.n = (int)m_->sources[e].size();
        sp.rank = plan_->rank;
        for (int k = 0; k < sp.n; ++k) {
            const SourceSpec& src = m_->sources[e][k];
            if (src.npos < plan_->rank)  // Dot(freq, Pos) would index past Pos (a Go panic)
                throw Error(strf("Source: Pos has %d coordinates, the grid has rank %d", src.npos, plan_->rank));
            for (int c = 0; c < 3; ++c) sp.pos[k][c] = src.pos[c];
            sp.amp[k] = src.fn(t, src.user);  // TimeDepSource on the host (sourceTerm.go:11)
        }
        const int id = tick("sources", 16.0 * (double)n);
        k_sources<<<ct_grid(n), 256, 0, s>>>(S_.s[si], plan_->freq_geom(), sp, n);
        tock(id);
        GOPF_CUDA(cudaGetLastError());
    }
}

// ---- epoch-boundary observers on the device-resident state --------
// block partials: row 0 sum, row 1 minimum of re a, row 2 maximum of re a
//   mode 0: sum of q0 * (q1 n^2 + q2 n^3 + q3 n^4), n = re a     IdealMixtureTerm.GetEnergy
//   mode 1: sum of re(a * b)                                      PairCorrlationTerm.GetEnergy
//   mode 2: extrema only                                          pfutil.MinReal / MaxReal
__global__ void __launch_bounds__(256)
    k_observe(const cplx* __restrict__ a, const cplx* __restrict__ b, int mode, double q0, double q1, double q2, double q3,
              double* __restrict__ partial, long long n) {
    __shared__ double sh[3][256];
    double acc = 0.0, mn = 1.0 / 0.0, mx = -1.0 / 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const cplx x = a[i];
        mn = fmin(mn, x.x);
        mx = fmax(mx, x.x);
        if (mode == 0) {
            const double v = x.x;
            acc += q0 * (q1 * v * v + q2 * v * v * v + q3 * v * v * v * v);  // pfc/ideal.go:26-28
        } else if (mode == 1) {
            const cplx y = b[i];
            acc += x.x * y.x - x.y * y.y;
        }
    }
    sh[0][threadIdx.x] = acc;
    sh[1][threadIdx.x] = mn;
    sh[2][threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
            sh[1][threadIdx.x] = fmin(sh[1][threadIdx.x], sh[1][threadIdx.x + s]);
            sh[2][threadIdx.x] = fmax(sh[2][threadIdx.x], sh[2][threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = sh[0][0];
        partial[gridDim.x + blockIdx.x] = sh[1][0];
        partial[2 * gridDim.x + blockIdx.x] = sh[2][0];
    }
}

// pairCorrelationTerm.go:66-70: field[k] *= Prefactor * C2(2 pi |f|)
__global__ void __launch_bounds__(256)
    k_pair_corr_weight(const cplx* __restrict__ in, cplx* __restrict__ out, PairCorrParams p, FreqGeom fg, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(fg, i, f);
        const double frad = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        const double w = p.prefactor * pair_corr_eval(p, 2.0 * GOPF_PI * frad);
        const cplx x = in[i];
        out[i] = mk(x.x * w, x.y * w);
    }
}

// RealPartAsUint8 (pf/util.go:108-117): uint8(255 * (re - min) / (max - min)), same operation order
__global__ void __launch_bounds__(256)
    k_real_as_uint8(const cplx* __restrict__ in, unsigned char* __restrict__ out, double mn, double mx, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double num = 255.0 * (in[i].x - mn);
        out[i] = (unsigned char)(num / (mx - mn));
    }
}

static const unsigned OBS_BLOCKS = 1024;

// sum / min / max over the array(s); finished on the host in block order
void Solver::observe(const cplx* a, const cplx* b, int mode, const double* q, double* sum, double* mn, double* mx) {
    cudaStream_t s = stream();
    const long long n = (long long)plan_->N;
    unsigned blocks = ct_grid(n);
    if (blocks > OBS_BLOCKS) blocks = OBS_BLOCKS;
    if (!obs_partial_) GOPF_CUDA(cudaMalloc(&obs_partial_, sizeof(double) * 3 * OBS_BLOCKS));
    k_observe<<<blocks, 256, 0, s>>>(a, b, mode, q[0], q[1], q[2], q[3], obs_partial_, n);
    GOPF_CUDA(cudaGetLastError());
    launches_++;
    std::vector<double> h(3 * (size_t)blocks);
    GOPF_CUDA(cudaMemcpyAsync(h.data(), obs_partial_, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, s));
    GOPF_CUDA(cudaStreamSynchronize(s));
    double acc = 0.0, lo = 1.0 / 0.0, hi = -1.0 / 0.0;
    for (unsigned k = 0; k < blocks; ++k) {
        acc += h[k];
        lo = std::fmin(lo, h[blocks + k]);
        hi = std::fmax(hi, h[2 * (size_t)blocks + k]);
    }
    if (sum) *sum = acc;
    if (mn) *mn = lo;
    if (mx) *mx = hi;
}

// IdealMixtureTerm.GetEnergy (pairCorrelationTerm.go:185-193) / PairCorrlationTerm.GetEnergy (:58-84)
// of the registered term `name` on the device-resident state
double Solver::term_energy(const std::string& name) {
    leave_blocked();
    if (!on_device_) throw Error("solver: nothing on the device (upload first)");
    auto it = m_->user_terms.find(name);
    if (it == m_->user_terms.end()) throw Error("GetEnergy: '" + name + "' is not a registered term");
    const UserTerm& u = it->second;
    const int fi = m_->field_index(u.field);
    if (fi < 0) throw Error("GetEnergy: unknown field " + u.field);
    plan_->use_device();
    ensure_buffers();
    inverse_to_real(S_.s[fi], Rw_[fi]);
    double sum = 0.0;
    if (u.kind == UserTermKind::IdealMixture) {
        const double q[4] = {u.prefactor, 0.5, -u.c3 / 6.0, u.c4 / 12.0};  // pfc/ideal.go:36-50
        observe(Rw_[fi], Rw_[fi], 0, q, &sum, nullptr, nullptr);
        return sum;
    }
    if (u.kind == UserTermKind::PairCorrelation || u.kind == UserTermKind::ExplicitPairCorrelation) {
        if (!sg_tmp_[0]) GOPF_CUDA(cudaMalloc(&sg_tmp_[0], sizeof(cplx) * plan_->N));
        PairCorrParams p = u.pc;
        p.prefactor = u.prefactor;
        const long long n = (long long)plan_->N;
        k_pair_corr_weight<<<ct_grid(n), 256, 0, stream()>>>(S_.s[fi], sg_tmp_[0], p, plan_->freq_geom(), n);
        GOPF_CUDA(cudaGetLastError());
        launches_++;
        inverse_to_real(sg_tmp_[0], sg_tmp_[0]);
        const double q[4] = {0.0, 0.0, 0.0, 0.0};
        observe(sg_tmp_[0], Rw_[fi], 1, q, &sum, nullptr, nullptr);
        return -0.5 * sum;
    }
    throw Error("GetEnergy: term '" + name + "' has no energy (IdealMixtureTerm and PairCorrlationTerm do)");
}

// Uint8IO.SaveFields payload of one field (pf/fileIO.go:29-44, pf/util.go:108-117): minimum and
// maximum of the real part, then the real part scaled to 0..255 -- 1 byte per cell over PCIe
void Solver::download_uint8(int field, unsigned char* host_out, double* mn_out, double* mx_out) {
    leave_blocked();
    if (!on_device_) throw Error("solver: nothing on the device to download");
    if (field < 0 || field >= (int)m_->fields.size()) throw Error("download_uint8: field index out of range");
    if (!host_out) throw Error("download_uint8: host_out is NULL");
    plan_->use_device();
    ensure_buffers();
    cudaStream_t s = stream();
    const long long n = (long long)plan_->N;
    if (!d_real_out_) GOPF_CUDA(cudaMalloc(&d_real_out_, sizeof(double) * n));
    inverse_to_real(S_.s[field], Rw_[field]);
    double mn = 0.0, mx = 0.0;
    const double q[4] = {0.0, 0.0, 0.0, 0.0};
    observe(Rw_[field], Rw_[field], 2, q, nullptr, &mn, &mx);
    if (mn_out) *mn_out = mn;
    if (mx_out) *mx_out = mx;
    if (std::fabs(mx - mn) < 1e-10) mx = mn + 1.0;  // util.go:110-112
    unsigned char* bytes = reinterpret_cast<unsigned char*>(d_real_out_);
    k_real_as_uint8<<<ct_grid(n), 256, 0, s>>>(Rw_[field], bytes, mn, mx, n);
    GOPF_CUDA(cudaGetLastError());
    launches_++;
    GOPF_CUDA(cudaMemcpyAsync(host_out, bytes, (size_t)n, cudaMemcpyDeviceToHost, s));
    GOPF_CUDA(cudaStreamSynchronize(s));
}

void Solver::catalog_terms() {
    if (m_->n_work_spectra == 0) return;
    charge_transport_terms();
    source_terms();
}

}  // namespace gopf
