// Shrinking-Dimer-Dynamics stepper on the device (pf/sdd.go:86-470).
//
// One step (sdd.go:158-300) on the device-resident spectra.  The reference shifts the real-space
// fields along the dimer and transforms them again; the transform is linear, so the shifted
// spectra are S +- s * v^ (v^ = FFT of the orientation block) and only the derived fields need
// the real-space image:
//   1. v^_i = FFT(v_i);  l = DimerLength(t)
//   2. S -= l/2 v^ : right-hand side at the start image;  S += l v^ : at the end image;  S -= l/2 v^
//   3. per field: w = alpha rhsStart[0:N] + (1 - alpha) rhsEnd[0:N]   (as written, :204-206: the
//      FIRST field's block for every field), Householder reflection w -= 2 v^_i <w, v^_i>/N,
//      d = S_i + dt w, then (D + u v^T)^-1 d by Sherman-Morrison with D = 1 - dt den_i (:327-344,
//      :445-470)
//   4. torque = rhsStart - rhsEnd - den l v^, projected (sigma = 1), inverse transformed; the
//      orientation follows it and is renormalised (:229-289)
// Global sums (inner products over k, norms) are reduced per block on the device and finished on
// the host in a fixed order, so a step is deterministic; the stepper is latency-, not
// bandwidth-bound by construction (a dozen host-visible scalars per step) and is not a bench path.
// Citations: /root/reference.
#include <cmath>
#include <vector>

#include "solver.h"

namespace gopf {

namespace {

constexpr int SDD_BLOCKS = 1024;
constexpr int SDD_VALUES = 5;  // up to 4 sums + 1 maximum per block

unsigned sdd_grid(long long n) {
    long long blocks = (n + 255) / 256;
    return (unsigned)(blocks < SDD_BLOCKS ? blocks : SDD_BLOCKS);
}

// sums[0..NS) and one maximum of a 256-thread block -> partial[v * gridDim.x + blockIdx.x]
template <int NS>
__device__ __forceinline__ void sdd_block_reduce(const double (&acc)[NS], double mx, double* __restrict__ partial) {
    __shared__ double sh[NS + 1][256];
#pragma unroll
    for (int v = 0; v < NS; ++v) sh[v][threadIdx.x] = acc[v];
    sh[NS][threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
#pragma unroll
            for (int v = 0; v < NS; ++v) sh[v][threadIdx.x] += sh[v][threadIdx.x + s];
            sh[NS][threadIdx.x] = fmax(sh[NS][threadIdx.x], sh[NS][threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int v = 0; v < NS; ++v) partial[(size_t)v * gridDim.x + blockIdx.x] = sh[v][0];
        partial[(size_t)NS * gridDim.x + blockIdx.x] = sh[NS][0];
    }
}

#define SDD_LOOP(i, n) \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__device__ __forceinline__ cplx mul_conj(cplx a, cplx b) {  // a * conj(b)
    return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

__device__ __forceinline__ cplx denominator(const DevKProgram& P, int i, const SpectraPtrs& sp, const FreqGeom& fg,
                                            long long idx) {  // Model.GetDenum(i) at one k (model.go:306-314)
    double f[3] = {0.0, 0.0, 0.0};
    ref_freq(fg, idx, f);
    const KPoint kp = make_kpoint(f[0], f[1], f[2]);
    const DevEquation& q = P.eq[i];
    cplx den = mk(0.0, 0.0);
    for (int j = 0; j < q.n_den; ++j) den += eval_term(P, q.den[j], kp, [&](int b) -> cplx { return sp.s[b][idx]; });
    return den;
}

__global__ void __launch_bounds__(256) k_sdd_to_cplx(const double* __restrict__ x, cplx* __restrict__ out, long long n) {
    SDD_LOOP(i, n) out[i] = mk(x[i], 0.0);
}

// Projection onto real fields: out[k] = (S[k] + conj(S[-k])) / 2, -k = ConjugateNode(k)
// (pfutil/fftWrap.go:78-95) in the reference's node numbering.
__global__ void __launch_bounds__(256)
    k_sdd_hermitian(const cplx* __restrict__ in, cplx* __restrict__ out, FreqGeom fg, long long n) {
    SDD_LOOP(i, n) {
        const long long c = i % fg.d1, q = i / fg.d1;
        const long long r = q % fg.d0, d = q / fg.d0;
        const long long cc = (fg.d1 - c) % fg.d1, rc = (fg.d0 - r) % fg.d0;
        const long long dc = fg.rank > 2 ? (fg.d2 - d) % fg.d2 : 0;
        const long long j = (dc * fg.d0 + rc) * fg.d1 + cc;
        const cplx a = in[i], b = in[j];
        out[i] = mk(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
    }
}

// FieldNorm (sdd.go:378-387) by Parseval: sum |c|^2 = sum |c^|^2 / N
__global__ void __launch_bounds__(256) k_sdd_norm(SpectraPtrs S, int F, double* __restrict__ partial, long long n) {
    double acc[1] = {0.0};
    SDD_LOOP(i, n)
        for (int f = 0; f < F; ++f) {
            const cplx c = S.s[f][i];
            acc[0] += c.x * c.x + c.y * c.y;
        }
    sdd_block_reduce<1>(acc, 0.0, partial);
}

// ShiftFieldsAlongDimer (:348-356) in k-space: S_f += scale * v^_f
__global__ void __launch_bounds__(256) k_sdd_shift(SpectraPtrs S, SpectraPtrs V, int F, double scale, long long n) {
    SDD_LOOP(i, n)
        for (int f = 0; f < F; ++f) {
            const cplx v = V.s[f][i];
            cplx c = S.s[f][i];
            c.x += scale * v.x;
            c.y += scale * v.y;
            S.s[f][i] = c;
        }
}

// extractRHS (:302-307): GetRHS of every equation, no in-place field update
__global__ void __launch_bounds__(256)
    k_sdd_rhs(const __grid_constant__ DevKProgram P, SpectraPtrs sp, SpectraPtrs out, FreqGeom fg, long long n) {
    SDD_LOOP(idx, n) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(fg, idx, f);
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

// :202-207: work = alpha rhsStart + (1 - alpha) rhsEnd, and <work, v^> for the reflection (:313-317)
__global__ void __launch_bounds__(256)
    k_sdd_work(const cplx* __restrict__ rs0, const cplx* __restrict__ re0, const cplx* __restrict__ vhat, double alpha,
               cplx* __restrict__ work, double* __restrict__ partial, long long n) {
    double acc[2] = {0.0, 0.0};
    SDD_LOOP(i, n) {
        const cplx a = rs0[i], b = re0[i];
        const cplx w = mk(alpha * a.x + (1.0 - alpha) * b.x, alpha * a.y + (1.0 - alpha) * b.y);
        work[i] = w;
        const cplx p = mul_conj(w, vhat[i]);
        acc[0] += p.x;
        acc[1] += p.y;
    }
    sdd_block_reduce<2>(acc, 0.0, partial);
}

// :209-216: reflected force, explicit Euler update d = S_i + dt w, and the two sums of
// diagonalShermannMorrison.dot (:457-461).  `dot` is <work, v^>/N.
__global__ void __launch_bounds__(256)
    k_sdd_predict(const __grid_constant__ DevKProgram P, int fi, SpectraPtrs sp, const cplx* __restrict__ work,
                  const cplx* __restrict__ vhat, cplx dot, double dt, double sigma, cplx* __restrict__ dout, FreqGeom fg,
                  double* __restrict__ partial, long long n) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const double inv_n = 1.0 / (double)n;
    SDD_LOOP(i, n) {
        const cplx v = vhat[i];
        const cplx w = work[i] - (v * dot) * sigma;  // householder (:319-323)
        const cplx s = sp.s[fi][i];
        const cplx d = mk(s.x + dt * w.x, s.y + dt * w.y);
        dout[i] = d;
        const cplx den = denominator(P, fi, sp, fg, i);
        const cplx inv_d = cdiv(mk(1.0, 0.0), mk(1.0 - dt * den.x, -dt * den.y));  // :338
        const cplx vs = mul_conj(mk(sigma * dt * den.x, sigma * dt * den.y), v) * inv_n;  // :339
        const cplx a = (v * inv_d) * vs;   // u * invD * v
        const cplx b = (vs * inv_d) * d;   // v * invD * vec
        acc[0] += a.x;
        acc[1] += a.y;
        acc[2] += b.x;
        acc[3] += b.y;
    }
    sdd_block_reduce<4>(acc, 0.0, partial);
}

// :463-469 and the monitor (:219-225): S_i = invD d - invD u vdot / denum
__global__ void __launch_bounds__(256)
    k_sdd_correct(const __grid_constant__ DevKProgram P, int fi, SpectraPtrs sp, const cplx* __restrict__ d,
                  const cplx* __restrict__ vhat, cplx ratio /* vdot / denum */, double dt, FreqGeom fg,
                  double* __restrict__ partial, long long n) {
    double acc[1] = {0.0};
    double mx = 0.0;
    const double inv_n = 1.0 / (double)n;
    SDD_LOOP(i, n) {
        const cplx den = denominator(P, fi, sp, fg, i);
        const cplx inv_d = cdiv(mk(1.0, 0.0), mk(1.0 - dt * den.x, -dt * den.y));
        const cplx nv = inv_d * d[i] - (inv_d * vhat[i]) * ratio;
        const cplx old = sp.s[fi][i];
        const double diff = hypot((nv.x - old.x) / dt, (nv.y - old.y) / dt);
        acc[0] += diff * diff * inv_n;
        mx = fmax(mx, diff);
        sp.s[fi][i] = nv;
    }
    sdd_block_reduce<1>(acc, mx, partial);
}

// :232-248: torque_i = rhsStart_i - rhsEnd_i - den_i v^_i l, and <torque, v^> for the projection
__global__ void __launch_bounds__(256)
    k_sdd_torque(const __grid_constant__ DevKProgram P, int fi, SpectraPtrs sp, cplx* __restrict__ rs,
                 const cplx* __restrict__ re, const cplx* __restrict__ vhat, double l, FreqGeom fg,
                 double* __restrict__ partial, long long n) {
    double acc[2] = {0.0, 0.0};
    SDD_LOOP(i, n) {
        const cplx den = denominator(P, fi, sp, fg, i);
        const cplx v = vhat[i];
        const cplx t = rs[i] - re[i] - (den * v) * l;
        rs[i] = t;
        const cplx p = mul_conj(t, v);
        acc[0] += p.x;
        acc[1] += p.y;
    }
    sdd_block_reduce<2>(acc, 0.0, partial);
}

// householder (:319-323) with the inner product already divided by N
__global__ void __launch_bounds__(256)
    k_sdd_reflect(cplx* __restrict__ data, const cplx* __restrict__ vhat, cplx dot, double sigma, long long n) {
    SDD_LOOP(i, n) data[i] = data[i] - (vhat[i] * dot) * sigma;
}

// :277-283: orientation -= coef * Re torque; max |Re torque|; sum orientation^2 (:286)
__global__ void __launch_bounds__(256)
    k_sdd_orient(double* __restrict__ orient, const cplx* __restrict__ torque, double coef, double* __restrict__ partial,
                 long long n) {
    double acc[1] = {0.0};
    double mx = 0.0;
    SDD_LOOP(i, n) {
        const double t = torque[i].x;
        const double o = orient[i] - coef * t;
        orient[i] = o;
        acc[0] += o * o;
        mx = fmax(mx, fabs(t));
    }
    sdd_block_reduce<1>(acc, mx, partial);
}

__global__ void __launch_bounds__(256) k_sdd_scale(double* __restrict__ a, double s, long long n) {
    SDD_LOOP(i, n) a[i] *= s;
}

}  // namespace

void Solver::sdd_free_buffers() {
    auto fr = [](void* p) {
        if (p) cudaFree(p);
    };
    fr(sdd_orient_);
    fr(sdd_work_);
    fr(sdd_d_);
    fr(sdd_partial_);
    sdd_orient_ = nullptr;
    sdd_work_ = sdd_d_ = nullptr;
    sdd_partial_ = nullptr;
    for (int i = 0; i < GOPF_MAX_FIELDS; ++i) {
        fr(sdd_vhat_[i]);
        fr(sdd_rs_[i]);
        fr(sdd_re_[i]);
        sdd_vhat_[i] = sdd_rs_[i] = sdd_re_[i] = nullptr;
    }
}

void Solver::sdd_ensure_buffers() {
    plan_->use_device();
    const int F = (int)m_->fields.size();
    const size_t bytes = sizeof(cplx) * plan_->N;
    if (!sdd_orient_) {
        GOPF_CUDA(cudaMalloc(&sdd_orient_, sizeof(double) * plan_->N * F));
        GOPF_CUDA(cudaMemset(sdd_orient_, 0, sizeof(double) * plan_->N * F));  // NewSDD: zeros (:124)
    }
    for (int i = 0; i < F; ++i) {
        if (!sdd_vhat_[i]) GOPF_CUDA(cudaMalloc(&sdd_vhat_[i], bytes));
        if (!sdd_rs_[i]) GOPF_CUDA(cudaMalloc(&sdd_rs_[i], bytes));
        if (!sdd_re_[i]) GOPF_CUDA(cudaMalloc(&sdd_re_[i], bytes));
    }
    if (!sdd_work_) GOPF_CUDA(cudaMalloc(&sdd_work_, bytes));
    if (!sdd_d_) GOPF_CUDA(cudaMalloc(&sdd_d_, bytes));
    if (!sdd_partial_) GOPF_CUDA(cudaMalloc(&sdd_partial_, sizeof(double) * SDD_VALUES * SDD_BLOCKS));
}

// finishes a block reduction on the host: sums[v] over blocks for v < n_sums, maximum of row n_sums
void Solver::sdd_collect(int n_sums, unsigned blocks, double* sums, double* mx) {
    std::vector<double> h((size_t)(n_sums + 1) * blocks);
    GOPF_CUDA(cudaMemcpyAsync(h.data(), sdd_partial_, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, stream()));
    GOPF_CUDA(cudaStreamSynchronize(stream()));
    for (int v = 0; v < n_sums; ++v) {
        double s = 0.0;
        for (unsigned b = 0; b < blocks; ++b) s += h[(size_t)v * blocks + b];
        sums[v] = s;
    }
    double m = 0.0;
    for (unsigned b = 0; b < blocks; ++b) m = std::fmax(m, h[(size_t)n_sums * blocks + b]);
    if (mx) *mx = m;
}

// sdd.go:364-370
double Solver::sdd_dimer_length(double t) const {
    const double l = sdd_.init_dimer_length * std::exp(-t / sdd_.tau_dimer_length);
    return l < sdd_.min_dimer_length ? sdd_.min_dimer_length : l;
}

// SetInitialOrientation (sdd.go:413-427); Init(init, final) is the same call with
// real(final - init) (:390-408)
void Solver::sdd_set_orientation(const double* orient, long long len) {
    if (stepper_ != StepperKind::SDD) throw Error("SDD: select the stepper first (set_stepper \"sdd\")");
    const long long M = (long long)plan_->N * (long long)m_->fields.size();
    if (!orient || len != M) throw Error("Inconsistent length of the passed orientaiton vector");
    sdd_ensure_buffers();
    double len2 = 0.0;
    for (long long i = 0; i < M; ++i) len2 += orient[i] * orient[i];
    sdd_.init_dimer_length = std::sqrt(len2);
    std::vector<double> v(orient, orient + M);
    for (double& x : v) x /= sdd_.init_dimer_length;
    GOPF_CUDA(cudaMemcpy(sdd_orient_, v.data(), sizeof(double) * M, cudaMemcpyHostToDevice));
    sdd_.initialized = true;
}

void Solver::sdd_get_orientation(double* host_out) {
    if (stepper_ != StepperKind::SDD) throw Error("SDD: not the active stepper");
    if (!host_out) throw Error("sdd_get_orientation: host_out is NULL");
    sdd_ensure_buffers();
    synchronize();
    const long long M = (long long)plan_->N * (long long)m_->fields.size();
    GOPF_CUDA(cudaMemcpy(host_out, sdd_orient_, sizeof(double) * M, cudaMemcpyDeviceToHost));
}

void Solver::sdd_set(const std::string& key, double value) {
    if (stepper_ != StepperKind::SDD) throw Error("SDD: select the stepper first (set_stepper \"sdd\")");
    if (key == "Alpha") sdd_.alpha = value;
    else if (key == "Dt") sdd_.dt = value;
    else if (key == "TimeConstants.Orientation") sdd_.tau_orientation = value;
    else if (key == "TimeConstants.DimerLength") sdd_.tau_dimer_length = value;
    else if (key == "MinDimerLength") sdd_.min_dimer_length = value;
    else if (key == "InitDimerLength") sdd_.init_dimer_length = value;
    else if (key == "CurrentStep") sdd_.current_step = (long long)value;
    else throw Error("SDD: unknown setting '" + key + "'");
}

double Solver::sdd_get(const std::string& key) {
    if (stepper_ != StepperKind::SDD) throw Error("SDD: not the active stepper");
    synchronize();
    if (key == "Alpha") return sdd_.alpha;
    if (key == "Dt") return sdd_.dt;
    if (key == "TimeConstants.Orientation") return sdd_.tau_orientation;
    if (key == "TimeConstants.DimerLength") return sdd_.tau_dimer_length;
    if (key == "MinDimerLength") return sdd_.min_dimer_length;
    if (key == "InitDimerLength") return sdd_.init_dimer_length;
    if (key == "CurrentStep") return (double)sdd_.current_step;
    if (key == "DimerLength") return sdd_dimer_length(get_time());
    if (key == "Monitor.MaxForce") return sdd_.max_force;
    if (key == "Monitor.ForcePowerSpectrum") return sdd_.force_power_spectrum;
    if (key == "Monitor.MaxTorque") return sdd_.max_torque;
    if (key == "Monitor.FieldNorm") return sdd_.field_norm;
    if (key == "Monitor.FieldNormChange") return sdd_.field_norm_change;
    throw Error("SDD: unknown quantity '" + key + "'");
}

// SDD.Step (sdd.go:158-300)
void Solver::sdd_step() {
    if (!sdd_.initialized) throw Error("SDD: The method have to be initialized first. See SDD.Init\n");
    if (sdd_.dt < 1e-16)
        throw Error("Timestep not set in SDD. Make sure that the Dt attribute has explicitly been set.");
    if (d_filter_) throw Error("SDD: Does not support modal filters");
    if (fused_) throw Error("SDD runs on the general path");
    sdd_ensure_buffers();
    cudaStream_t s = stream();
    const int F = (int)m_->fields.size();
    const long long n = (long long)plan_->N;
    const FreqGeom fg = plan_->freq_geom();
    const unsigned grid = sdd_grid(n);
    const double dt = sdd_.dt;
    DevKProgram P = prog_;
    P.dt = dt;
    auto launched = [&]() {
        GOPF_CUDA(cudaGetLastError());
        launches_++;
    };
    SpectraPtrs V{}, RS{}, RE{};
    for (int i = 0; i < F; ++i) {
        V.s[i] = sdd_vhat_[i];
        RS.s[i] = sdd_rs_[i];
        RE.s[i] = sdd_re_[i];
    }
    double sums[4], mx = 0.0;

    // Registered functions are evaluated on the real parts of the fields (include/gopf_cuda.h), so
    // the non-linearity does not damp the imaginary rounding residue of a real field as the
    // reference's complex-valued closures do; under the reflected dynamics that residue grows along
    // the dimer.  The fields of this stepper are real (ShiftFieldsAlongDimer adds reals, :348-356):
    // project the spectra onto real fields once per step.
    if (!plan_->freq_axis_consistent())
        throw Error("SDD on the device needs a 2-D or a cubic 3-D grid (FFTWWrapper.Freq, SURVEY.md 7)");
    for (int i = 0; i < F; ++i) {
        k_sdd_hermitian<<<grid, 256, 0, s>>>(S_.s[i], sdd_d_, fg, n);
        launched();
        GOPF_CUDA(cudaMemcpyAsync(S_.s[i], sdd_d_, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, s));
    }

    // :164-166 FieldNorm of the real-space fields
    k_sdd_norm<<<grid, 256, 0, s>>>(S_, F, sdd_partial_, n);
    launched();
    sdd_collect(1, grid, sums, nullptr);
    const double fnorm = sums[0] / (double)n;
    sdd_.field_norm_change = sdd_.field_norm - fnorm;
    sdd_.field_norm = fnorm;

    // :168-175 Fourier transform of the orientation blocks
    for (int i = 0; i < F; ++i) {
        k_sdd_to_cplx<<<grid, 256, 0, s>>>(sdd_orient_ + (size_t)i * n, sdd_vhat_[i], n);
        launched();
        plan_->exec_device(sdd_vhat_[i], -1, s);
    }
    const double l = sdd_dimer_length(get_time());

    bool any_derived = false;
    for (const DerivedSpec& d : m_->derived) any_derived |= d.used;
    auto rhs_at_image = [&](SpectraPtrs out) {  // sdd.fft(m) + extractRHS
        if (any_derived || (has_elastic() && elast_valid_)) {
            eval_real_fields();
            for (size_t d = 0; d < m_->derived.size(); ++d)
                if (m_->derived[d].used) forward_derived((int)d);
        }
        squared_gradient_terms();
        catalog_terms();
        if (has_elastic()) elastic_terms();
        k_sdd_rhs<<<grid, 256, 0, s>>>(P, S_, out, fg, n);
        launched();
    };
    auto shift = [&](double scale) {
        k_sdd_shift<<<grid, 256, 0, s>>>(S_, V, F, scale, n);
        launched();
    };
    shift(-0.5 * l);  // :182-184
    rhs_at_image(RS);
    shift(l);         // :187-190
    rhs_at_image(RE);
    shift(-0.5 * l);  // :193-195 (the derived fields at the centre are not read again)

    // :197-226
    double power = 0.0;
    sdd_.max_force = 0.0;
    for (int i = 0; i < F; ++i) {
        k_sdd_work<<<grid, 256, 0, s>>>(sdd_rs_[0], sdd_re_[0], sdd_vhat_[i], sdd_.alpha, sdd_work_, sdd_partial_, n);
        launched();
        sdd_collect(2, grid, sums, nullptr);
        const cplx dot = mk(sums[0] / (double)n, sums[1] / (double)n);
        k_sdd_predict<<<grid, 256, 0, s>>>(P, i, S_, sdd_work_, sdd_vhat_[i], dot, dt, 2.0, sdd_d_, fg, sdd_partial_, n);
        launched();
        sdd_collect(4, grid, sums, nullptr);
        // vDotInvDDotVec / denum (:457-468), complex division as Go does it (Smith)
        const double dr = 1.0 + sums[0], di = sums[1];
        const double nr = sums[2], ni = sums[3];
        cplx ratio;
        if (std::fabs(dr) >= std::fabs(di)) {
            const double r = di / dr, d = dr + r * di;
            ratio = mk((nr + ni * r) / d, (ni - nr * r) / d);
        } else {
            const double r = dr / di, d = di + r * dr;
            ratio = mk((nr * r + ni) / d, (ni * r - nr) / d);
        }
        k_sdd_correct<<<grid, 256, 0, s>>>(P, i, S_, sdd_d_, sdd_vhat_[i], ratio, dt, fg, sdd_partial_, n);
        launched();
        sdd_collect(1, grid, sums, &mx);
        power += sums[0];
        sdd_.max_force = std::fmax(sdd_.max_force, mx);
    }
    sdd_.force_power_spectrum = std::sqrt(power / (double)((long long)F * n));  // :228-230

    // :232-260 torque and its projection
    double tr = 0.0, ti = 0.0;
    for (int i = 0; i < F; ++i) {
        k_sdd_torque<<<grid, 256, 0, s>>>(P, i, S_, sdd_rs_[i], sdd_re_[i], sdd_vhat_[i], l, fg, sdd_partial_, n);
        launched();
        sdd_collect(2, grid, sums, nullptr);
        tr += sums[0];
        ti += sums[1];
    }
    const cplx tdot = mk(tr / (double)n, ti / (double)n);
    for (int i = 0; i < F; ++i) {
        k_sdd_reflect<<<grid, 256, 0, s>>>(sdd_rs_[i], sdd_vhat_[i], tdot, 1.0, n);
        launched();
        inverse_to_real(sdd_rs_[i], sdd_rs_[i]);  // :263-266
    }
    // :277-289 orientation update and normalisation
    const double coef = dt / (sdd_dimer_length(get_time()) * sdd_.tau_orientation);
    double len2 = 0.0;
    sdd_.max_torque = 0.0;
    for (int i = 0; i < F; ++i) {
        k_sdd_orient<<<grid, 256, 0, s>>>(sdd_orient_ + (size_t)i * n, sdd_rs_[i], coef, sdd_partial_, n);
        launched();
        sdd_collect(1, grid, sums, &mx);
        len2 += sums[0];
        sdd_.max_torque = std::fmax(sdd_.max_torque, mx);
    }
    k_sdd_scale<<<sdd_grid(n * F), 256, 0, s>>>(sdd_orient_, 1.0 / std::sqrt(len2), n * F);
    launched();
    sdd_.current_step++;  // :299
    volume_lp_hooks();    // Solver.Propagate: OnStepFinished after Stepper.Step (solver.go:74-82)
    elastic_hooks();
}

}  // namespace gopf
