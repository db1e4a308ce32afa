// Gradient-based catalog terms at operator level (the remainder):
//   GradientCalculator.Calculate   pf/gradientCalculator.go:19-31
//   DivGrad                        pf/gradientCalculator.go:49-112   div(F grad field)
//   WeightedLaplacian.Construct    pf/gradientCalculator.go:131-172  F LAP field
//   Advection                      pf/advection.go:50-96             -v . grad field
// In the reference as shipped these types cannot be registered with a model (their OnStepFinished lacks the
// `bricks` argument of pf.PureTerm): they are reached by calling PrepareModel / Construct by hand,
// as their tests do.  The entry points here are that surface: one call = what Construct's closure leaves in
// `field` after the derived fields PrepareModel registered have been updated, on device-resident arrays of an
// FFT plan (host-buffer wrappers at the end).  Every transform is the plan's own pass kernels; the multiplier
// i 2 pi f_c sits in the load of the pass along axis c where the shape allows (LK_GRADIENT_LINE), otherwise in a
// pointwise k-space kernel with the literal Freq.
#include "../../include/gopf_cuda.h"
#include "fft_kernels.cuh"
#include "fft_plan.h"
#include "host_util.h"

using namespace gopf;

struct gopf_fft_plan {
    FftPlan* p;
};

namespace {

// data[i] *= i 2 pi f_comp(i)   (zero_nyquist: f = +1/2 -> 0, gradientCalculator.go:24-27)
__global__ void k_mul_gradient(cplx* __restrict__ data, FreqGeom fg, int comp, int zero_nyquist, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(fg, i, f);
        double fd = f[comp];
        if (zero_nyquist && fabs(fd - 0.5) < 1e-10) fd = 0.0;
        const double w = 2.0 * GOPF_PI * fd;
        const cplx u = data[i];
        data[i] = mk(-u.y * w, u.x * w);
    }
}

// data[i] *= -(2 pi |f(i)|)^2   (LaplacianN{Power: 1}.Eval, pf/diffOp.go:25-30)
__global__ void k_mul_laplacian(cplx* __restrict__ data, FreqGeom fg, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(fg, i, f);
        const KPoint kp = make_kpoint(f[0], f[1], f[2]);
        const cplx u = data[i];
        data[i] = mk(u.x * kp.L, u.y * kp.L);
    }
}

// out[i] = sign * sum_d a_d[i] * b_d[i]  (+ out[i] when accumulate)
__global__ void k_sum_products(cplx* out, const cplx* a0, const cplx* b0, const cplx* a1, const cplx* b1,
                               const cplx* a2, const cplx* b2, int dim, double sign, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        cplx acc = a0[i] * b0[i];
        if (dim > 1) acc += a1[i] * b1[i];
        if (dim > 2) acc += a2[i] * b2[i];
        out[i] = mk(acc.x * sign, acc.y * sign);
    }
}

// out[i] (+)= i 2 pi f_comp(i) * in[i]   (DivGrad.Construct, gradientCalculator.go:96-108: no Nyquist zeroing)
__global__ void k_accumulate_gradient(cplx* __restrict__ out, const cplx* __restrict__ in, FreqGeom fg, int comp, int first,
                                      long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(fg, i, f);
        const double w = 2.0 * GOPF_PI * f[comp];
        const cplx u = in[i];
        const cplx t = mk(-u.y * w, u.x * w);
        out[i] = first ? t : mk(out[i].x + t.x, out[i].y + t.y);
    }
}

__global__ void k_scale(cplx* __restrict__ data, double s, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        data[i] = mk(data[i].x * s, data[i].y * s);
}

unsigned grid_of(long long n) {
    long long b = (n + 255) / 256;
    if (b > 148LL * 16) b = 148LL * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

struct DevArrays {  // scratch arrays of one call, freed on scope exit
    std::vector<cplx*> v;
    cplx* get(size_t n) {
        cplx* p = nullptr;
        GOPF_CUDA(cudaMalloc(&p, sizeof(cplx) * n));
        v.push_back(p);
        return p;
    }
    ~DevArrays() {
        for (cplx* p : v) cudaFree(p);
    }
};

void check_comp(const FftPlan& p, int comp) {
    if (p.rank < 2) throw Error("gradient terms: rank must be 2 or 3 (Freq indexes res[1], fftWrap.go:61)");
    if (comp < 0 || comp >= p.rank) throw Error(strf("gradient terms: component %d out of range for rank %d", comp, p.rank));
}

// GradientCalculator.Calculate: out = IFFT(i 2 pi f_comp FFT(in)) / N.  in == out allowed.
void gradient_calculate(FftPlan& p, const cplx* in, cplx* out, int comp, bool keep_nyquist, cudaStream_t s) {
    check_comp(p, comp);
    p.use_device();
    const long long n = (long long)p.N;
    if (in != out) GOPF_CUDA(cudaMemcpyAsync(out, in, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, s));
    p.exec_device(out, -1, s);
    bool all_fast = p.freq_axis_consistent();
    for (int ax = 0; ax < 3; ++ax)
        if (p.extent(ax) > 1 && !p.axis_fast(ax)) all_fast = false;
    const double inv_n = 1.0 / (double)n;
    if (all_fast && !keep_nyquist) {
        // multiplier in the load of the pass along the component's axis, 1/N on the last store
        int last = -1;
        for (int ax = 0; ax < 3; ++ax)
            if (p.extent(ax) > 1) last = ax;
        for (int ax = 0; ax < 3; ++ax) {
            if (p.extent(ax) <= 1) continue;
            PassIO io = plain_io(out, out, true, ax == last ? inv_n : 1.0);
            if (ax == p.axis_of_component(comp)) {
                io.load_kind = LK_GRADIENT_LINE;
                io.rtab = p.freq_axis(ax);
            }
            cudaError_t e = launch_pass(p.geom(ax), p.tx_want, io, p.twiddle(ax), s);
            if (e != cudaSuccess) throw Error(strf("gradient pass along axis %d: %s", ax, cudaGetErrorString(e)));
        }
        return;
    }
    k_mul_gradient<<<grid_of(n), 256, 0, s>>>(out, p.freq_geom(), comp, keep_nyquist ? 0 : 1, n);
    GOPF_CUDA(cudaGetLastError());
    p.exec_device(out, 1, s);
    k_scale<<<grid_of(n), 256, 0, s>>>(out, inv_n, n);
    GOPF_CUDA(cudaGetLastError());
}

// Advection: -sum_d v_d * GRAD_d(field)  (advection.go:59-82, 87-94); transformed: its forward transform, which is
// what the closure reads inside a step (derived fields are transformed before the terms are evaluated)
void advection_construct(FftPlan& p, const cplx* field, const cplx* const* vel, cplx* out, bool transformed, cudaStream_t s) {
    p.use_device();
    const long long n = (long long)p.N;
    DevArrays tmp;
    cplx* g[3] = {nullptr, nullptr, nullptr};
    for (int d = 0; d < p.rank; ++d) {
        if (!vel[d]) throw Error("advection: one velocity field per dimension (advection.go:53-55)");
        g[d] = tmp.get((size_t)n);
        gradient_calculate(p, field, g[d], d, false, s);
    }
    k_sum_products<<<grid_of(n), 256, 0, s>>>(out, vel[0], g[0], vel[1], g[1], p.rank > 2 ? vel[2] : vel[1],
                                              p.rank > 2 ? g[2] : g[1], p.rank, -1.0, n);
    GOPF_CUDA(cudaGetLastError());
    if (transformed) p.exec_device(out, -1, s);
    GOPF_CUDA(cudaStreamSynchronize(s));  // scratch is freed on return
}

// DivGrad.Construct: sum_d i 2 pi f_d FFT(F * GRAD_d(field))  (gradientCalculator.go:72-108)
void div_grad_construct(FftPlan& p, const cplx* field, const cplx* F, cplx* out, cudaStream_t s) {
    p.use_device();
    const long long n = (long long)p.N;
    DevArrays tmp;
    cplx* g = tmp.get((size_t)n);
    for (int d = 0; d < p.rank; ++d) {
        gradient_calculate(p, field, g, d, false, s);
        k_sum_products<<<grid_of(n), 256, 0, s>>>(g, F, g, F, g, F, g, 1, 1.0, n);
        GOPF_CUDA(cudaGetLastError());
        p.exec_device(g, -1, s);
        k_accumulate_gradient<<<grid_of(n), 256, 0, s>>>(out, g, p.freq_geom(), d, d == 0 ? 1 : 0, n);
        GOPF_CUDA(cudaGetLastError());
    }
    GOPF_CUDA(cudaStreamSynchronize(s));
}

// WeightedLaplacian.Construct: FFT( IFFT(L field^)/N * IFFT(prefactor^)/N )  (gradientCalculator.go:131-172)
void weighted_laplacian_construct(FftPlan& p, const cplx* field_hat, const cplx* prefactor_hat, cplx* out, cudaStream_t s) {
    p.use_device();
    const long long n = (long long)p.N;
    const double inv_n = 1.0 / (double)n;
    DevArrays tmp;
    cplx* lap = tmp.get((size_t)n);
    cplx* work = tmp.get((size_t)n);
    GOPF_CUDA(cudaMemcpyAsync(lap, field_hat, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, s));
    k_mul_laplacian<<<grid_of(n), 256, 0, s>>>(lap, p.freq_geom(), n);
    GOPF_CUDA(cudaGetLastError());
    p.exec_device(lap, 1, s);
    k_scale<<<grid_of(n), 256, 0, s>>>(lap, inv_n, n);
    GOPF_CUDA(cudaMemcpyAsync(work, prefactor_hat, sizeof(cplx) * n, cudaMemcpyDeviceToDevice, s));
    p.exec_device(work, 1, s);
    k_scale<<<grid_of(n), 256, 0, s>>>(work, inv_n, n);
    k_sum_products<<<grid_of(n), 256, 0, s>>>(out, lap, work, lap, work, lap, work, 1, 1.0, n);
    GOPF_CUDA(cudaGetLastError());
    p.exec_device(out, -1, s);
    GOPF_CUDA(cudaStreamSynchronize(s));
}

// host-buffer plumbing: arrays of N complex128 up, `fn` on the plan's stream, result down
struct HostCall {
    FftPlan& p;
    DevArrays dev;
    explicit HostCall(FftPlan& plan) : p(plan) { p.use_device(); }
    cplx* up(const double* host) {
        if (!host) throw Error("gradient terms: host array is NULL");
        cplx* d = dev.get(p.N);
        GOPF_CUDA(cudaMemcpyAsync(d, host, sizeof(cplx) * p.N, cudaMemcpyHostToDevice, p.stream));
        return d;
    }
    cplx* fresh() { return dev.get(p.N); }
    void down(double* host, const cplx* d) {
        if (!host) throw Error("gradient terms: host output array is NULL");
        GOPF_CUDA(cudaMemcpyAsync(host, d, sizeof(cplx) * p.N, cudaMemcpyDeviceToHost, p.stream));
        GOPF_CUDA(cudaStreamSynchronize(p.stream));
    }
};

cudaStream_t stream_of(gopf_fft_plan* plan, void* stream) {
    return stream ? reinterpret_cast<cudaStream_t>(stream) : plan->p->stream;
}

}  // namespace

extern "C" {

int gopf_gradient_calculate_device(gopf_fft_plan* plan, const void* in, void* out, int comp, int keep_nyquist, void* stream) {
    GOPF_API_BEGIN
    if (!plan || !in || !out) throw Error("gopf_gradient_calculate_device: NULL argument");
    gradient_calculate(*plan->p, reinterpret_cast<const cplx*>(in), reinterpret_cast<cplx*>(out), comp, keep_nyquist != 0,
                       stream_of(plan, stream));
    GOPF_API_END
}

int gopf_advection_construct_device(gopf_fft_plan* plan, const void* field, const void* const* velocity, int n_velocity,
                                    void* out, int transformed, void* stream) {
    GOPF_API_BEGIN
    if (!plan || !field || !velocity || !out) throw Error("gopf_advection_construct_device: NULL argument");
    if (n_velocity != plan->p->rank) throw Error("Advection: Inconsistent number of velocity fields");  // advection.go:53-55
    const cplx* v[3] = {nullptr, nullptr, nullptr};
    for (int d = 0; d < n_velocity; ++d) v[d] = reinterpret_cast<const cplx*>(velocity[d]);
    advection_construct(*plan->p, reinterpret_cast<const cplx*>(field), v, reinterpret_cast<cplx*>(out), transformed != 0,
                        stream_of(plan, stream));
    GOPF_API_END
}

int gopf_div_grad_construct_device(gopf_fft_plan* plan, const void* field, const void* func_values, void* out, void* stream) {
    GOPF_API_BEGIN
    if (!plan || !field || !func_values || !out) throw Error("gopf_div_grad_construct_device: NULL argument");
    div_grad_construct(*plan->p, reinterpret_cast<const cplx*>(field), reinterpret_cast<const cplx*>(func_values),
                       reinterpret_cast<cplx*>(out), stream_of(plan, stream));
    GOPF_API_END
}

int gopf_weighted_laplacian_construct_device(gopf_fft_plan* plan, const void* field_hat, const void* prefactor_hat, void* out,
                                             void* stream) {
    GOPF_API_BEGIN
    if (!plan || !field_hat || !prefactor_hat || !out) throw Error("gopf_weighted_laplacian_construct_device: NULL argument");
    weighted_laplacian_construct(*plan->p, reinterpret_cast<const cplx*>(field_hat), reinterpret_cast<const cplx*>(prefactor_hat),
                                 reinterpret_cast<cplx*>(out), stream_of(plan, stream));
    GOPF_API_END
}

// ---- host-buffer forms (Field.Data / DerivedField.Data are host slices in the reference) ----------------------
int gopf_gradient_calculate(gopf_fft_plan* plan, const double* in, double* out, int comp, int keep_nyquist) {
    GOPF_API_BEGIN
    if (!plan) throw Error("gopf_gradient_calculate: plan is NULL");
    HostCall h(*plan->p);
    cplx* d = h.up(in);
    gradient_calculate(h.p, d, d, comp, keep_nyquist != 0, h.p.stream);
    h.down(out, d);
    GOPF_API_END
}

int gopf_advection_construct(gopf_fft_plan* plan, const double* field, const double* const* velocity, int n_velocity,
                             double* out, int transformed) {
    GOPF_API_BEGIN
    if (!plan || !velocity) throw Error("gopf_advection_construct: NULL argument");
    if (n_velocity != plan->p->rank) throw Error("Advection: Inconsistent number of velocity fields");
    HostCall h(*plan->p);
    const cplx* v[3] = {nullptr, nullptr, nullptr};
    for (int d = 0; d < n_velocity; ++d) v[d] = h.up(velocity[d]);
    cplx* f = h.up(field);
    cplx* o = h.fresh();
    advection_construct(h.p, f, v, o, transformed != 0, h.p.stream);
    h.down(out, o);
    GOPF_API_END
}

int gopf_div_grad_construct(gopf_fft_plan* plan, const double* field, const double* func_values, double* out) {
    GOPF_API_BEGIN
    if (!plan) throw Error("gopf_div_grad_construct: plan is NULL");
    HostCall h(*plan->p);
    cplx* f = h.up(field);
    cplx* F = h.up(func_values);
    cplx* o = h.fresh();
    div_grad_construct(h.p, f, F, o, h.p.stream);
    h.down(out, o);
    GOPF_API_END
}

int gopf_weighted_laplacian_construct(gopf_fft_plan* plan, const double* field_hat, const double* prefactor_hat, double* out) {
    GOPF_API_BEGIN
    if (!plan) throw Error("gopf_weighted_laplacian_construct: plan is NULL");
    HostCall h(*plan->p);
    cplx* f = h.up(field_hat);
    cplx* w = h.up(prefactor_hat);
    cplx* o = h.fresh();
    weighted_laplacian_construct(h.p, f, w, o, h.p.stream);
    h.down(out, o);
    GOPF_API_END
}

}  // extern "C"
