// FftPlan implementation: twiddle / k tables and the plain (unfused) transform.
#include <cmath>
#include <cstring>

#include "fft_plan.h"

namespace gopf {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

// ---- any length (non power of two): O(n^2) DFT per line, out of place -----------
// Completes the FFTWWrapper contract (FFTW accepts any n); not a performance path.
template <bool INV>
__global__ void k_pass_dft_generic(PassGeom g, const cplx* __restrict__ in, cplx* __restrict__ out,
                                   const cplx* __restrict__ tw, double scale) {
    const long long total = g.A * g.N * g.B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i % g.B;
        const long long k = (i / g.B) % g.N;
        const long long a = i / (g.B * g.N);
        const cplx* src = in + (size_t)a * g.N * g.B + b;
        double sx = 0.0, sy = 0.0;
        long long e = 0;  // (j*k) mod N
        for (int j = 0; j < g.N; ++j) {
            cplx w = tw[e];
            if (INV) w.y = -w.y;
            cplx x = src[(size_t)j * g.B];
            sx += x.x * w.x - x.y * w.y;
            sy += x.x * w.y + x.y * w.x;
            e += k;
            if (e >= g.N) e -= g.N;
        }
        out[i] = mk(sx * scale, sy * scale);
    }
}

__global__ void k_fill_freq(double* f, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double v = (double)i / (double)n;  // same IEEE divide as fftWrap.go:62-67
        if (v > 0.5) v -= 1.0;             // fftWrap.go:69-71, Nyquist stays +0.5
        f[i] = v;
    }
}

__global__ void k_freq_nodes(FreqGeom g, bool use_tables, int n1, int n2, const double* f0, const double* f1,
                             const double* f2, const long long* nodes, long long count, double* out) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    long long i = nodes[q];
    double res[3] = {0.0, 0.0, 0.0};
    if (use_tables) {
        // decomposition used by the fused kernels: FFTW coordinates -> per-axis tables
        long long i2 = i % n2, i1 = (i / n2) % n1, i0 = i / ((long long)n1 * n2);
        res[0] = f1[i1];
        res[1] = f2[i2];
        if (g.rank > 2) res[2] = f0[i0];
    } else {
        ref_freq(g, i, res);
    }
    for (int k = 0; k < g.rank; ++k) out[q * g.rank + k] = res[k];
}

FftPlan::FftPlan(int rank_, const int* n, int device_) : rank(rank_), d_scratch_(nullptr), d_buf_(nullptr) {
    if (rank < 1 || rank > 3) throw Error(strf("fft plan: rank must be 1, 2 or 3 (got %d)", rank));
    for (int i = 0; i < 3; ++i) dims[i] = 1;
    for (int i = 0; i < rank; ++i) {
        if (n[i] < 1) throw Error(strf("fft plan: extent %d of axis %d is not positive", n[i], i));
        dims[i] = n[i];
    }
    n0 = n1 = n2 = 1;
    if (rank == 1) { n2 = n[0]; }
    else if (rank == 2) { n1 = n[0]; n2 = n[1]; }
    else { n0 = n[0]; n1 = n[1]; n2 = n[2]; }
    N = (size_t)n0 * n1 * n2;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw Error(strf("fft plan: no CUDA device available (%s); libgopfcuda has no CPU fallback",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)));
    if (device_ < 0) GOPF_CUDA(cudaGetDevice(&device_));
    if (device_ >= ndev) throw Error(strf("fft plan: device %d out of range (%d visible)", device_, ndev));
    device = device_;
    use_device();
    GOPF_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    tx_want = 8;
    for (int ax = 0; ax < 3; ++ax) {
        d_freq_[ax] = nullptr;
        const int len = extent(ax);
        GOPF_CUDA(cudaMalloc(&d_freq_[ax], sizeof(double) * len));
        k_fill_freq<<<(len + 255) / 256, 256, 0, stream>>>(d_freq_[ax], len);
        GOPF_CUDA(cudaGetLastError());
        if (len > 1 && d_tw_.find(len) == d_tw_.end()) {
            std::vector<cplx> tw(len);
            for (int j = 0; j < len; ++j) {
                // exact octant reduction is unnecessary in long double: |err| < 1e-19
                long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)len;
                tw[j] = mk((double)cosl(ang), (double)sinl(ang));
            }
            // exact values on the axes
            tw[0] = mk(1.0, 0.0);
            if (len % 2 == 0) tw[len / 2] = mk(-1.0, 0.0);
            if (len % 4 == 0) { tw[len / 4] = mk(0.0, -1.0); tw[3 * len / 4] = mk(0.0, 1.0); }
            cplx* d = nullptr;
            GOPF_CUDA(cudaMalloc(&d, sizeof(cplx) * len));
            GOPF_CUDA(cudaMemcpyAsync(d, tw.data(), sizeof(cplx) * len, cudaMemcpyHostToDevice, stream));
            GOPF_CUDA(cudaStreamSynchronize(stream));
            d_tw_[len] = d;
        }
    }
    GOPF_CUDA(cudaStreamSynchronize(stream));
}

FftPlan::~FftPlan() {
    cudaSetDevice(device);
    for (auto& kv : d_tw_) cudaFree(kv.second);
    for (int ax = 0; ax < 3; ++ax) cudaFree(d_freq_[ax]);
    if (d_scratch_) cudaFree(d_scratch_);
    if (d_buf_) cudaFree(d_buf_);
    cudaStreamDestroy(stream);
}

void FftPlan::use_device() const { GOPF_CUDA(cudaSetDevice(device)); }

FreqGeom FftPlan::freq_geom() const {
    FreqGeom g;
    g.rank = rank;
    g.d0 = dims[0];
    g.d1 = dims[1];
    g.d2 = dims[2];
    return g;
}

bool FftPlan::freq_axis_consistent() const {
    if (rank == 2) return true;
    if (rank == 3) return dims[0] == dims[1] && dims[1] == dims[2];
    return false;
}

int FftPlan::axis_of_component(int c) const { return c == 0 ? 1 : (c == 1 ? 2 : 0); }

const cplx* FftPlan::twiddle(int axis) const {
    auto it = d_tw_.find(extent(axis));
    return it == d_tw_.end() ? nullptr : it->second;
}

bool FftPlan::axis_fast(int axis) const {
    const PassGeom g = geom(axis);
    if (!fast_length(g.N)) return false;
    if (g.B == 1) return true;
    return (g.B % 2) == 0;
}

cplx* FftPlan::scratch() {
    if (!d_scratch_) {
        use_device();
        GOPF_CUDA(cudaMalloc(&d_scratch_, sizeof(cplx) * N));
    }
    return d_scratch_;
}

void FftPlan::generic_pass(cplx* data, int axis, int sign, cudaStream_t s) {
    const PassGeom g = geom(axis);
    cplx* tmp = scratch();
    const int threads = 256;
    long long blocks = ((long long)N + threads - 1) / threads;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    if (sign < 0)
        k_pass_dft_generic<false><<<(unsigned)blocks, threads, 0, s>>>(g, data, tmp, twiddle(axis), 1.0);
    else
        k_pass_dft_generic<true><<<(unsigned)blocks, threads, 0, s>>>(g, data, tmp, twiddle(axis), 1.0);
    GOPF_CUDA(cudaGetLastError());
    GOPF_CUDA(cudaMemcpyAsync(data, tmp, sizeof(cplx) * N, cudaMemcpyDeviceToDevice, s));
}

void FftPlan::exec_device(cplx* data, int sign, cudaStream_t s) {
    if (sign != -1 && sign != 1) throw Error(strf("fft exec: sign must be -1 or +1 (got %d)", sign));
    use_device();
    if (!s) s = stream;
    for (int axis = 2; axis >= 0; --axis) {
        if (extent(axis) <= 1) continue;
        if (axis_fast(axis)) {
            const PassGeom g = geom(axis);
            cudaError_t e = launch_pass(g, tx_want, plain_io(data, data, sign > 0, 1.0), twiddle(axis), s);
            if (e != cudaSuccess)
                throw Error(strf("fft exec: pass along axis %d (n=%d) failed: %s", axis, g.N, cudaGetErrorString(e)));
        } else {
            generic_pass(data, axis, sign, s);
        }
    }
}

void FftPlan::exec_host(double* host, int sign) {
    if (!host) throw Error("fft exec: host pointer is NULL");
    use_device();
    if (!d_buf_) GOPF_CUDA(cudaMalloc(&d_buf_, sizeof(cplx) * N));
    GOPF_CUDA(cudaMemcpyAsync(d_buf_, host, sizeof(cplx) * N, cudaMemcpyHostToDevice, stream));
    exec_device(d_buf_, sign, stream);
    GOPF_CUDA(cudaMemcpyAsync(host, d_buf_, sizeof(cplx) * N, cudaMemcpyDeviceToHost, stream));
    GOPF_CUDA(cudaStreamSynchronize(stream));
}

void FftPlan::freq_device(const long long* nodes, long long count, double* out) {
    if (rank < 2) throw Error("Freq: the reference indexes res[1] unconditionally (fftWrap.go:61); rank must be 2 or 3");
    use_device();
    long long* d_nodes = nullptr;
    double* d_out = nullptr;
    GOPF_CUDA(cudaMalloc(&d_nodes, sizeof(long long) * count));
    GOPF_CUDA(cudaMalloc(&d_out, sizeof(double) * count * rank));
    GOPF_CUDA(cudaMemcpyAsync(d_nodes, nodes, sizeof(long long) * count, cudaMemcpyHostToDevice, stream));
    k_freq_nodes<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(freq_geom(), freq_axis_consistent(), n1, n2,
                                                                      d_freq_[0], d_freq_[1], d_freq_[2], d_nodes,
                                                                      count, d_out);
    GOPF_CUDA(cudaGetLastError());
    GOPF_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * count * rank, cudaMemcpyDeviceToHost, stream));
    GOPF_CUDA(cudaStreamSynchronize(stream));
    cudaFree(d_nodes);
    cudaFree(d_out);
}

}  // namespace gopf
