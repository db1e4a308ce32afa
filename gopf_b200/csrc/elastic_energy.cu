// elasticity.HomogeneousModulusEnergy on the device (elasticity/linearElasticity.go:101-165;
// SURVEY.md 8f rank 4): elastic energy per unit precipitate volume of a misfitting inclusion in a
// homogeneous matrix.  A standalone post-processing call, not a step kernel:
//   H^ = FFT(indicator);  u^ from G u^ = F^ per k (Displacements, :16-63);
//   eps^_ij = i pi (f'_j u^_i + f'_i u^_j) (Strain, :67-83) = s_ij(k) H^ with the REAL factor
//   s_ij = 1/2 (f'_j h_i + f'_i h_j),  h = Gamma^-1 b,  b_c = sum_j sigma*_cj f_j  (all three c: the
//   function loops comp < 3 also in 2-D, :114-119, unlike the HomogeneousModulusLinElast term);
//   six inverse transforms, eps_ij(x) = Re IFFT / N - eps*_ij where Re indicator > 0.5;
//   energy = sum_x 1/2 C:eps:eps / sum_x Re indicator.
// Citations: /root/reference.
#include <cmath>
#include <vector>

#include "../../include/gopf_cuda.h"
#include "elastic.cuh"
#include "fft_kernels.cuh"
#include "fft_plan.h"
#include "host_util.h"

namespace gopf {

// s_ij(k) for the pair index p = (00, 01, 02, 11, 12, 22)
__host__ __device__ inline double elastic_strain_factor(const ElastParams& E, double f0, double f1, double f2, int pi, int pj) {
    if (fabs(f0) < 1e-10 && fabs(f1) < 1e-10 && fabs(f2) < 1e-10) return 0.0;  // :43-51
    const double q[6] = {f0 * f0, f1 * f1, f2 * f2, f0 * f1, f0 * f2, f1 * f2};
    double g[9];
    for (int e = 0; e < 9; ++e) {
        double s = 0.0;
        for (int p = 0; p < 6; ++p) s = fma(E.K[e][p], q[p], s);
        g[e] = s;
    }
    const double f[3] = {f0, f1, f2};
    double b[3];
    for (int c = 0; c < 3; ++c) b[c] = E.sigma[c * 3 + 0] * f0 + E.sigma[c * 3 + 1] * f1 + E.sigma[c * 3 + 2] * f2;
    const double c00 = g[4] * g[8] - g[5] * g[7], c01 = g[5] * g[6] - g[3] * g[8], c02 = g[3] * g[7] - g[4] * g[6];
    const double c10 = g[2] * g[7] - g[1] * g[8], c11 = g[0] * g[8] - g[2] * g[6], c12 = g[1] * g[6] - g[0] * g[7];
    const double c20 = g[1] * g[5] - g[2] * g[4], c21 = g[2] * g[3] - g[0] * g[5], c22 = g[0] * g[4] - g[1] * g[3];
    const double det = g[0] * c00 + g[1] * c01 + g[2] * c02;
    const double inv = 1.0 / det;
    double h[3];
    h[0] = (c00 * b[0] + c10 * b[1] + c20 * b[2]) * inv;
    h[1] = (c01 * b[0] + c11 * b[1] + c21 * b[2]) * inv;
    h[2] = (c02 * b[0] + c12 * b[1] + c22 * b[2]) * inv;
    const double fi = (fabs(fabs(f[pi]) - 0.5) < 1e-10) ? 0.0 : f[pi];
    const double fj = (fabs(fabs(f[pj]) - 0.5) < 1e-10) ? 0.0 : f[pj];
    return 0.5 * (fj * h[pi] + fi * h[pj]);
}

struct Stiffness {
    double c[81];
};

namespace {

unsigned ee_grid(long long n) {
    long long blocks = (n + 255) / 256;
    return (unsigned)(blocks < 1024 ? blocks : 1024);
}

__global__ void __launch_bounds__(256)
    k_strain_hat(const cplx* __restrict__ hhat, cplx* __restrict__ out, const __grid_constant__ ElastParams E, FreqGeom fg,
                 int pi, int pj, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};  // padded to three components (:124-132)
        ref_freq(fg, i, f);
        const double s = elastic_strain_factor(E, f[0], f[1], f[2], pi, pj);
        const cplx h = hhat[i];
        out[i] = mk(h.x * s, h.y * s);
    }
}

// :147-158: re = real(strain)/N, minus the misfit where real(indicator) > 0.5
__global__ void __launch_bounds__(256)
    k_strain_store(const cplx* __restrict__ eps, const cplx* __restrict__ indicator, double misfit, double inv_n,
                   double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double re = eps[i].x * inv_n;
        if (indicator[i].x > 0.5) re -= misfit;
        out[i] = re;
    }
}

// :160-163 with EnergyDensity (:86-98): sum_x 1/2 C_ijkl eps_ij eps_kl; one partial per block
__global__ void __launch_bounds__(256)
    k_energy_density(const double* __restrict__ strain6, const __grid_constant__ Stiffness C, double* __restrict__ partial,
                     long long n) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (long long)gridDim.x * blockDim.x) {
        double e[9];
        const double e00 = strain6[x], e01 = strain6[n + x], e02 = strain6[2 * n + x];
        const double e11 = strain6[3 * n + x], e12 = strain6[4 * n + x], e22 = strain6[5 * n + x];
        e[0] = e00; e[1] = e01; e[2] = e02;
        e[3] = e01; e[4] = e11; e[5] = e12;
        e[6] = e02; e[7] = e12; e[8] = e22;
        double res = 0.0;
        for (int ij = 0; ij < 9; ++ij)
            for (int kl = 0; kl < 9; ++kl) res += C.c[ij * 9 + kl] * e[ij] * e[kl];
        acc += 0.5 * res;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { GOPF_CUDA(cudaMalloc(&p, bytes)); }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

}  // namespace

double homogeneous_modulus_energy(int rank, const int* n, const double* indicator_c128, const double* misfit9,
                                  const double* stiffness81, int device) {
    if (rank != 2 && rank != 3) throw Error("HomogeneousModulusEnergy: domain size has to be of length 2 or 3");
    FftPlan plan(rank, n, device);
    plan.use_device();
    const long long N = (long long)plan.N;
    double volume = 0.0;
    for (long long i = 0; i < N; ++i) volume += indicator_c128[2 * i];  // :102-105
    ElastParams E;
    make_elast_params(&E, stiffness81, misfit9, 3);
    Stiffness C;
    for (int i = 0; i < 81; ++i) C.c[i] = stiffness81[i];
    DevBuf ind(sizeof(cplx) * N), hat(sizeof(cplx) * N), eps(sizeof(cplx) * N), strain(sizeof(double) * 6 * N),
        partial(sizeof(double) * 1024);
    cudaStream_t s = plan.stream;
    GOPF_CUDA(cudaMemcpyAsync(ind.p, indicator_c128, sizeof(cplx) * N, cudaMemcpyHostToDevice, s));
    GOPF_CUDA(cudaMemcpyAsync(hat.p, ind.p, sizeof(cplx) * N, cudaMemcpyDeviceToDevice, s));
    plan.exec_device((cplx*)hat.p, -1, s);
    const FreqGeom fg = plan.freq_geom();
    static const int PI[6] = {0, 0, 0, 1, 1, 2}, PJ[6] = {0, 1, 2, 1, 2, 2};
    const unsigned grid = ee_grid(N);
    for (int p = 0; p < 6; ++p) {
        k_strain_hat<<<grid, 256, 0, s>>>((const cplx*)hat.p, (cplx*)eps.p, E, fg, PI[p], PJ[p], N);
        GOPF_CUDA(cudaGetLastError());
        plan.exec_device((cplx*)eps.p, +1, s);
        k_strain_store<<<grid, 256, 0, s>>>((const cplx*)eps.p, (const cplx*)ind.p, misfit9[PI[p] * 3 + PJ[p]], 1.0 / (double)N,
                                            (double*)strain.p + (size_t)p * N, N);
        GOPF_CUDA(cudaGetLastError());
    }
    k_energy_density<<<grid, 256, 0, s>>>((const double*)strain.p, C, (double*)partial.p, N);
    GOPF_CUDA(cudaGetLastError());
    std::vector<double> h(grid);
    GOPF_CUDA(cudaMemcpyAsync(h.data(), partial.p, sizeof(double) * grid, cudaMemcpyDeviceToHost, s));
    GOPF_CUDA(cudaStreamSynchronize(s));
    double energy = 0.0;
    for (double v : h) energy += v;
    return energy / volume;
}

}  // namespace gopf

extern "C" {

int gopf_elasticity_homogeneous_modulus_energy(int rank, const int* n, const double* indicator_c128, const double* misfit9,
                                               const double* stiffness81, int device, double* energy) {
    GOPF_API_BEGIN
    if (!n || !indicator_c128 || !misfit9 || !stiffness81 || !energy) throw gopf::Error("NULL argument");
    *energy = gopf::homogeneous_modulus_energy(rank, n, indicator_c128, misfit9, stiffness81, device);
    GOPF_API_END
}

int gopf_elasticity_strain_factor(const double* stiffness81, const double* misfit9, const double* freq3, int64_t count,
                                  int i, int j, double* out) {
    GOPF_API_BEGIN
    if (!stiffness81 || !misfit9 || !freq3 || !out) throw gopf::Error("NULL argument");
    if (i < 0 || j < 0 || i > 2 || j > 2) throw gopf::Error("strain component out of range");
    gopf::ElastParams E;
    gopf::make_elast_params(&E, stiffness81, misfit9, 3);
    for (int64_t k = 0; k < count; ++k)
        out[k] = gopf::elastic_strain_factor(E, freq3[3 * k], freq3[3 * k + 1], freq3[3 * k + 2], i, j);
    GOPF_API_END
}

}  // extern "C"
