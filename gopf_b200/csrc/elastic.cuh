// Khachaturyan homogeneous-modulus elasticity in k-space.
//
// pf.HomogeneousModulusLinElast.Construct (pf/homoLinElast.go:47-99) builds, per evaluation,
//   force_c  = -i 2 pi sum_j sigma*_cj f_j H^            (elasticity/effectiveForce.go:26-35; c < Dim)
//   G u^     = force,  G_mn = (2 pi)^2 sum_jl C_mjnl f_j f_l   (elasticity/linearElasticity.go:16-63)
//   eps^_mn  = i pi (f'_n u^_m + f'_m u^_n),  f' = f with |f| = 1/2 zeroed  (linearElasticity.go:67-83)
//   term     = sum_{i<=j<Dim} w_ij A_ij FFT(H'(phi) IFFT(eps^_ij)/N) - 2 E FFT(H H'),  A = sigma* = C:eps*
// Every quantity before the inverse transform is H^(k) times a function of k only, and the
// inverse transform is linear, so the dim(dim+1)/2 strain round trips collapse to ONE:
//   term = FFT( H'(phi) IFFT(M(k) H^)/N - 2 E H(phi) H'(phi) ),
//   M(k) = 1/2 sum_{i<=j<Dim} w_ij A_ij (f'_j h_i + f'_i h_j),   h = Gamma^-1 b,
//   Gamma_mn = sum_jl C_mjnl f_j f_l,  b_c = sum_j sigma*_cj f_j (c < Dim; 0 otherwise)
// (the factors -i 2 pi, (2 pi)^-2 and i pi multiply to the real 1/2).  M(k) is real.
#pragma once
#include "cplx.cuh"

namespace gopf {

struct ElastParams {
    // Gamma_mn = sum_p K[m*3+n][p] * q_p,  q = (f0^2, f1^2, f2^2, f0 f1, f0 f2, f1 f2)
    double K[9][6];
    double sigma[9];  // sigma* = C : misfit (= A, homoLinElast.go:67), row-major 3x3
    int dim, pad;
};

// gonum mat.Dense.Solve (LU) restated as the adjugate solve of the 3x3 system Gamma h = b.
__host__ __device__ __forceinline__ double elastic_multiplier(const ElastParams& E, double f0, double f1, double f2) {
    // linearElasticity.go:43-51: the zero mode keeps u = 0
    if (fabs(f0) < 1e-10 && fabs(f1) < 1e-10 && fabs(f2) < 1e-10) return 0.0;
    const double q[6] = {f0 * f0, f1 * f1, f2 * f2, f0 * f1, f0 * f2, f1 * f2};
    double g[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) {
        double s = 0.0;
#pragma unroll
        for (int p = 0; p < 6; ++p) s = fma(E.K[e][p], q[p], s);
        g[e] = s;
    }
    const double f[3] = {f0, f1, f2};
    double b[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
        b[c] = (c < E.dim) ? E.sigma[c * 3 + 0] * f0 + E.sigma[c * 3 + 1] * f1 + E.sigma[c * 3 + 2] * f2 : 0.0;
    // adjugate (cofactor transpose)
    const double c00 = g[4] * g[8] - g[5] * g[7], c01 = g[5] * g[6] - g[3] * g[8], c02 = g[3] * g[7] - g[4] * g[6];
    const double c10 = g[2] * g[7] - g[1] * g[8], c11 = g[0] * g[8] - g[2] * g[6], c12 = g[1] * g[6] - g[0] * g[7];
    const double c20 = g[1] * g[5] - g[2] * g[4], c21 = g[2] * g[3] - g[0] * g[5], c22 = g[0] * g[4] - g[1] * g[3];
    const double det = g[0] * c00 + g[1] * c01 + g[2] * c02;
    const double inv = 1.0 / det;
    double h[3];
    h[0] = (c00 * b[0] + c10 * b[1] + c20 * b[2]) * inv;
    h[1] = (c01 * b[0] + c11 * b[1] + c21 * b[2]) * inv;
    h[2] = (c02 * b[0] + c12 * b[1] + c22 * b[2]) * inv;
    double fp[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) fp[c] = (fabs(fabs(f[c]) - 0.5) < 1e-10) ? 0.0 : f[c];
    double m = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
            if (j >= E.dim) continue;
            const double w = (i == j) ? 1.0 : 2.0;
            m += w * E.sigma[i * 3 + j] * (fp[j] * h[i] + fp[i] * h[j]);
        }
    return 0.5 * m;
}

// host: fold the rank-4 stiffness (index i*27 + j*9 + k*3 + l, elasticity/rank4.go:22-24) and the
// misfit strain into ElastParams; returns the misfit energy density 1/2 C:eps*:eps*
// (elasticity/linearElasticity.go:86-98).
inline double make_elast_params(ElastParams* E, const double* C, const double* misfit, int dim) {
    static const int pj[6] = {0, 1, 2, 0, 0, 1}, pl[6] = {0, 1, 2, 1, 2, 2};
    for (int m = 0; m < 3; ++m)
        for (int n = 0; n < 3; ++n)
            for (int p = 0; p < 6; ++p) {
                const int j = pj[p], l = pl[p];
                double v = C[m * 27 + j * 9 + n * 3 + l];
                if (j != l) v += C[m * 27 + l * 9 + n * 3 + j];
                E->K[m * 3 + n][p] = v;
            }
    double energy = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;  // rank4.go:62-75 ContractLast
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l < 3; ++l) s += C[i * 27 + j * 9 + k * 3 + l] * misfit[k * 3 + l];
            E->sigma[i * 3 + j] = s;
            energy += s * misfit[i * 3 + j];
        }
    E->dim = dim;
    E->pad = 0;
    return 0.5 * energy;
}

}  // namespace gopf
