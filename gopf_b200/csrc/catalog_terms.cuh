// Per-cell arithmetic of the catalog terms that live outside the k-space program:
// ChargeTransport (pf/chargeTransport.go:56-119) and point Sources
// (pf/sourceTerm.go:25-30).  Every function is __host__ __device__ so the same
// code is checked on the CPU through the gopf_charge_transport_* / gopf_source_*
// table entry points of the C ABI (tests/test_catalog_cpu.py) and runs inside
// the kernels of catalog_terms.cu.  Citations: /root/reference.
#pragma once
#include <math.h>

namespace gopf {

#define GOPF_CT_PI 3.14159265358979323846
#define GOPF_MAX_SOURCES 8

// ChargeTransport: device view of one registered term
struct ChargeParams {
    int dim;        // len(FT.Freq(0)) == rank of the grid
    int n_voigt;    // 3 (2-D: s_xx, s_yy, s_xy) or 6 (3-D: s_xx, s_yy, s_zz, s_xz, s_yz, s_xy)
    double ext[3];  // ExternalField
};

// Sources of one equation: sum_s amp_s * exp(-i 2 pi f . pos_s)
struct SourceParams {
    int n, rank;
    double pos[GOPF_MAX_SOURCES][3];
    double amp[GOPF_MAX_SOURCES];  // f(t) evaluated on the host for the current time
};

// voigtIndex (pf/chargeTransport.go:151-171)
__host__ __device__ inline int ct_voigt(int i, int j, int dim) {
    if (dim == 2) return i == j ? i : 2;
    if (i == j) return i;
    const int s = i + j;  // (0,1) -> 5, (0,2) -> 4, (1,2) -> 3
    return 6 - s;
}

// chargeTransport.go:64-73: effField^(k) = rho^(k) * complex(0, w), w = f_d / (2 pi f.f + 1e-16);
// 0 where abs(abs(f_d) - 0.5) <= 1e-10
__host__ __device__ inline double ct_field_multiplier(const double* f, int rank, int comp) {
    double ksq = 0.0;
    for (int k = 0; k < rank; ++k) ksq += f[k] * f[k];
    if (fabs(fabs(f[comp]) - 0.5) > 1e-10) return f[comp] / (2.0 * GOPF_CT_PI * ksq + 1e-16);
    return 0.0;
}

// chargeTransport.go:106-112: field += complex(0, 2 pi f_d2) * FFT(J_d2), Nyquist plane skipped
__host__ __device__ inline double ct_divergence_multiplier(const double* f, int comp) {
    const double k = f[comp];
    return fabs(fabs(k) - 0.5) > 1e-10 ? 2.0 * GOPF_CT_PI * k : 0.0;
}

// sourceTerm.go:25-30 for one source: complex(amp, 0) * cmplx.Exp(-complex(0, 2 pi Dot(f, pos)))
__host__ __device__ inline void source_value(const double* f, const double* pos, int rank, double amp, double* re,
                                             double* im) {
    double dot = 0.0;
    for (int k = 0; k < rank; ++k) dot += f[k] * pos[k];
    const double theta = 2.0 * GOPF_CT_PI * dot;
    double s, c;
#ifdef __CUDA_ARCH__
    sincos(-theta, &s, &c);
#else
    s = sin(-theta);
    c = cos(-theta);
#endif
    *re = amp * c;
    *im = amp * s;
}

}  // namespace gopf
