// Plain-old-data "programs" the host model compiles the reference's closures into,
// and the device evaluators that run them inside the pass kernels.
//
//   DevKProgram  <- RHS{Terms,Denum} of every equation (pf/rhsBuilder.go:19-54,
//                   125-190), evaluated per k-point in declaration order
//                   (Gauss-Seidel, pf/euler.go:27-39), plus the semi-implicit divide
//                   (pf/euler.go:33) and the modal filter (pf/util.go:125-132).
//   DevDerived   <- DerivedField.Calc: monomials (pf/util.go:38-65), registered
//                   functions (pf/model.go:400-412) as an RPN over real parts, and
//                   white noise (pf/noise.go:20-23).
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#include "cplx.cuh"

// Device pointers of the program (modal-filter table, VolumeConservingLP multipliers).  The
// run-time specialisation (jit.cu) defines these as literal addresses so that the program itself
// can be a compile-time constant.
#ifndef GOPF_FILTER
#define GOPF_FILTER(P) ((P).filter)
#define GOPF_LP(P, slot) ((P).lp_multiplier[slot])
#endif
// Loops over the program (terms, peaks, powers): their trip counts are literals in the run-time
// specialisation, which asks for full unrolling; the library build leaves them rolled.
#ifndef GOPF_JIT_UNROLL
#define GOPF_JIT_UNROLL
#endif

namespace gopf {

#define GOPF_MAX_FIELDS 4
#define GOPF_MAX_SPECTRA 16
#define GOPF_MAX_TERMS 8
#define GOPF_MAX_FACTORS 4
#define GOPF_MAX_RPN 64
#define GOPF_MAX_PEAKS 4
#define GOPF_MAX_SPECIAL 2
#define GOPF_RPN_STACK 12
#define GOPF_MAX_POLY 8

enum TermKind {
    TK_MONOMIAL = 0,       // coef * [brick] * L^lap                       rhsBuilder.go:158-188
    TK_SPECTRAL_VISC = 1,  // coef * (-eps Q(|f|) |f|^p) * L^lap            spectralViscosity.go:44-54
    TK_PAIR_CORR = 2,      // coef * (-A C2(2 pi |f|)) [* brick] * L^lap    pairCorrelationTerm.go:37-51,96-110
    TK_CONS_NOISE = 3,     // coef * sum_c 2i sin(pi f_c) xi_c * L^lap      noise.go:60-78
    TK_VOLUME_LP = 4,      // coef * lambda * [brick] * L^lap               volumeConserving.go:19-29
    TK_TENSOR_HESSIAN = 5, // coef * (-4 pi^2 sum_ij K_ij f_i f_j) * L^lap  tensorialHessian.go:38-63
    TK_WHITE_NOISE_K = 6   // coef * xi^(k) * L^lap: the spectrum of WhiteNoise (noise.go:20-23) generated in k-space
};

struct DevTerm {
    double cre, cim;  // sign * prod scalar^p (Go cmplx.Pow semantics, evaluated on the host)
    int brick;        // spectrum index or -1
    int lap;          // total Laplacian power (own LAP^n plus LAP prefixes)
    int kind;         // TermKind
    int param;        // index into the matching special-parameter array
};

struct DevEquation {
    int n_rhs, n_den;
    DevTerm rhs[GOPF_MAX_TERMS];
    DevTerm den[GOPF_MAX_TERMS];
};

struct SpectralViscParams {
    double eps, threshold;
    int power, pad;
};

struct PairCorrParams {
    double prefactor, eff_temp;
    int n_peaks, pad;
    double plane_density[GOPF_MAX_PEAKS], location[GOPF_MAX_PEAKS], width[GOPF_MAX_PEAKS];
    double num_planes[GOPF_MAX_PEAKS];
};

struct TensorHessianParams {
    double K[9];  // row-major d x d (d = 3 when 9 coefficients were given, else 2; tensorialHessian.go:65-71)
    int d, pad;
    // A TK_WHITE_NOISE_K term borrows a slot of this array (no layout change for the other kernels):
    // K[0] = std * sqrt(N), K[1] = bits of the 64-bit seed, K[2] = bits of the step counter, which the
    // solver stamps before every step.
};

__host__ __device__ inline unsigned long long gopf_bits_of(double v) {
    union { double d; unsigned long long u; } c;
    c.d = v;
    return c.u;
}
__host__ __device__ inline double gopf_double_of(unsigned long long u) {
    union { double d; unsigned long long u; } c;
    c.u = u;
    return c.d;
}

struct ConsNoiseParams {
    int dim;
    int brick[3];  // spectrum index of the current component fields
};

struct DevKProgram {
    int rank, n_fields;
    double dt;
    DevEquation eq[GOPF_MAX_FIELDS];
    const double* filter;  // modal-filter table on the device (NULL: no filter)
    int filter_n, pad;
    SpectralViscParams sv[GOPF_MAX_SPECIAL];
    PairCorrParams pc[GOPF_MAX_SPECIAL];
    ConsNoiseParams cn[GOPF_MAX_SPECIAL];
    TensorHessianParams th[GOPF_MAX_SPECIAL];
    const double* lp_multiplier[GOPF_MAX_SPECIAL];  // device scalars (VolumeConservingLP.Multiplier)
    // Single-field fast form (fused kernels): when every term of eq[0] is a monomial with a
    // real coefficient the update collapses to polynomials in L = -(2 pi |f|)^2:
    //   c^ <- (c^ + dt*(p_nl(L)*g^ + p_self(L)*c^)) / (1 - dt*q(L))
    // fast == 2, the tabulated form: the explicit side is still polynomial (plus at most one k-space white-noise
    // term), the implicit side and the modal filter are arbitrary but real and time-independent, tabulated once:
    //   c^ <- (b(L)*c^ + a(L)*g^ + n(L)*xi^(k)) * dtab[k],   dtab = filter / (1 - dt*den),  b = 1 + dt*p_self
    // (phase-field crystal: pair-correlation + ideal-mixture terms, Vandeven filter, white noise: cfg 5).
    int fast, deg_nl, deg_self, deg_q;
    double p_nl[GOPF_MAX_POLY + 1], p_self[GOPF_MAX_POLY + 1], q[GOPF_MAX_POLY + 1];
    double fa[5], fq[5];  // fast form folded with dt: fa = dt*p_nl, fq = 1 - dt*q (degree <= 4)
    double fself[5], fnz[5];  // tabulated form: 1 + dt*p_self, dt * (noise coefficient) * L^lap
    int noise_param, pad2;    // tabulated form: P.th slot of the white-noise term (-1: none)
    const double* dtab;       // tabulated form: one real per k-point, same order as the spectrum
};

// ---- k-point ---------------------------------------------------------------------
struct KPoint {
    double f[3];  // reference Freq components [row, col, depth]
    double frad;  // sqrt(Dot(f, f))
    double L;     // -(2 pi |f|)^2, the LaplacianN base (pf/diffOp.go:27)
};

#define GOPF_PI 3.14159265358979323846

__device__ __forceinline__ KPoint make_kpoint(double f0, double f1, double f2) {
    KPoint kp;
    kp.f[0] = f0;
    kp.f[1] = f1;
    kp.f[2] = f2;
    kp.frad = sqrt(f0 * f0 + f1 * f1 + f2 * f2);
    const double k = 2.0 * GOPF_PI * kp.frad;
    kp.L = -(k * k);
    return kp;
}

__device__ __forceinline__ double ipow(double x, int n) {
    double r = 1.0;
    GOPF_JIT_UNROLL
    for (int i = 0; i < n; ++i) r *= x;
    return r;
}

// pf/spectralViscosity.go:31-40
__device__ __forceinline__ double sv_interpolant(double f, double peak) {
    const double frac = 1.0 / 3.0;
    if (f < frac * peak) return 0.0;
    if (f > peak) return 1.0;
    const double x = 1.5 * (f - frac * peak) / peak;
    return 2.0 * x * x - 3.0 * x * x * x;
}

// pfc/pairCorrelation.go:27-37
__device__ __forceinline__ double pair_corr_eval(const PairCorrParams& p, double k) {
    double result = 0.0;
    GOPF_JIT_UNROLL
    for (int i = 0; i < p.n_peaks; ++i) {
        const double pref = exp(-p.eff_temp * p.eff_temp * k * k / (2.0 * p.plane_density[i] * p.num_planes[i]));
        const double z = (k - p.location[i]) / p.width[i];
        const double value = pref * exp(-0.5 * (z * z));
        if (value > result) result = value;
    }
    return result;
}

// pf/vandeven.go:30-40 applied at x = |f| * 2 / pi (pf/util.go:125-132)
__device__ __forceinline__ double filter_eval(const double* __restrict__ tab, int n, double x) {
    const double nn = (double)(n - 1);
    const int idx = (int)(x * nn);
    if (idx >= n - 1) return tab[n - 1];
    const double dx = 1.0 / nn;
    const double y0 = tab[idx];
    const double dy = tab[idx + 1] - y0;
    const double x0 = (double)idx * dx;
    return y0 + (x - x0) * dy / dx;
}

#ifdef GOPF_KNOISE
__device__ __forceinline__ cplx knoise_value(double amp, unsigned long long seed, unsigned long long step, const KPoint& kp);
#endif

// One RHS / denominator term at one k-point.  `get(brick)` returns the current
// spectrum value of that brick at this k-point.
template <class Get>
__device__ __forceinline__ cplx eval_term(const DevKProgram& P, const DevTerm& t, const KPoint& kp, Get get) {
    cplx val;
    switch (t.kind) {
        case TK_SPECTRAL_VISC: {
            const SpectralViscParams& s = P.sv[t.param];
            val = mk(-s.eps * sv_interpolant(kp.frad, s.threshold) * ipow(kp.frad, s.power), 0.0);
            break;
        }
        case TK_PAIR_CORR: {
            const PairCorrParams& p = P.pc[t.param];
            const double m = -(p.prefactor * pair_corr_eval(p, 2.0 * GOPF_PI * kp.frad));
            val = mk(m, 0.0);
            if (t.brick >= 0) val = get(t.brick) * m;
            break;
        }
        case TK_CONS_NOISE: {
            const ConsNoiseParams& c = P.cn[t.param];
            val = mk(0.0, 0.0);
            GOPF_JIT_UNROLL
            for (int comp = 0; comp < c.dim; ++comp) {
                const double f = kp.f[comp];
                if (fabs(fabs(f) - 0.5) > 1e-6) {
                    const double s2 = 2.0 * sin(GOPF_PI * f);
                    const cplx xi = get(c.brick[comp]);
                    val += mk(-s2 * xi.y, s2 * xi.x);  // (0 + i s2) * xi
                }
            }
            break;
        }
        case TK_TENSOR_HESSIAN: {
            const TensorHessianParams& h = P.th[t.param];
            double acc = 0.0;
            GOPF_JIT_UNROLL
            for (int j = 0; j < P.rank; ++j) acc += -4.0 * GOPF_PI * GOPF_PI * kp.f[j] * kp.f[j] * h.K[j * h.d + j];
            GOPF_JIT_UNROLL
            for (int j = 0; j < P.rank; ++j)
                GOPF_JIT_UNROLL
                for (int k = j + 1; k < P.rank; ++k) acc += -8.0 * GOPF_PI * GOPF_PI * kp.f[j] * kp.f[k] * h.K[j * h.d + k];
            val = mk(acc, 0.0);
            break;
        }
#ifdef GOPF_KNOISE
        case TK_WHITE_NOISE_K: {
            const TensorHessianParams& h = P.th[t.param];
            val = knoise_value(h.K[0], gopf_bits_of(h.K[1]), gopf_bits_of(h.K[2]), kp);
            break;
        }
#endif
        case TK_VOLUME_LP: {
            const double lam = *GOPF_LP(P, t.param);
            val = get(t.brick) * lam;
            break;
        }
        default:
            val = (t.brick >= 0) ? get(t.brick) : mk(1.0, 0.0);
            break;
    }
    val = val * mk(t.cre, t.cim);
    if (t.lap > 0) val = val * ipow(kp.L, t.lap);
    return val;
}

// Semi-implicit Euler update of field `i` at one k-point (pf/euler.go:28-38).
template <class Get>
__device__ __forceinline__ cplx euler_update(const DevKProgram& P, int i, const KPoint& kp, cplx d, Get get) {
    const DevEquation& q = P.eq[i];
    cplx rhs = mk(0.0, 0.0), den = mk(0.0, 0.0);
    GOPF_JIT_UNROLL
    for (int j = 0; j < q.n_rhs; ++j) rhs += eval_term(P, q.rhs[j], kp, get);
    GOPF_JIT_UNROLL
    for (int j = 0; j < q.n_den; ++j) den += eval_term(P, q.den[j], kp, get);
    const cplx num = mk(d.x + P.dt * rhs.x, d.y + P.dt * rhs.y);
    const cplx dn = mk(1.0 - P.dt * den.x, -P.dt * den.y);
    cplx r = cdiv(num, dn);
    if (GOPF_FILTER(P)) {
        const double s = filter_eval(GOPF_FILTER(P), P.filter_n, kp.frad * 2.0 / GOPF_PI);
        r = mk(r.x * s, r.y * s);
    }
    return r;
}

// ---- derived fields (real space) ---------------------------------------------------
enum DerivedKind { DK_MONOMIAL = 0, DK_RPN = 1, DK_WHITE_NOISE = 2, DK_TABLE = 3 };

enum RpnOp {
    OP_CONST = 0,   // push arg
    OP_FIELD_RE,    // push real(field[(int)arg])
    OP_FIELD_IM,
    OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG,
    OP_POWI,        // x^(int)arg
    OP_POW,         // pop y, x -> pow(x, y)
    OP_H,           // 3x^2 - 2x^3        (pf/homoLinElast.go:10-12)
    OP_DH,          // 6x - 6x^2          (pf/homoLinElast.go:15-17)
    OP_LANDAU,      // x^2 - 2x^3 + x^4
    OP_DLANDAU,     // 2x - 6x^2 + 4x^3
    OP_EXP, OP_LOG, OP_SIN, OP_COS, OP_TANH, OP_SQRT, OP_ABS,
    OP_NEGPART      // min(x, 0): NegativeValuePenalty acts on negative values only (pf/negative_value_penalty.go:15-21)
};

struct DevDerived {
    int kind;
    int n_factors;
    int field[GOPF_MAX_FACTORS];
    double power[GOPF_MAX_FACTORS];
    int ipower[GOPF_MAX_FACTORS];  // >= 0: integer power, -1: use the polar form
    int n_ops, pad;
    unsigned char op[GOPF_MAX_RPN];
    double arg[GOPF_MAX_RPN];
    double noise_std;
    unsigned long long seed;  // white noise: Philox key; counter = (step, node)
    const double* table;      // DK_TABLE: prescribed real values, row (step mod table_steps)
    long long table_steps;
    long long table_n;
    // DK_RPN whose expression is a real polynomial of re(field 0) alone (model.cu detects it): the fused
    // real-space kernels evaluate it by Horner on the register-resident line instead of interpreting the program
    int poly_deg, pad3;       // -1: not a polynomial
    double poly[GOPF_MAX_POLY + 1];
};

// Go cmplx.Pow(x, p) for real p (math/cmplx/pow.go), polar form
__device__ __forceinline__ cplx go_cpow(cplx x, double p) {
    if (x.x == 0.0 && x.y == 0.0) {
        if (p == 0.0) return mk(1.0, 0.0);
        if (p < 0.0) return mk(1.0 / 0.0, 0.0);
        return mk(0.0, 0.0);
    }
    const double modulus = hypot(x.x, x.y);
    const double r = pow(modulus, p);
    const double theta = p * atan2(x.y, x.x);
    double s, c;
    sincos(theta, &s, &c);
    return mk(r * c, r * s);
}

__device__ __forceinline__ cplx cpow_int(cplx x, int n) {
    cplx r = mk(1.0, 0.0);
    for (int i = 0; i < n; ++i) r = r * x;
    return r;
}

// single-precision log / sincos: the hardware approximations on the device, libm on the host (tests/host_emul)
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float gopf_fast_logf(float x) { return __logf(x); }
__device__ __forceinline__ void gopf_fast_sincosf(float x, float* s, float* c) { __sincosf(x, s, c); }
#else
__device__ __forceinline__ float gopf_fast_logf(float x) { return logf(x); }
__device__ __forceinline__ void gopf_fast_sincosf(float x, float* s, float* c) {
    *s = sinf(x);
    *c = cosf(x);
}
#endif

// Philox-4x32-10 (Salmon et al., SC'11) -> two standard normals by Box-Muller.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
}

__device__ __forceinline__ double philox_normal(unsigned long long seed, unsigned long long step,
                                                unsigned long long node) {
    uint32_t c[4] = {(uint32_t)node, (uint32_t)(node >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    const unsigned long long a = ((unsigned long long)c[0] << 32) | c[1];
    const unsigned long long b = ((unsigned long long)c[2] << 32) | c[3];
    const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
    const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

#ifdef GOPF_KNOISE
// The spectrum of white noise generated where it is used.  For xi(x) iid N(0, s^2) on N nodes
// (pf/noise.go:20-23), xi^(k) = sum_x xi(x) e^{-ikx} is complex Gaussian with E|xi^|^2 = N s^2,
// xi^(-k) = conj xi^(k), and real on the modes that are their own conjugate (every component 0 or
// Nyquist).  Both members of a pair {k, -k} draw from the Philox stream of the pair's canonical
// member, found from the frequency components alone (first component that is neither 0 nor +-1/2
// positive), so the generic, fused and slab-sharded kernels produce the same field.  The reference's
// stream (Go math/rand) is unpinned: parity is statistical, as for the real-space generator.
// amp = s * sqrt(N).
// One frequency component as the generator sees it: its sign class and its coordinate on a 2^-30 lattice for
// either orientation of the pair (f = i/n to 1 ulp on either member; n <= 2^20).  Depends on one component only,
// so the fused kernels build the two line-constant components once per line and the running one from the integer
// position (knoise_comp_index; bit-identical for power-of-two extents).
struct KnComp {
    int s;              // 0: the component is its own negative (0 or +-1/2); +-1: its sign
    uint32_t pos, neg;  // lattice coordinate when the pair's canonical member is k / -k
};
__device__ __forceinline__ KnComp knoise_comp(double f) {
    KnComp c;
    const bool nyq = fabs(f) == 0.5;
    c.s = (f == 0.0 || nyq) ? 0 : (f > 0.0 ? 1 : -1);
    const double g = nyq ? 0.5 : f;
    c.pos = (uint32_t)(long long)(rint(g * 1073741824.0) + 1073741824.0);
    c.neg = nyq ? c.pos : (uint32_t)(long long)(rint(-g * 1073741824.0) + 1073741824.0);
    return c;
}
// component i/N, i in (-N/2, N/2], N a power of two <= 2^20: i * 2^30 / N is an integer
template <int N>
__device__ __forceinline__ KnComp knoise_comp_index(int i) {
    KnComp c;
    const bool nyq = i == N / 2 || i == -(N / 2);
    c.s = (i == 0 || nyq) ? 0 : (i > 0 ? 1 : -1);
    const int g = (nyq ? N / 2 : i) * (int)(1073741824LL / N);
    c.pos = (uint32_t)(1073741824 + g);
    c.neg = nyq ? c.pos : (uint32_t)(1073741824 - g);
    return c;
}
// the draw at the k-point whose components (reference order row, col, depth) are x0, x1, x2
__device__ __forceinline__ cplx knoise_draw(double amp, unsigned long long seed, unsigned long long step, const KnComp& x0,
                                            const KnComp& x1, const KnComp& x2) {
    const int sgn = x0.s != 0 ? x0.s : (x1.s != 0 ? x1.s : x2.s);  // first component that is neither 0 nor +-1/2
    uint32_t c[4], k[2];
    c[0] = sgn < 0 ? x0.neg : x0.pos;
    c[1] = sgn < 0 ? x1.neg : x1.pos;
    c[2] = sgn < 0 ? x2.neg : x2.pos;
    c[3] = (uint32_t)step;
    k[0] = (uint32_t)seed;
    k[1] = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32);
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    // Box-Muller in single precision on the special-function unit: the draw is a random variate, so 2^-21
    // relative accuracy is statistically invisible, and the double-precision log / sincospi cost more fp64
    // instructions than both FFTs of the k-space kernel together.  u1 = (c0 + 1/2) 2^-32 in (0, 1]: radius up
    // to 6.7 sigma; angle = c2 as a signed 32-bit fraction of pi in [-pi, pi).
    const float u1 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f;
    const float ang = (float)(int)c[2] * 1.4629180792671596e-9f;  // pi * 2^-31
    float snf, csf;
    const double r = (double)sqrtf(-2.0f * gopf_fast_logf(u1));
    gopf_fast_sincosf(ang, &snf, &csf);
    const double sn = (double)snf, cs = (double)csf;
    if (sgn == 0) return mk(amp * r * cs, 0.0);
    const double h = amp * 0.70710678118654752440 * r;
    return mk(h * cs, sgn > 0 ? h * sn : -(h * sn));
}
__device__ __forceinline__ cplx knoise_value(double amp, unsigned long long seed, unsigned long long step, const KPoint& kp) {
    return knoise_draw(amp, seed, step, knoise_comp(kp.f[0]), knoise_comp(kp.f[1]), knoise_comp(kp.f[2]));
}
#endif

// Value of one derived field at one node.  `fld(j)` returns field j's real-space value.
template <class Fld>
__device__ __forceinline__ cplx eval_derived(const DevDerived& D, Fld fld, unsigned long long step,
                                             unsigned long long node) {
    if (D.kind == DK_MONOMIAL) {
        cplx r = mk(1.0, 0.0);
        for (int j = 0; j < D.n_factors; ++j) {
            const cplx x = fld(D.field[j]);
            r = r * (D.ipower[j] >= 0 ? cpow_int(x, D.ipower[j]) : go_cpow(x, D.power[j]));
        }
        return r;
    }
    if (D.kind == DK_WHITE_NOISE) return mk(philox_normal(D.seed, step, node) * D.noise_std, 0.0);
    if (D.kind == DK_TABLE) return mk(D.table[(step % (unsigned long long)D.table_steps) * D.table_n + node], 0.0);
    // Operand stack in local memory (L1-resident).  A register-resident variant that shifts eight
    // named registers on every push / pop was measured 45 % slower in the pointwise kernel
    // (16 moves per push against one L1 access).
    double st[GOPF_RPN_STACK];
    int sp = 0;
    for (int i = 0; i < D.n_ops; ++i) {
        const double a = D.arg[i];
        switch (D.op[i]) {
            case OP_CONST: st[sp++] = a; break;
            case OP_FIELD_RE: st[sp++] = fld((int)a).x; break;
            case OP_FIELD_IM: st[sp++] = fld((int)a).y; break;
            case OP_ADD: sp--; st[sp - 1] += st[sp]; break;
            case OP_SUB: sp--; st[sp - 1] -= st[sp]; break;
            case OP_MUL: sp--; st[sp - 1] *= st[sp]; break;
            case OP_DIV: sp--; st[sp - 1] /= st[sp]; break;
            case OP_NEG: st[sp - 1] = -st[sp - 1]; break;
            case OP_POWI: st[sp - 1] = ipow(st[sp - 1], (int)a); break;
            case OP_POW: sp--; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
            case OP_H: { const double x = st[sp - 1]; st[sp - 1] = 3.0 * x * x - 2.0 * x * x * x; break; }
            case OP_DH: { const double x = st[sp - 1]; st[sp - 1] = 6.0 * x - 6.0 * x * x; break; }
            case OP_LANDAU: { const double x = st[sp - 1]; st[sp - 1] = x * x - 2.0 * x * x * x + x * x * x * x; break; }
            case OP_DLANDAU: { const double x = st[sp - 1]; st[sp - 1] = 2.0 * x - 6.0 * x * x + 4.0 * x * x * x; break; }
            case OP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
            case OP_LOG: st[sp - 1] = log(st[sp - 1]); break;
            case OP_SIN: st[sp - 1] = sin(st[sp - 1]); break;
            case OP_COS: st[sp - 1] = cos(st[sp - 1]); break;
            case OP_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
            case OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
            case OP_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
            case OP_NEGPART: st[sp - 1] = fmin(st[sp - 1], 0.0); break;
            default: break;
        }
    }
    return mk(sp > 0 ? st[sp - 1] : 0.0, 0.0);
}

// ---- single-field fast forms used inside the fused kernels ----------------------------
// The fused kernels keep a thread's 16 cells in registers.  The common cases below are
// branch-free and stay in registers; every other case goes through the general evaluators
// above in a ROLLED loop over cells staged in shared memory (see step_kernels.cuh), so the
// interpreter is instantiated once and never forces the register-resident line to spill.
__device__ __forceinline__ bool derived_is_monomial_fast(const DevDerived& D) {
    return D.kind == DK_MONOMIAL && D.n_factors == 1 && D.ipower[0] >= 0 && D.ipower[0] <= 15;
}
__device__ __forceinline__ bool derived_is_poly_fast(const DevDerived& D) { return D.kind == DK_RPN && D.poly_deg >= 0; }
__device__ __forceinline__ bool derived_is_fast(const DevDerived& D) {
    return derived_is_monomial_fast(D) || derived_is_poly_fast(D);
}
// real polynomial of the real part (registered functions like 3 a v^2 + 4 b v^3 of the PFC ideal-mixture term)
__device__ __forceinline__ cplx derived_poly(const DevDerived& D, cplx c) {
    double r = D.poly[GOPF_MAX_POLY];
#pragma unroll
    for (int i = GOPF_MAX_POLY - 1; i >= 0; --i) r = fma(r, c.x, D.poly[i]);
    return mk(r, 0.0);
}

// c^p by square-and-multiply (p <= 15): c^3 = c * (c * c)
__device__ __forceinline__ cplx derived_fast(int p, cplx c) {
    cplx r = (p & 1) ? c : mk(1.0, 0.0);
    cplx sq = c;
#pragma unroll
    for (int b = 1; b < 4; ++b) {
        if ((p >> b) == 0) break;
        sq = sq * sq;
        if ((p >> b) & 1) r = r * sq;
    }
    return r;
}

// Fast form of the single-field update (DevKProgram::fast): coefficients of two real
// degree-4 polynomials in L held in registers, no branches in the per-cell code.
//   c^ <- (c^ + dt * a(L) * g^) / (1 - dt * q(L)),  L = -(2 pi)^2 (f0^2 + f1^2 + f2^2)
// The coefficients (P.fa = dt*p_nl, P.fq = 1 - dt*q, folded on the host) are read as
// constant-bank operands of the DFMAs, so they occupy no registers.
__device__ __forceinline__ cplx fast_update(const DevKProgram& P, double frad2, cplx cur, cplx nl) {
    const double L = -(4.0 * GOPF_PI * GOPF_PI) * frad2;
    const double a = fma(fma(fma(fma(P.fa[4], L, P.fa[3]), L, P.fa[2]), L, P.fa[1]), L, P.fa[0]);
    const double d = fma(fma(fma(fma(P.fq[4], L, P.fq[3]), L, P.fq[2]), L, P.fq[1]), L, P.fq[0]);
    const double inv = 1.0 / d;
    return mk(fma(a, nl.x, cur.x) * inv, fma(a, nl.y, cur.y) * inv);
}

// Tabulated form (DevKProgram::fast == 2):  c^ <- (b(L)*c^ + a(L)*g^ + n(L)*xi^(k)) * dk,  dk = P.dtab at this k-point.
// tab_self_and_noise gives b*c^ + n*xi^ (the part a kernel folds into its staged spectrum cell in a rolled loop);
// tab_update finishes: `folded` says cur already is that sum.
__device__ __forceinline__ double tab_poly(const double (&c)[5], double L) {
    return fma(fma(fma(fma(c[4], L, c[3]), L, c[2]), L, c[1]), L, c[0]);
}
#ifdef GOPF_KNOISE
__device__ __forceinline__ cplx tab_self_and_noise(const DevKProgram& P, double frad2, const KnComp& x0, const KnComp& x1,
                                                   const KnComp& x2, cplx cur) {
    const double L = -(4.0 * GOPF_PI * GOPF_PI) * frad2;
    const double b = tab_poly(P.fself, L);
    const TensorHessianParams& h = P.th[P.noise_param];
    const cplx xi = knoise_draw(h.K[0], gopf_bits_of(h.K[1]), gopf_bits_of(h.K[2]), x0, x1, x2);
    const double c = tab_poly(P.fnz, L);
    return mk(fma(c, xi.x, b * cur.x), fma(c, xi.y, b * cur.y));
}
#endif
__device__ __forceinline__ cplx tab_update(const DevKProgram& P, double frad2, double dk, cplx cur, cplx nl, bool folded) {
    const double L = -(4.0 * GOPF_PI * GOPF_PI) * frad2;
    const double a = tab_poly(P.fa, L);
    const double b = folded ? 1.0 : tab_poly(P.fself, L);
    return mk(fma(a, nl.x, b * cur.x) * dk, fma(a, nl.y, b * cur.y) * dk);
}

}  // namespace gopf
