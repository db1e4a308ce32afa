// The generic k-space update (pf/euler.go:27-39) as header-only device code: solver.cu
// instantiates it with the program as a kernel parameter (k_update_generic), and jit.cu hands
// the same text to NVRTC with the program baked in as a constant, so that the compiler folds the
// term list, the Freq divisions and the table / filter switches.
#pragma once
#include "step_program.h"

// Presence of field i's tabulated implicit factor.  The run-time specialisation defines this as a
// literal bit test.
#ifndef GOPF_TAB_PRESENT
#define GOPF_TAB_PRESENT(tab, i) ((tab).t[i] != nullptr)
#endif

namespace gopf {

// Reference k-table geometry.  Freq() decomposes the node number with
// Dimensions[1] / Dimensions[0] (pfutil/fftWrap.go:42-54), which matches the FFTW
// row-major layout for every 2-D shape and for cubic 3-D shapes only.
struct FreqGeom {
    int rank;
    int d0, d1, d2;  // reference Dimensions[0..2] (d2 = 1 for rank 2)
};

// Literal restatement of FFTWWrapper.Freq for node i (pfutil/fftWrap.go:57-74).
__host__ __device__ inline void ref_freq(const FreqGeom& g, long long i, double* res) {
    long long c = i % g.d1;
    long long r = (i / g.d1) % g.d0;
    res[1] = (double)c / (double)g.d1;
    res[0] = (double)r / (double)g.d0;
    if (g.rank > 2) {
        long long d = i / ((long long)g.d0 * g.d1);
        res[2] = (double)d / (double)g.d2;
    }
    for (int k = 0; k < g.rank; ++k)
        if (res[k] > 0.5) res[k] -= 1.0;
}

struct SpectraPtrs {
    cplx* s[GOPF_MAX_SPECTRA];  // fields first (k-space state, updated in place), then derived / work spectra
};

// generic pointwise update over every k (any shape, literal Freq from the node number)
// Freq of node idx with 32-bit index arithmetic when the grid allows it (64-bit divisions cost
// ~100 instructions each); same IEEE divides as ref_freq, so the result is bit-identical.
__device__ __forceinline__ void ref_freq_fast(const FreqGeom& g, long long idx, bool small, double* res) {
    if (!small) {
        ref_freq(g, idx, res);
        return;
    }
    const unsigned i = (unsigned)idx, d1 = (unsigned)g.d1, d0 = (unsigned)g.d0;
    const unsigned q = i / d1, c = i - q * d1;
    const unsigned d = q / d0, r = q - d * d0;
    res[1] = (double)c / (double)g.d1;
    res[0] = (double)r / (double)g.d0;
    if (g.rank > 2) res[2] = (double)d / (double)g.d2;
    GOPF_JIT_UNROLL
    for (int k = 0; k < g.rank; ++k)
        if (res[k] > 0.5) res[k] -= 1.0;
}

// ImplicitTab: per field, filter(k) / (1 - dt*den(k)) tabulated once (the implicit side and the
// modal filter depend on k only; the pair-correlation and viscosity multipliers cost exp / sqrt
// per k).  NULL entries: evaluate literally.
struct ImplicitTab {
    const cplx* t[GOPF_MAX_FIELDS];
};

// C cells per thread.  Measured at 512^3 (cfg 4): C = 4 is 40 % slower than C = 1 (registers cost
// more occupancy than the extra loads in flight give back), so the kernel runs C = 1.
template <int C>
__device__ __forceinline__ void update_cells(const DevKProgram& P, const SpectraPtrs& sp, const ImplicitTab& tab,
                                             const FreqGeom& fg, bool small, const long long (&idx)[C]) {
    KPoint kp[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq_fast(fg, idx[c], small, f);
        kp[c] = make_kpoint(f[0], f[1], f[2]);
    }
    GOPF_JIT_UNROLL
    for (int i = 0; i < P.n_fields; ++i) {
        const DevEquation& q = P.eq[i];
        cplx d[C], rhs[C], den[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            d[c] = sp.s[i][idx[c]];
            rhs[c] = den[c] = mk(0.0, 0.0);
        }
        GOPF_JIT_UNROLL
        for (int j = 0; j < q.n_rhs; ++j) {
#pragma unroll
            for (int c = 0; c < C; ++c) rhs[c] += eval_term(P, q.rhs[j], kp[c], [&](int b) -> cplx { return sp.s[b][idx[c]]; });
        }
        if (GOPF_TAB_PRESENT(tab, i)) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                sp.s[i][idx[c]] = mk(d[c].x + P.dt * rhs[c].x, d[c].y + P.dt * rhs[c].y) * tab.t[i][idx[c]];
        } else {
            GOPF_JIT_UNROLL
            for (int j = 0; j < q.n_den; ++j) {
#pragma unroll
                for (int c = 0; c < C; ++c) den[c] += eval_term(P, q.den[j], kp[c], [&](int b) -> cplx { return sp.s[b][idx[c]]; });
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const cplx num = mk(d[c].x + P.dt * rhs[c].x, d[c].y + P.dt * rhs[c].y);
                cplx r = cdiv(num, mk(1.0 - P.dt * den[c].x, -P.dt * den[c].y));  // euler.go:33
                if (GOPF_FILTER(P)) {
                    const double sc = filter_eval(GOPF_FILTER(P), P.filter_n, kp[c].frad * 2.0 / GOPF_PI);
                    r = mk(r.x * sc, r.y * sc);
                }
                sp.s[i][idx[c]] = r;  // later equations read the updated value (euler.go:27-39)
            }
        }
    }
}


// every k of the local spectrum, one cell per thread and iteration
__device__ __forceinline__ void update_all(const DevKProgram& P, const SpectraPtrs& sp, const ImplicitTab& tab,
                                           const FreqGeom& fg, long long n) {
    const bool small = n < (1LL << 31);
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        const long long idx[1] = {i};
        update_cells<1>(P, sp, tab, fg, small, idx);
    }
}

// filter(k) / (1 - dt * den_i(k)) for every node (euler.go:33, util.go:125-132)
__device__ __forceinline__ void implicit_table_all(const DevKProgram& P, int i, cplx* out, const FreqGeom& fg, long long n) {
    const bool small = n < (1LL << 31);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq_fast(fg, idx, small, f);
        const KPoint kp = make_kpoint(f[0], f[1], f[2]);
        const DevEquation& q = P.eq[i];
        cplx den = mk(0.0, 0.0);
        GOPF_JIT_UNROLL
        for (int j = 0; j < q.n_den; ++j) den += eval_term(P, q.den[j], kp, [&](int) -> cplx { return mk(1.0, 0.0); });
        cplx r = cdiv(mk(1.0, 0.0), mk(1.0 - P.dt * den.x, -P.dt * den.y));
        if (GOPF_FILTER(P)) {
            const double sc = filter_eval(GOPF_FILTER(P), P.filter_n, kp.frad * 2.0 / GOPF_PI);
            r = mk(r.x * sc, r.y * sc);
        }
        out[idx] = r;
    }
}

// VolumeConservingLP.OnStepFinished (pf/volumeConserving.go:31-50).  sum_i Re c_i is the
// DC mode of the updated spectrum, the indicator integral the DC mode of the indicator
// spectrum.  state = {multiplier, current integral, first-update flag}
__device__ __forceinline__ void volume_lp_update(double* state, const cplx* field_spec, const cplx* indicator_spec, double dt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double field_integral = field_spec[0].x;
        const double indicator_integral = indicator_spec[0].x;
        if (state[2] != 0.0) {
            state[1] = field_integral;
            state[2] = 0.0;
        } else {
            const double delta = field_integral - state[1];
            state[1] = field_integral;
            state[0] = state[0] - delta / (dt * indicator_integral);
        }
    }
}

// RK4 pointwise passes (pf/rk4.go:58-68, 77-84, 87-96, 101-111, 123-126)
__device__ __forceinline__ void rk4_rhs_all(const DevKProgram& P, SpectraPtrs sp, SpectraPtrs kout, FreqGeom fg, long long n) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq(fg, idx, f);
        const KPoint kp = make_kpoint(f[0], f[1], f[2]);
        auto get = [&](int b) -> cplx { return sp.s[b][idx]; };
        GOPF_JIT_UNROLL
        for (int i = 0; i < P.n_fields; ++i) {
            const DevEquation& q = P.eq[i];
            cplx rhs = mk(0.0, 0.0);
            GOPF_JIT_UNROLL
            for (int j = 0; j < q.n_rhs; ++j) rhs += eval_term(P, q.rhs[j], kp, get);
            kout.s[i][idx] = rhs;
        }
    }
}

// mode 0: final += fdt*k ; field = initial                       (PrepareNextCorrection)
// mode 1: field = (field + fdt*k) / (1 - fdt*den)                 (correction, first loop)
// mode 2: final /= (1 - fdt*den); field = final; filter           (end of Step)
__device__ __forceinline__ void rk4_point_all(const DevKProgram& P, int mode, double fdt, SpectraPtrs field,
                                              SpectraPtrs initial, SpectraPtrs final_, SpectraPtrs kf, FreqGeom fg,
                                              long long n) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        KPoint kp;
        if (mode != 0) {
            double f[3] = {0.0, 0.0, 0.0};
            ref_freq(fg, idx, f);
            kp = make_kpoint(f[0], f[1], f[2]);
        }
        auto get = [&](int b) -> cplx { return field.s[b][idx]; };
        GOPF_JIT_UNROLL
        for (int i = 0; i < P.n_fields; ++i) {
            if (mode == 0) {
                const cplx k = kf.s[i][idx];
                cplx fv = final_.s[i][idx];
                fv = mk(fv.x + fdt * k.x, fv.y + fdt * k.y);
                final_.s[i][idx] = fv;
                field.s[i][idx] = initial.s[i][idx];
                continue;
            }
            const DevEquation& q = P.eq[i];
            cplx den = mk(0.0, 0.0);
            GOPF_JIT_UNROLL
            for (int j = 0; j < q.n_den; ++j) den += eval_term(P, q.den[j], kp, get);
            const cplx dn = mk(1.0 - fdt * den.x, -fdt * den.y);
            if (mode == 1) {
                const cplx k = kf.s[i][idx];
                cplx fv = field.s[i][idx];
                fv = mk(fv.x + fdt * k.x, fv.y + fdt * k.y);
                field.s[i][idx] = cdiv(fv, dn);
            } else {
                cplx fv = cdiv(final_.s[i][idx], dn);
                final_.s[i][idx] = fv;
                if (GOPF_FILTER(P)) {
                    const double s = filter_eval(GOPF_FILTER(P), P.filter_n, kp.frad * 2.0 / GOPF_PI);
                    fv = mk(fv.x * s, fv.y * s);
                }
                field.s[i][idx] = fv;
            }
        }
    }
}

}  // namespace gopf
