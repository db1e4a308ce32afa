// Test infrastructure: the device evaluators of the general path compiled for the host
// (see cuda_shim.h).  Built by tests/test_host_emulation_cpu.py with
//   g++ -O1 -ffp-contract=off -DGOPF_KNOISE -shared -fPIC -I gopf_b200/csrc tests/host_emul/emul.cpp
#include "cuda_shim.h"

#include "kupdate.cuh"

using namespace gopf;

extern "C" {

int emul_sizeof_program(void) { return (int)sizeof(DevKProgram); }
int emul_sizeof_derived(void) { return (int)sizeof(DevDerived); }

// update_all (pf/euler.go:27-39) over n nodes.  spectra: GOPF_MAX_SPECTRA pointers to interleaved
// complex128 arrays (NULL where unused), updated in place; tabs: GOPF_MAX_FIELDS pointers or NULL.
void emul_update(const void* program, double** spectra, const double** tabs, const double* filter, int filter_n,
                 const double* lp0, const double* lp1, int rank, int d0, int d1, int d2, long long n) {
    DevKProgram P;
    memcpy(&P, program, sizeof(P));
    P.filter = filter;
    P.filter_n = filter ? filter_n : 0;
    P.lp_multiplier[0] = lp0;
    P.lp_multiplier[1] = lp1;
    SpectraPtrs sp;
    for (int i = 0; i < GOPF_MAX_SPECTRA; ++i) sp.s[i] = reinterpret_cast<cplx*>(spectra[i]);
    ImplicitTab tab;
    for (int i = 0; i < GOPF_MAX_FIELDS; ++i) tab.t[i] = tabs ? reinterpret_cast<const cplx*>(tabs[i]) : nullptr;
    FreqGeom fg;
    fg.rank = rank;
    fg.d0 = d0;
    fg.d1 = d1;
    fg.d2 = d2;
    update_all(P, sp, tab, fg, n);
}

static DevKProgram load_program(const void* program, const double* filter, int filter_n, const double* lp0, const double* lp1) {
    DevKProgram P;
    memcpy(&P, program, sizeof(P));
    P.filter = filter;
    P.filter_n = filter ? filter_n : 0;
    P.lp_multiplier[0] = lp0;
    P.lp_multiplier[1] = lp1;
    return P;
}

static SpectraPtrs load_spectra(double** spectra) {
    SpectraPtrs sp;
    for (int i = 0; i < GOPF_MAX_SPECTRA; ++i) sp.s[i] = spectra ? reinterpret_cast<cplx*>(spectra[i]) : nullptr;
    return sp;
}

static FreqGeom geom(int rank, int d0, int d1, int d2) {
    FreqGeom fg;
    fg.rank = rank;
    fg.d0 = d0;
    fg.d1 = d1;
    fg.d2 = d2;
    return fg;
}

// implicit_table_all: filter(k) / (1 - dt den_i(k)) for every node
void emul_implicit_table(const void* program, const double* filter, int filter_n, int i, double* out, int rank, int d0,
                         int d1, int d2, long long n) {
    const DevKProgram P = load_program(program, filter, filter_n, nullptr, nullptr);
    implicit_table_all(P, i, reinterpret_cast<cplx*>(out), geom(rank, d0, d1, d2), n);
}

// volume_lp_update (pf/volumeConserving.go:31-50) on state = {multiplier, integral, first flag}
void emul_volume_lp_update(double* state, const double* field_spec, const double* indicator_spec, double dt) {
    volume_lp_update(state, reinterpret_cast<const cplx*>(field_spec), reinterpret_cast<const cplx*>(indicator_spec), dt);
}

// the RK4 passes (pf/rk4.go:29-127)
void emul_rk4_rhs(const void* program, double** spectra, double** kout, const double* lp0, const double* lp1, int rank,
                  int d0, int d1, int d2, long long n) {
    const DevKProgram P = load_program(program, nullptr, 0, lp0, lp1);
    rk4_rhs_all(P, load_spectra(spectra), load_spectra(kout), geom(rank, d0, d1, d2), n);
}

void emul_rk4_point(const void* program, const double* filter, int filter_n, const double* lp0, const double* lp1, int mode,
                    double fdt, double** field, double** initial, double** final_, double** kf, int rank, int d0, int d1,
                    int d2, long long n) {
    const DevKProgram P = load_program(program, filter, filter_n, lp0, lp1);
    rk4_point_all(P, mode, fdt, load_spectra(field), load_spectra(initial), load_spectra(final_), load_spectra(kf),
                  geom(rank, d0, d1, d2), n);
}

// Solver's stamp_noise_step (solver.cu): the step counter of the k-space noise stream lives in the program
void emul_stamp_noise_step(void* program, unsigned long long step) {
    DevKProgram* P = reinterpret_cast<DevKProgram*>(program);
    for (int i = 0; i < P->n_fields; ++i)
        for (int j = 0; j < P->eq[i].n_rhs; ++j)
            if (P->eq[i].rhs[j].kind == TK_WHITE_NOISE_K) P->th[P->eq[i].rhs[j].param].K[2] = gopf_double_of(step);
}

// eval_derived (pf/model.go:237-241) at every node: fields = GOPF_MAX_FIELDS pointers to real-space
// complex128 arrays, out = n complex128.
void emul_derived(const void* derived, const double** fields, const double* table, long long table_n, double* out,
                  unsigned long long step, long long n) {
    DevDerived D;
    memcpy(&D, derived, sizeof(D));
    D.table = table;
    D.table_n = table_n;
    cplx* o = reinterpret_cast<cplx*>(out);
    for (long long i = 0; i < n; ++i)
        o[i] = eval_derived(D, [&](int f) -> cplx { return reinterpret_cast<const cplx*>(fields[f])[i]; }, step,
                            (unsigned long long)i);
}

// ref_freq for every node: out[n][3]
void emul_freq(int rank, int d0, int d1, int d2, long long n, double* out) {
    FreqGeom fg;
    fg.rank = rank;
    fg.d0 = d0;
    fg.d1 = d1;
    fg.d2 = d2;
    for (long long i = 0; i < n; ++i) {
        double f[3] = {0.0, 0.0, 0.0};
        ref_freq_fast(fg, i, true, f);
        out[3 * i + 0] = f[0];
        out[3 * i + 1] = f[1];
        out[3 * i + 2] = f[2];
    }
}

}  // extern "C"
