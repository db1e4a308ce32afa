// Test infrastructure: the device evaluators of the general path compiled for the host
// (see cuda_shim.h).  Built by tests/test_host_emulation_cpu.py with
//   g++ -O1 -ffp-contract=off -shared -fPIC -I gopf_b200/csrc tests/host_emul/emul.cpp
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
