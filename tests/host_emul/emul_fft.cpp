// Test infrastructure: the axis-pass kernels of gopf_b200/csrc/fft_kernels.cuh (register / shared-memory
// Stockham engine, fft_engine.cuh) compiled for the host and run with one OS thread per CUDA thread
// (cuda_shim.h, GOPF_EMUL_THREADS).  Built by tests/test_host_emulation_fft_cpu.py:
//   g++ -std=c++17 -O1 -pthread -I gopf_b200/csrc -I tests/host_emul [-include <generated loader>] emul_fft.cpp
// When EMUL_JIT_UNIT is defined it names the translation unit jit::derived_pass_source generated (the text NVRTC
// receives), which brings fft_kernels.cuh in itself together with the loader of one registered function.
#define GOPF_EMUL_THREADS 1
#include "cuda_shim.h"

#ifdef EMUL_JIT_UNIT
#include EMUL_JIT_UNIT
#else
#include "fft_kernels.cuh"
#endif

#include <cmath>
#include <vector>

using namespace gopf;

static std::vector<cplx> twiddles(int len) {  // fft_plan.cu
    std::vector<cplx> tw(len);
    for (int j = 0; j < len; ++j) {
        long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)len;
        tw[j] = mk((double)cosl(ang), (double)sinl(ang));
    }
    tw[0] = mk(1.0, 0.0);
    if (len % 2 == 0) tw[len / 2] = mk(-1.0, 0.0);
    if (len % 4 == 0) {
        tw[len / 4] = mk(0.0, -1.0);
        tw[3 * len / 4] = mk(0.0, 1.0);
    }
    return tw;
}

template <int N>
static int run_contig(const PassGeom& g, const PassIO& io, const cplx* tw) {
    unsigned grid = 0, block = 0;
    size_t smem = 0;
    contig_config_n<N>(g.A, &grid, &block, &smem);
    if (smem > sizeof(gopf_smem_raw)) return 2;
    emul_launch(k_pass_contig<N>, grid, block, g, io, tw);
    return 0;
}

template <int N, int TX>
static int run_strided(const PassGeom& g, const PassIO& io, const cplx* tw) {
    if ((size_t)N * TX * sizeof(cplx) > sizeof(gopf_smem_raw) || g.B % TX != 0) return 2;
    emul_launch(k_pass_strided<N, TX, false>, (unsigned)(g.A * (g.B / TX)), (unsigned)(PlanFor<N>::T * TX), g, io, tw);
    return 0;
}

template <int N>
static int run_pass(const PassGeom& g, const PassIO& io, int tx) {
    const std::vector<cplx> tw = twiddles(N);
    if (g.B == 1) return run_contig<N>(g, io, tw.data());
    switch (tx) {
        case 2: return run_strided<N, 2>(g, io, tw.data());
        case 4: return run_strided<N, 4>(g, io, tw.data());
        case 8: return run_strided<N, 8>(g, io, tw.data());
        default: return 3;
    }
}

static int dispatch(const PassGeom& g, const PassIO& io, int tx) {
    switch (g.N) {
        case 4: return run_pass<4>(g, io, tx);
        case 8: return run_pass<8>(g, io, tx);
        case 16: return run_pass<16>(g, io, tx);
        case 32: return run_pass<32>(g, io, tx);
        case 64: return run_pass<64>(g, io, tx);
        case 128: return run_pass<128>(g, io, tx);
        case 256: return run_pass<256>(g, io, tx);
        case 512: return run_pass<512>(g, io, tx);
        case 1024: return run_pass<1024>(g, io, tx);
        case 2048: return run_pass<2048>(g, io, tx);
        case 4096: return run_pass<4096>(g, io, tx);
        default: return 1;
    }
}

extern "C" {

// one pass along `axis` of the row-major [n0][n1][n2] array, out of place or in place; tx = strided tile width
int emul_fft_pass(int n0, int n1, int n2, int axis, int inverse, double scale, const double* in, double* out, int tx) {
    const PassGeom g = make_geom(n0, n1, n2, axis);
    const PassIO io = plain_io(reinterpret_cast<const cplx*>(in), reinterpret_cast<cplx*>(out), inverse != 0, scale);
    return dispatch(g, io, tx);
}

// the contiguous forward pass with a derived field evaluated in its load (LK_DERIVED): the interpreter of the
// library build, or -- in a unit built with EMUL_JIT_UNIT -- the generated loader of the run-time specialisation
int emul_fft_pass_derived(int n0, int n1, int n2, const void* derived, const double** fields, unsigned long long step,
                          double* out) {
    const PassGeom g = make_geom(n0, n1, n2, 2);
    PassIO io = plain_io(reinterpret_cast<cplx*>(out), reinterpret_cast<cplx*>(out), false, 1.0);
    io.load_kind = LK_DERIVED;
    memcpy(&io.D, derived, sizeof(DevDerived));
    for (int f = 0; f < GOPF_MAX_FIELDS; ++f) io.R.r[f] = reinterpret_cast<const cplx*>(fields[f]);
    io.step = step;
    return dispatch(g, io, 0);
}

// SquaredGradient loads (pf/squareGradientTerm.go:38-65).  emul_fft_pass_gradient: one inverse pass along `axis`
// with the multiplier i 2 pi f_comp in its load -- per node from FFTWWrapper.Freq (LK_GRADIENT, line_table == NULL)
// or from the axis' own Freq table (LK_GRADIENT_LINE; only valid when `axis` carries component `comp`).
int emul_fft_pass_gradient(int n0, int n1, int n2, int axis, int rank, int d0, int d1, int d2, int comp,
                           const double* line_table, const double* in, double* out, int tx) {
    const PassGeom g = make_geom(n0, n1, n2, axis);
    PassIO io = plain_io(reinterpret_cast<const cplx*>(in), reinterpret_cast<cplx*>(out), true, 1.0);
    if (line_table) {
        io.load_kind = LK_GRADIENT_LINE;
        io.rtab = line_table;
    } else {
        io.load_kind = LK_GRADIENT;
        io.fg.rank = rank;
        io.fg.d0 = d0;
        io.fg.d1 = d1;
        io.fg.d2 = d2;
        io.comp = comp;
    }
    return dispatch(g, io, tx);
}

// forward pass along `axis` of sum_d g_d^2 (LK_SUM_SQUARES)
int emul_fft_pass_sum_squares(int n0, int n1, int n2, int axis, int dim, const double* g0, const double* g1, const double* g2,
                              double* out, int tx) {
    const PassGeom g = make_geom(n0, n1, n2, axis);
    PassIO io = plain_io(reinterpret_cast<cplx*>(out), reinterpret_cast<cplx*>(out), false, 1.0);
    io.load_kind = LK_SUM_SQUARES;
    io.dim = dim;
    io.g[0] = reinterpret_cast<const cplx*>(g0);
    io.g[1] = reinterpret_cast<const cplx*>(g1);
    io.g[2] = reinterpret_cast<const cplx*>(g2 ? g2 : g1);
    return dispatch(g, io, tx);
}

int emul_fft_has_jit_loader(void) {
#ifdef GOPF_JIT_LOAD_LINE
    return 1;
#else
    return 0;
#endif
}

}  // extern "C"
