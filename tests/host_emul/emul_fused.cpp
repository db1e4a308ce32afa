// Test infrastructure: the fused single-field step (gopf_b200/csrc/step_kernels.cuh: k_fused_real,
// k_fused_kspace) compiled for the host and run with one OS thread per CUDA thread
// (cuda_shim.h, GOPF_EMUL_THREADS), in the order Solver::euler_step_fused launches the kernels.
// Built by tests/test_host_emulation_fused_cpu.py with -DGOPF_KNOISE (the k-space noise generator).
#define GOPF_EMUL_THREADS 1
#include "cuda_shim.h"

#include <type_traits>

#include "step_kernels.cuh"

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

static std::vector<double> freq_axis(int n) {  // fft_plan.cu: f[idx] = idx / n, -1 if > 0.5 (fftWrap.go:62-71)
    std::vector<double> f(n);
    for (int i = 0; i < n; ++i) {
        double v = (double)i / (double)n;
        if (v > 0.5) v -= 1.0;
        f[i] = v;
    }
    return f;
}

template <int N>
static void plain_pass(const PassGeom& g, const PassIO& io, const cplx* tw, int tx) {
    if (g.B == 1) {
        unsigned grid = 0, block = 0;
        size_t smem = 0;
        contig_config_n<N>(g.A, &grid, &block, &smem);
        emul_launch(k_pass_contig<N>, grid, block, g, io, tw);
    } else if (tx == 2) {
        emul_launch(k_pass_strided<N, 2, false>, (unsigned)(g.A * (g.B / 2)), (unsigned)(PlanFor<N>::T * 2), g, io, tw);
    } else {
        emul_launch(k_pass_strided<N, 4, false>, (unsigned)(g.A * (g.B / 4)), (unsigned)(PlanFor<N>::T * 4), g, io, tw);
    }
}

template <int N>
static void fused_real(const PassGeom& g, cplx* W, const DevDerived& D, double inv_n, unsigned long long step, const cplx* tw) {
    unsigned grid = 0, block = 0;
    size_t smem = 0;
    contig_config_n<N>(g.A, &grid, &block, &smem);  // fused_real_n_mode: the same shape
    emul_launch(k_fused_real<N, 0>, grid, block, g, W, (cplx*)nullptr, D, inv_n, step, tw);
}

template <int N>
static void fused_kspace(const PassGeom& g, cplx* W, cplx* S, const DevKProgram& P, const FreqTabs& ft, const cplx* tw, int tx,
                         bool late) {
    const unsigned tiles = (unsigned)(g.A * (g.bcount / tx));
    const unsigned threads = (unsigned)(PlanFor<N>::T * tx);
    const cplx* Win = W;
    if (late && tx == 2) emul_launch(k_fused_kspace<N, 2, false, true>, tiles, threads, g, Win, W, S, P, ft, tw);
    else if (late) emul_launch(k_fused_kspace<N, 4, false, true>, tiles, threads, g, Win, W, S, P, ft, tw);
    else if (tx == 2) emul_launch(k_fused_kspace<N, 2, false, false>, tiles, threads, g, Win, W, S, P, ft, tw);
    else emul_launch(k_fused_kspace<N, 4, false, false>, tiles, threads, g, Win, W, S, P, ft, tw);
}

// Solver::stamp_noise_step
static void stamp_noise_step(DevKProgram* P, unsigned long long step) {
    for (int j = 0; j < P->eq[0].n_rhs; ++j)
        if (P->eq[0].rhs[j].kind == TK_WHITE_NOISE_K) P->th[P->eq[0].rhs[j].param].K[2] = gopf_double_of(step);
}

// nsteps x Solver::euler_step_fused on a cubic (rank 3) or square (rank 2) grid of edge N
template <int N>
static int steps(int rank, const DevKProgram& P_in, const DevDerived& D, cplx* S, int nsteps, unsigned long long step0, int tx,
                 int late_request) {
    DevKProgram P = P_in;
    const int n0 = rank == 3 ? N : 1;
    const long long cells = (long long)n0 * N * N;
    const int slow = rank == 3 ? 0 : 1;
    const PassGeom gs = make_geom(n0, N, N, slow), g1 = make_geom(n0, N, N, 1), g2 = make_geom(n0, N, N, 2);
    const std::vector<cplx> tw = twiddles(N);
    const std::vector<double> f = freq_axis(N);
    FreqTabs ft;
    ft.f0 = ft.f1 = ft.f2 = f.data();
    ft.rank = rank;
    ft.off1 = 0;
    const bool late = P.fast != 0 && late_request != 0;  // LATE is launched for fast-form programs only (fused_launch.h)
    std::vector<cplx> W(cells);
    plain_pass<N>(gs, plain_io(S, W.data(), true, 1.0), tw.data(), tx);  // first inverse pass of the current spectrum
    for (int i = 0; i < nsteps; ++i) {
        stamp_noise_step(&P, step0 + i);
        if (rank == 3) plain_pass<N>(g1, plain_io(W.data(), W.data(), true, 1.0), tw.data(), tx);
        fused_real<N>(g2, W.data(), D, 1.0 / (double)cells, step0 + i, tw.data());
        if (rank == 3) plain_pass<N>(g1, plain_io(W.data(), W.data(), false, 1.0), tw.data(), tx);
        fused_kspace<N>(gs, W.data(), S, P, ft, tw.data(), tx, late);
    }
    return 0;
}

extern "C" {

int emul_fused_sizeof_program(void) { return (int)sizeof(DevKProgram); }

// program: gopf_model_fused_program_image; derived: gopf_model_derived_image of its derived field;
// S: the field's spectrum (edge^rank interleaved complex128), advanced in place
int emul_fused_steps(int rank, int edge, const void* program, const void* derived, double* S, int nsteps,
                     unsigned long long step0, int tx, int late) {
    DevKProgram P;
    memcpy(&P, program, sizeof(P));
    DevDerived D;
    memcpy(&D, derived, sizeof(D));
    cplx* s = reinterpret_cast<cplx*>(S);
    if (tx != 2 && tx != 4) return 3;
    switch (edge) {
        case 8: return steps<8>(rank, P, D, s, nsteps, step0, tx, late);
        case 16: return steps<16>(rank, P, D, s, nsteps, step0, tx, late);
        case 32: return steps<32>(rank, P, D, s, nsteps, step0, tx, late);
        case 64: return steps<64>(rank, P, D, s, nsteps, step0, tx, late);
        default: return 1;
    }
}

}  // extern "C"
