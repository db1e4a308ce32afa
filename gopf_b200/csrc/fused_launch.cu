// Dispatch of the fused single-field kernels to their per-length instantiations.
#include "fused_launch.h"
#include "tma_launch.h"

#include <cstdlib>

namespace gopf {

int env_int(const char* name, int fallback) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : fallback;
}

#define GOPF_FUSED_N(X) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

#define X(n)                                                                                                     \
    cudaError_t fused_kspace_##n(const PassGeom&, int, const cplx*, cplx*, cplx*, const DevKProgram&,            \
                                 const FreqTabs&, const cplx*, cudaStream_t);                                                   \
    cudaError_t fused_real_##n(const PassGeom&, int, cplx*, cplx*, const DevDerived&, double, unsigned long long, \
                               const cplx*, cudaStream_t);
GOPF_FUSED_N(X)
#undef X

bool fused_length_supported(int n) {
    switch (n) {
#define X(v) case v: return true;
        GOPF_FUSED_N(X)
#undef X
        default: return false;
    }
}

cudaError_t launch_fused_kspace(const PassGeom& g, int tx_want, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P,
                                const FreqTabs& ft, const cplx* tw, cudaStream_t s) {
    {
        const cudaError_t e = launch_fused_kspace_tma(g, W, Wout, S, P, ft, tw, s);
        if (e != cudaErrorNotSupported) return e;
    }
    switch (g.N) {
#define X(v) case v: return fused_kspace_##v(g, tx_want, W, Wout, S, P, ft, tw, s);
        GOPF_FUSED_N(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fused_real(const PassGeom& g, int mode, cplx* W, cplx* real_out, const DevDerived& D, double inv_n,
                              unsigned long long step, const cplx* tw, cudaStream_t s) {
    if (mode == 0 && real_out == nullptr) {
        const cudaError_t e = launch_fused_real_tma(g, W, D, inv_n, step, tw, s);
        if (e != cudaErrorNotSupported) return e;
    }
    switch (g.N) {
#define X(v) case v: return fused_real_##v(g, mode, W, real_out, D, inv_n, step, tw, s);
        GOPF_FUSED_N(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace gopf
