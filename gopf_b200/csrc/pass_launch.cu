// Dispatch of one axis pass to the per-length instantiations (pass_inst_*.cu).
#include "fft_kernels.cuh"
#include "tma_launch.h"

#include <cstdlib>

namespace gopf {

int pass_tx_override() {
    const char* v = std::getenv("GOPF_PASS_TX");
    return (v && *v) ? std::atoi(v) : 0;
}

int prefetch_enabled() {
    const char* v = std::getenv("GOPF_PREFETCH");
    return (v && *v) ? std::atoi(v) : 0;
}

#define GOPF_DECL(n) cudaError_t launch_pass_##n(const PassGeom&, int, const PassIO&, const cplx*, cudaStream_t);
GOPF_DECL(2) GOPF_DECL(4) GOPF_DECL(8) GOPF_DECL(16) GOPF_DECL(32) GOPF_DECL(64) GOPF_DECL(128) GOPF_DECL(256)
GOPF_DECL(512) GOPF_DECL(1024) GOPF_DECL(2048) GOPF_DECL(4096)
#undef GOPF_DECL

cudaError_t launch_pass(const PassGeom& g, int tx_want, const PassIO& io, const cplx* tw, cudaStream_t s) {
    // long strided lines: the copy-engine-fed kernel (tma_kernels.cuh) when the shape is covered
    {
        const cudaError_t e = launch_pass_tma(g, io, tw, s);
        if (e != cudaErrorNotSupported) return e;
    }
    switch (g.N) {
#define X(n) case n: return launch_pass_##n(g, tx_want, io, tw, s);
        X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

bool contig_launch_config(int N, long long A, unsigned* grid, unsigned* block, size_t* smem) {
    switch (N) {
#define X(n) case n: contig_config_n<n>(A, grid, block, smem); return true;
        X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)
#undef X
        default: return false;
    }
}

}  // namespace gopf
