// Axis-pass kernel instantiations (see fft_kernels.cuh); split by length so nvcc runs in parallel.
#include "fft_kernels.cuh"
namespace gopf {
cudaError_t launch_pass_4096(const PassGeom& g, int tx, const PassIO& io, const cplx* tw, cudaStream_t s) { return launch_pass_n<4096>(g, tx, io, tw, s); }
}  // namespace gopf
