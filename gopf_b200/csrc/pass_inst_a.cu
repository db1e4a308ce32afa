// Axis-pass kernel instantiations (see fft_kernels.cuh); split by length so nvcc runs in parallel.
#include "fft_kernels.cuh"
namespace gopf {
cudaError_t launch_pass_2(const PassGeom& g, int tx, const PassIO& io, const cplx* tw, cudaStream_t s) { return launch_pass_n<2>(g, tx, io, tw, s); }
cudaError_t launch_pass_4(const PassGeom& g, int tx, const PassIO& io, const cplx* tw, cudaStream_t s) { return launch_pass_n<4>(g, tx, io, tw, s); }
cudaError_t launch_pass_8(const PassGeom& g, int tx, const PassIO& io, const cplx* tw, cudaStream_t s) { return launch_pass_n<8>(g, tx, io, tw, s); }
cudaError_t launch_pass_16(const PassGeom& g, int tx, const PassIO& io, const cplx* tw, cudaStream_t s) { return launch_pass_n<16>(g, tx, io, tw, s); }
cudaError_t launch_pass_32(const PassGeom& g, int tx, const PassIO& io, const cplx* tw, cudaStream_t s) { return launch_pass_n<32>(g, tx, io, tw, s); }
cudaError_t launch_pass_64(const PassGeom& g, int tx, const PassIO& io, const cplx* tw, cudaStream_t s) { return launch_pass_n<64>(g, tx, io, tw, s); }
}  // namespace gopf
