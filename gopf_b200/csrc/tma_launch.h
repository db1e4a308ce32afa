// Host entry points of the copy-engine-fed line kernels (tma_kernels.cuh).  Each returns
// cudaErrorNotSupported when the shape is outside what those kernels cover (the caller then takes the
// register-resident kernel of fft_kernels.cuh / step_kernels.cuh: both are device paths).
//
// Environment (tuning): GOPF_TMA=0 switches all three off; GOPF_TMA_PASS / GOPF_TMA_REAL / GOPF_TMA_KSPACE = 0
// switch one off; GOPF_TMA_MIN_N (default 1024) is the shortest line that takes them; GOPF_TMA_L2 (0..3,
// default 3) is the tensor maps' L2 promotion (none / 64 / 128 / 256 B).
#pragma once
#include "fft_kernels.cuh"
#include "step_kernels.cuh"

namespace gopf {

cudaError_t launch_pass_tma(const PassGeom& g, const PassIO& io, const cplx* tw, cudaStream_t s);
cudaError_t launch_fused_real_tma(const PassGeom& g, cplx* W, const DevDerived& D, double inv_n, unsigned long long step,
                                  const cplx* tw, cudaStream_t s);
cudaError_t launch_fused_kspace_tma(const PassGeom& g, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P,
                                    const FreqTabs& ft, const cplx* tw, cudaStream_t s);
// number of launches that took the copy-engine kernels since the last reset (tests assert the path taken)
long long tma_launch_count(bool reset);

}  // namespace gopf
