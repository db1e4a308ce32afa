// Fused-kernel instantiations (see fused_launch.h); split by length so nvcc runs in parallel.
#include "fused_launch.h"
namespace gopf {
cudaError_t fused_kspace_1024(const PassGeom& g, int tx, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P, const FreqTabs& ft, const cplx* tw, cudaStream_t s) { return fused_kspace_n<1024>(g, tx, W, Wout, S, P, ft, tw, s); }
cudaError_t fused_real_1024(const PassGeom& g, int mode, cplx* W, cplx* ro, const DevDerived& D, double inv_n, unsigned long long step, const cplx* tw, cudaStream_t s) { return fused_real_n<1024>(g, mode, W, ro, D, inv_n, step, tw, s); }
}  // namespace gopf
