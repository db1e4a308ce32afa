// Fused-kernel instantiations (see fused_launch.h); split by length so nvcc runs in parallel.
#include "fused_launch.h"
namespace gopf {
cudaError_t fused_kspace_4(const PassGeom& g, int tx, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P, const FreqTabs& ft, const cplx* tw, cudaStream_t s) { return fused_kspace_n<4>(g, tx, W, Wout, S, P, ft, tw, s); }
cudaError_t fused_real_4(const PassGeom& g, int mode, cplx* W, cplx* ro, const DevDerived& D, double inv_n, unsigned long long step, const cplx* tw, cudaStream_t s) { return fused_real_n<4>(g, mode, W, ro, D, inv_n, step, tw, s); }
cudaError_t fused_kspace_8(const PassGeom& g, int tx, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P, const FreqTabs& ft, const cplx* tw, cudaStream_t s) { return fused_kspace_n<8>(g, tx, W, Wout, S, P, ft, tw, s); }
cudaError_t fused_real_8(const PassGeom& g, int mode, cplx* W, cplx* ro, const DevDerived& D, double inv_n, unsigned long long step, const cplx* tw, cudaStream_t s) { return fused_real_n<8>(g, mode, W, ro, D, inv_n, step, tw, s); }
cudaError_t fused_kspace_16(const PassGeom& g, int tx, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P, const FreqTabs& ft, const cplx* tw, cudaStream_t s) { return fused_kspace_n<16>(g, tx, W, Wout, S, P, ft, tw, s); }
cudaError_t fused_real_16(const PassGeom& g, int mode, cplx* W, cplx* ro, const DevDerived& D, double inv_n, unsigned long long step, const cplx* tw, cudaStream_t s) { return fused_real_n<16>(g, mode, W, ro, D, inv_n, step, tw, s); }
cudaError_t fused_kspace_32(const PassGeom& g, int tx, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P, const FreqTabs& ft, const cplx* tw, cudaStream_t s) { return fused_kspace_n<32>(g, tx, W, Wout, S, P, ft, tw, s); }
cudaError_t fused_real_32(const PassGeom& g, int mode, cplx* W, cplx* ro, const DevDerived& D, double inv_n, unsigned long long step, const cplx* tw, cudaStream_t s) { return fused_real_n<32>(g, mode, W, ro, D, inv_n, step, tw, s); }
cudaError_t fused_kspace_64(const PassGeom& g, int tx, const cplx* W, cplx* Wout, cplx* S, const DevKProgram& P, const FreqTabs& ft, const cplx* tw, cudaStream_t s) { return fused_kspace_n<64>(g, tx, W, Wout, S, P, ft, tw, s); }
cudaError_t fused_real_64(const PassGeom& g, int mode, cplx* W, cplx* ro, const DevDerived& D, double inv_n, unsigned long long step, const cplx* tw, cudaStream_t s) { return fused_real_n<64>(g, mode, W, ro, D, inv_n, step, tw, s); }
}  // namespace gopf
