// Run-time specialisation of registered functions (pf.Model.RegisterFunction,
// /root/reference/pf/model.go:400-412).  The RPN a GenericFunction expression compiles to
// (step_program.h DevDerived) is interpreted per cell by k_eval_derived, which is
// instruction-bound (~2 TB/s).  Here the same RPN is turned into straight-line
// CUDA C, compiled for sm_100a with NVRTC when the solver first needs it and launched through the
// driver API.  libnvrtc and libcuda are opened with dlopen at first use, so the library has no
// link-time dependency on either and still loads on a machine without a GPU; any failure (library
// missing, compile or load error) leaves the interpreter kernel in charge -- both are CUDA paths.
//
// The k-space update (pf/euler.go:27-39, k_update_generic) is specialised the other way round:
// no code is generated.  NVRTC compiles the very headers the library is built from (kupdate.cuh,
// step_program.h, cplx.cuh, embedded at build time) with the solver's DevKProgram laid down as a
// constant word array, the grid geometry, the node count and the device addresses of the filter
// table / multipliers as literals; constant folding then removes the term-list loops, the
// TermKind switches and turns the Freq divisions into multiplications.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "kupdate.cuh"
#include "step_program.h"

namespace gopf {
namespace jit {

// The per-cell function of one DK_RPN derived field as portable C (also valid CUDA C):
//   static inline double gopf_expr(double r0, double i0, ..., double r3, double i3)
// r_f / i_f = real / imaginary part of field f at the cell.  `used_mask` gets bit f set when field f
// is read.  Throws on a malformed program.
std::string expression_source(const DevDerived& D, unsigned* used_mask, unsigned* imag_mask = nullptr);

// Full translation unit: gopf_expr + extern "C" __global__ gopf_jit_derived(f0..f3, out, n)
std::string derived_kernel_source(const DevDerived& D, unsigned* used_mask);

// The first forward pass of a registered function's transform with the function evaluated in its
// load (fft_kernels.cuh pass_load_line, hook GOPF_JIT_LOAD_LINE): the library's own k_pass_contig<N>
// recompiled with a generated loader, so the function costs no pointwise kernel and no 32-B round
// trip of its values.  *name_expr receives the C++ name of the kernel instance for
// nvrtcGetLoweredName.
std::string derived_pass_source(const DevDerived& D, int N, std::string* name_expr);

// Translation unit of the specialised k-space kernels: extern "C" __global__
// gopf_jit_kupdate(SpectraPtrs sp, ImplicitTab tab), gopf_jit_rk4_rhs(SpectraPtrs sp, SpectraPtrs kout),
// gopf_jit_rk4_point(int mode, double fdt, SpectraPtrs field, initial, final_, kf).  P.filter / P.lp_multiplier are taken as the
// literal device addresses to bake in; tab_mask bit i = field i has a tabulated implicit factor.
std::string kupdate_kernel_source(const DevKProgram& P, const FreqGeom& fg, long long n, unsigned tab_mask);

// NVRTC: CUDA C -> sm_100a cubin.  Returns false (with the log) when NVRTC is unavailable or the
// source does not compile.  Needs no GPU.
// name_expr / lowered: optional C++ name of a kernel template instance and its mangled name in the image.
bool compile_cubin(const std::string& source, std::vector<char>* cubin, std::string* log, const std::string* name_expr = nullptr,
                   std::string* lowered = nullptr);

struct Kernel;  // a loaded module + function
// load on the current device (the CUDA runtime's primary context must be current); NULL on failure
Kernel* load(const std::vector<char>& cubin, const char* entry, std::string* log);
void unload(Kernel* k);
// launch with a 1-D grid; false on failure.  smem = dynamic shared memory in bytes (opted in above 48 KB)
bool launch(Kernel* k, unsigned grid, unsigned block, void** args, cudaStream_t stream, std::string* log, size_t smem = 0);

bool inpass_enabled();  // GOPF_JIT_INPASS=1: also evaluate registered functions inside their first forward pass
bool enabled();  // default on; GOPF_JIT=0 in the environment keeps the interpreter kernels for new solvers

}  // namespace jit
}  // namespace gopf
