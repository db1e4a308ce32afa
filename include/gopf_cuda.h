/*
 * gopf_cuda.h -- C ABI of libgopfcuda.so, the sm_100a implementation of gopf's
 * spectral time-stepping hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes only, no torch / C++
 * types.  Every entry point names the reference interface it replaces
 * (file:line relative to the davidkleiven/gopf tree).  The Go side binds these
 * through cgo (see INTEGRATION.md and go/); tests and bench.py bind them through
 * ctypes.
 *
 * Conventions
 *   - complex128 arrays are interleaved (re, im) doubles, exactly Go's
 *     []complex128 / FFTW's fftw_complex; `double*` arguments named *_c128 point
 *     at 2*N doubles.
 *   - every function returns 0 on success and non-zero on failure;
 *     gopf_last_error() then describes the failure (thread local).  The
 *     reference panics on these paths (pf/solver.go:58, pf/rhsBuilder.go:29,147,
 *     pf/model.go:51); the Go shim turns a non-zero status back into panic().
 *   - handles are not thread-safe, matching FFTWWrapper (shared Data buffer,
 *     pfutil/fftWrap.go:8-13) and the single-goroutine reference.
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with a non-zero status.
 */
#ifndef GOPF_CUDA_H
#define GOPF_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOPF_ABI_VERSION 1

typedef struct gopf_fft_plan gopf_fft_plan;
typedef struct gopf_model gopf_model;
typedef struct gopf_solver gopf_solver;

/* ---- library ------------------------------------------------------------- */
const char* gopf_last_error(void);
int gopf_abi_version(void);
/* number of visible CUDA devices (0 and status != 0 when the driver is absent) */
int gopf_device_count(int* count);

/* ---- pfutil index / k-table helpers (host side, integer-exact) ------------- */
/* pfutil.NodeIdx (pfutil/indexPositionConversion.go:4-22); pos = [row, col(, depth)] */
int gopf_node_idx(int rank, const int* domain_size, const int* pos, int64_t* node);
/* pfutil.Pos (pfutil/indexPositionConversion.go:24-44) */
int gopf_pos(int rank, const int* domain_size, int64_t node, int* pos_out);
/* FFTWWrapper.Freq (pfutil/fftWrap.go:57-74): out[rank] cycles/sample, Nyquist stays +0.5 */
int gopf_freq(int rank, const int* n, int64_t i, double* out);
/* FFTWWrapper.ConjugateNode (pfutil/fftWrap.go:78-95) */
int gopf_conjugate_node(int rank, const int* n, int64_t i, int64_t* out);

/* ---- transform level: pfutil.FFTWWrapper ---------------------------------- */
/* pfutil.NewFFTW(n) (pfutil/fftWrap.go:16-23).  rank 1..3, row-major, last axis
 * fastest.  device < 0 selects the current device. */
int gopf_fft_plan_create(int rank, const int* n, int device, gopf_fft_plan** out);
/* FFTWWrapper.FFT / IFFT (pfutil/fftWrap.go:26-39): in place on the caller's host
 * slice, unnormalised both ways.  sign = -1 forward, +1 inverse. */
int gopf_fft_exec(gopf_fft_plan* plan, double* host_inout_c128, int sign);
/* same on a device-resident array; stream is a cudaStream_t (NULL = plan stream) */
int gopf_fft_exec_device(gopf_fft_plan* plan, void* dev_inout_c128, int sign, void* stream);
/* Freq(i) for `count` node numbers evaluated ON THE DEVICE by the same code the
 * k-space kernels use (bit-exactness check of the device k-table). */
int gopf_fft_freq_device(gopf_fft_plan* plan, const int64_t* nodes, int64_t count, double* out);
int gopf_fft_plan_destroy(gopf_fft_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* GOPF_CUDA_H */
