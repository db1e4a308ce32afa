// complex128 helpers for the sm_100a spectral kernels.  A cell is one double2
// (re, im interleaved, 16 B) exactly like Go's complex128 / FFTW's fftw_complex,
// so host slices map 1:1 onto device arrays and every global access is a
// 128-bit LDG/STG.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#else
typedef unsigned int uint32_t;  // NVRTC has no <stdint.h>
#endif

namespace gopf {

typedef double2 cplx;

__host__ __device__ __forceinline__ cplx mk(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return mk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cplx operator*(cplx a, double s) { return mk(a.x * s, a.y * s); }
__device__ __forceinline__ cplx operator*(double s, cplx a) { return mk(a.x * s, a.y * s); }
__device__ __forceinline__ cplx& operator+=(cplx& a, cplx b) { a.x += b.x; a.y += b.y; return a; }
__device__ __forceinline__ cplx& operator-=(cplx& a, cplx b) { a.x -= b.x; a.y -= b.y; return a; }
__device__ __forceinline__ cplx cswap(cplx a) { return mk(a.y, a.x); }
__device__ __forceinline__ cplx cconj(cplx a) { return mk(a.x, -a.y); }
// multiply by -i / +i
__device__ __forceinline__ cplx mul_mi(cplx a) { return mk(a.y, -a.x); }
__device__ __forceinline__ cplx mul_pi(cplx a) { return mk(-a.y, a.x); }

// Go's complex128 '/' (runtime.complex128div) is Smith's algorithm (CACM 5(8):435,
// 1962) plus inf/nan fix-ups; this is the finite-operand part, same operation order.
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x;
        double d = b.x + r * b.y;
        return mk((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y;
        double d = b.y + r * b.x;
        return mk((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}

// 128-bit global accessors.  ld_stream / st_stream carry an evict-first hint for
// data touched once per pass (the field arrays never fit L2 at 256^3 and above).
__device__ __forceinline__ cplx ld_g(const cplx* p) { return *p; }
__device__ __forceinline__ void st_g(cplx* p, cplx v) { *p = v; }
__device__ __forceinline__ cplx ld_tab(const cplx* p) { return __ldg(p); }

}  // namespace gopf
