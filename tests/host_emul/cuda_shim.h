// Test infrastructure: just enough of the CUDA C++ surface to compile the header-only device
// evaluators of gopf_b200/csrc (cplx.cuh, step_program.h, kupdate.cuh) with g++ and run them on the
// host, one "thread" in a 1 x 1 grid.  Nothing under gopf_b200/ includes this; the product has no
// CPU path.  Used by tests/test_host_emulation_cpu.py.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE  // sincos
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __CUDACC_RTC__ 1  // the headers then skip <cuda_runtime.h>
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__

struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) {
    double2 r;
    r.x = x;
    r.y = y;
    return r;
}
struct emul_dim3 {
    unsigned x, y, z;
};
static const emul_dim3 gridDim = {1, 1, 1}, blockDim = {1, 1, 1}, blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0};

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
template <class T>
static inline T __ldg(const T* p) {
    return *p;
}
static inline void sincospi(double x, double* s, double* c);
// cos(pi x) with the argument reduced exactly first, like the device function
static inline double cospi(double x) {
    double r = fmod(fabs(x), 2.0);  // exact
    if (r > 1.0) r = 2.0 - r;       // cos is even and 2-periodic in x
    if (r == 0.5) return 0.0;
    return r < 0.5 ? cos(M_PI * r) : -cos(M_PI * (1.0 - r));
}
static inline double sinpi(double x) { return cospi(x - 0.5); }
static inline void sincospi(double x, double* s, double* c) {
    *s = sinpi(x);
    *c = cospi(x);
}
