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
#ifndef GOPF_EMUL_THREADS
// pointwise evaluators: one "thread" in a 1 x 1 grid
static const emul_dim3 gridDim = {1, 1, 1}, blockDim = {1, 1, 1}, blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0};
#else
// kernels with shared memory and barriers (the FFT passes): one OS thread per CUDA thread of a block, blocks
// one after the other; emul_launch below drives them
#include <pthread.h>

#include <thread>
#include <vector>
#define GOPF_HOST_EMUL 1
static thread_local emul_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
static emul_dim3 gridDim = {1, 1, 1}, blockDim = {1, 1, 1};
static pthread_barrier_t emul_block_barrier;
static pthread_barrier_t emul_warp_barrier[32];  // one per warp of the block: __syncwarp orders that warp only
#ifdef GOPF_EMUL_NO_BARRIERS  // self-test of the sanitizer runs: without the barriers the races must be reported
static inline void __syncthreads() {}
static inline void __syncwarp() {}
#else
static inline void __syncthreads() { pthread_barrier_wait(&emul_block_barrier); }
static inline void __syncwarp() { pthread_barrier_wait(&emul_warp_barrier[threadIdx.x / 32]); }
#endif
#define __shared__
#define __align__(n) __attribute__((aligned(n)))
namespace gopf {
alignas(16) unsigned char gopf_smem_raw[232448];  // the kernels' `extern __shared__` array (227 KB)
}
template <class Kernel, class... Args>
static void emul_launch(Kernel kernel, unsigned grid, unsigned block, Args... args) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b) {
        pthread_barrier_init(&emul_block_barrier, nullptr, block);
        const unsigned warps = (block + 31) / 32;
        for (unsigned w = 0; w < warps; ++w) pthread_barrier_init(&emul_warp_barrier[w], nullptr, w + 1 < warps ? 32 : block - 32 * w);
        std::vector<std::thread> threads;
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([=]() {
                threadIdx.x = t;
                blockIdx.x = b;
                kernel(args...);
            });
        for (std::thread& th : threads) th.join();
        pthread_barrier_destroy(&emul_block_barrier);
        for (unsigned w = 0; w < warps; ++w) pthread_barrier_destroy(&emul_warp_barrier[w]);
    }
}
#endif

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
