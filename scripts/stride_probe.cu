// Copy-rate probe for the tile shapes of the strided line passes: a tile is ROWS row segments of TX cells
// (16 B each); the rows of a tile sit at (r >> rlog) * r_hi + (r & rmask) * r_lo, the tiles of a launch at
// (o >> olog) * o_hi + (o & omask) * o_lo + chunk * TX * 16, handed out in order by the block scheduler.
// Measures how the row stride (pages touched per tile) bounds what any line kernel can reach on this layout.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/stride_probe scripts/stride_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct Pat {
    long long r_lo, r_hi, o_lo, o_hi;  // bytes
    int rlog, olog;
};

template <int TX>
__global__ void __launch_bounds__(256) k_copy(const char* __restrict__ in, char* __restrict__ out, Pat pi, Pat po,
                                              int rows, int chunks) {
    const long long tile = blockIdx.x;
    const long long o = tile / chunks, c = tile - o * chunks;
    const long long bi = (o >> pi.olog) * pi.o_hi + (o & ((1LL << pi.olog) - 1)) * pi.o_lo + c * TX * 16;
    const long long bo = (o >> po.olog) * po.o_hi + (o & ((1LL << po.olog) - 1)) * po.o_lo + c * TX * 16;
    const int per_row = TX;  // 16-byte pieces per row
    const int lane = threadIdx.x % per_row, r0 = threadIdx.x / per_row, rstep = 256 / per_row;
    constexpr int MAXE = 1024 * TX / 256;
    double2 v[MAXE];
#pragma unroll
    for (int m = 0; m < MAXE; ++m) {
        const int r = r0 + m * rstep;
        if (r < rows) {
            const long long a = bi + (long long)(r >> pi.rlog) * pi.r_hi + (long long)(r & ((1 << pi.rlog) - 1)) * pi.r_lo;
            v[m] = *reinterpret_cast<const double2*>(in + a + lane * 16);
        }
    }
#pragma unroll
    for (int m = 0; m < MAXE; ++m) {
        const int r = r0 + m * rstep;
        if (r < rows) {
            const long long a = bo + (long long)(r >> po.rlog) * po.r_hi + (long long)(r & ((1 << po.rlog) - 1)) * po.r_lo;
            v[m].x += 1.0;
            *reinterpret_cast<double2*>(out + a + lane * 16) = v[m];
        }
    }
}

static const long long KB = 1024, MB = 1024 * 1024;

template <int TX>
static void run(const char* name, const char* in, char* out, Pat pi, Pat po, long long outer, int inplace) {
    const int rows = 1024, chunks = 1024 / TX;
    const long long tiles = outer * chunks;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_copy<TX><<<(unsigned)tiles, 256>>>(in, inplace ? (char*)in : out, pi, po, rows, chunks);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double bytes = 2.0 * tiles * rows * TX * 16;
    printf("%-58s TX=%d %s  %7.3f ms  %6.0f GB/s  %s\n", name, TX, inplace ? "in-place " : "out-of-pl", best,
           bytes / best * 1e-6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
}

int main(int argc, char** argv) {
    const long long n = 1024;
    const long long bytes = n * n * n * 16;
    char *a, *b;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) {
        printf("alloc failed\n");
        return 1;
    }
    cudaMemset(a, 0, bytes);
    cudaMemset(b, 0, bytes);
    // standard layout [a0][a1][a2]
    const Pat ax1_std = {16 * KB, 0, 16 * MB, 0, 30, 30};   // rows a1, outer a0
    const Pat ax0_std = {16 * MB, 0, 16 * KB, 0, 30, 30};   // rows a0, outer a1
    // layout B1 [a0_hi 32][a1 1024][a0_lo 32][a2]: cell = ((a0_hi*1024 + a1)*32 + a0_lo)*16K
    const Pat ax0_b1 = {16 * KB, 512 * MB, 512 * KB, 0, 5, 30};          // rows a0 (lo 16K, hi 512M), outer a1 (512K)
    const Pat ax1_b1 = {512 * KB, 0, 16 * KB, 512 * MB, 30, 5};          // rows a1 (512K), outer a0 (lo 16K, hi 512M)
    // layout B2 [a0_hi 32][a1_hi 128][a0_lo 32][a1_lo 8][a2]
    const Pat ax0_b2 = {128 * KB, 512 * MB, 16 * KB, 4 * MB, 5, 3};      // rows a0; outer a1 (lo 16K x8, hi 4M)
    const Pat ax1_b2 = {16 * KB, 4 * MB, 128 * KB, 512 * MB, 3, 5};      // rows a1 (lo 16K x8, hi 4M); outer a0
    // layout B3 [a0_hi 32][a1_hi 32][a0_lo 32][a1_lo 32][a2]
    const Pat ax0_b3 = {512 * KB, 512 * MB, 16 * KB, 16 * MB, 5, 5};
    const Pat ax1_b3 = {16 * KB, 16 * MB, 512 * KB, 512 * MB, 5, 5};
    // layout T [a1][a0][a2] (axes 0 and 1 swapped): rows a0 at 16K, outer a1 at 16M
    const Pat ax0_t = {16 * KB, 0, 16 * MB, 0, 30, 30};
    const long long outer = argc > 1 ? atoll(argv[1]) : 1024;

    run<4>("axis1 std -> std (stride 16K)", a, b, ax1_std, ax1_std, outer, 0);
    run<4>("axis1 std -> std (stride 16K)", a, b, ax1_std, ax1_std, outer, 1);
    run<8>("axis1 std -> std (stride 16K)", a, b, ax1_std, ax1_std, outer, 1);
    run<4>("axis0 std -> std (stride 16M)", a, b, ax0_std, ax0_std, outer, 0);
    run<4>("axis0 std -> std (stride 16M)", a, b, ax0_std, ax0_std, outer, 1);
    run<8>("axis0 std -> std (stride 16M)", a, b, ax0_std, ax0_std, outer, 1);
    run<4>("axis0 B1 [a0h][a1][a0l][a2] (32 pages/tile)", a, b, ax0_b1, ax0_b1, outer, 0);
    run<4>("axis0 B1 [a0h][a1][a0l][a2] (32 pages/tile)", a, b, ax0_b1, ax0_b1, outer, 1);
    run<8>("axis0 B1 [a0h][a1][a0l][a2] (32 pages/tile)", a, b, ax0_b1, ax0_b1, outer, 1);
    run<4>("axis1 std -> B1 (out stride 512K, 256 pages)", a, b, ax1_std, ax1_b1, outer, 0);
    run<4>("axis1 B1 -> std", a, b, ax1_b1, ax1_std, outer, 0);
    run<4>("axis0 B2 [a0h][a1h][a0l][a1l 8][a2]", a, b, ax0_b2, ax0_b2, outer, 1);
    run<4>("axis1 std -> B2", a, b, ax1_std, ax1_b2, outer, 0);
    run<4>("axis0 B3 [a0h][a1h][a0l][a1l 32][a2]", a, b, ax0_b3, ax0_b3, outer, 1);
    run<4>("axis1 std -> B3", a, b, ax1_std, ax1_b3, outer, 0);
    run<4>("axis0 on T [a1][a0][a2] (stride 16K)", a, b, ax0_t, ax0_t, outer, 1);
    run<4>("axis1 std -> T (out stride 16M)", a, b, ax1_std, ax0_std, outer, 0);
    return 0;
}
