// Probe: throughput of 128-bit stores from a kernel on GPU 0 into GPU 1's memory over NVLink as
// a function of the contiguous segment a group of lanes writes (the row width of a pass tile).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o peer_store_probe peer_store_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// total cells = rows * seg_cells; segment s lives at dst + s*row_stride_cells; lanes of a warp fill
// segments cell by cell (seg_cells consecutive lanes per segment)
__global__ void k_store(double2* __restrict__ dst, long long n_cells, int seg_cells, long long row_stride, int rows_per_col) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (long long)gridDim.x * blockDim.x) {
        const long long seg = i / seg_cells;
        const int c = (int)(i - seg * seg_cells);
        // segments of one column block are row_stride apart; successive column blocks are adjacent
        const long long col = seg / rows_per_col, row = seg - col * rows_per_col;
        dst[row * row_stride + col * seg_cells + c] = make_double2((double)i, 1.0);
    }
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    const long long n = 1LL << 28;  // 4 GiB of double2
    double2 *local = nullptr, *remote = nullptr;
    CK(cudaMalloc(&local, n * 16));
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&remote, n * 16));
    CK(cudaSetDevice(0));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    const int rows = 1024;
    for (int pass = 0; pass < 2; ++pass) {
        double2* dst = pass == 0 ? local : remote;
        for (int seg = 2; seg <= 64; seg *= 2) {
            const long long row_stride = n / rows;  // column blocks tile each row
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(a));
                k_store<<<148 * 8, 256>>>(dst, n, seg, row_stride, rows);
                CK(cudaEventRecord(b));
                CK(cudaEventSynchronize(b));
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                if (ms < best) best = ms;
            }
            printf("%s segment %4d B: %7.1f GB/s (%.2f ms)\n", pass == 0 ? "local " : "remote", seg * 16, n * 16 / (best * 1e-3) / 1e9, best);
        }
    }
    return 0;
}
