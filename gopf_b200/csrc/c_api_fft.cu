// C ABI, part 1: library info, pfutil index helpers and the transform-level
// FFTWWrapper replacement.  Contract: include/gopf_cuda.h.
#include "tma_launch.h"
#include "../../include/gopf_cuda.h"
#include "fft_plan.h"

using namespace gopf;

struct gopf_fft_plan {
    FftPlan* p;
};

extern "C" {

const char* gopf_last_error(void) { return get_last_error(); }

int gopf_abi_version(void) { return GOPF_ABI_VERSION; }

int gopf_tma_launch_count(int reset, int64_t* launches) {
    GOPF_API_BEGIN
    if (!launches) throw Error("gopf_tma_launch_count: launches is NULL");
    *launches = (int64_t)tma_launch_count(reset != 0);
    GOPF_API_END
}

int gopf_device_count(int* count) {
    GOPF_API_BEGIN
    if (!count) throw Error("gopf_device_count: count is NULL");
    *count = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) throw Error(strf("cudaGetDeviceCount: %s", cudaGetErrorString(e)));
    *count = n;
    GOPF_API_END
}

// pfutil/indexPositionConversion.go:4-22
int gopf_node_idx(int rank, const int* d, const int* pos, int64_t* node) {
    GOPF_API_BEGIN
    if (!d || !pos || !node) throw Error("gopf_node_idx: NULL argument");
    if (rank == 2) *node = (int64_t)pos[0] * d[1] + pos[1];
    else if (rank == 3) *node = (int64_t)pos[2] * d[0] * d[1] + (int64_t)pos[0] * d[1] + pos[1];
    else throw Error("util: Domain size and idx has to be of length 2 or 3");
    GOPF_API_END
}

// pfutil/indexPositionConversion.go:24-44
int gopf_pos(int rank, const int* d, int64_t node, int* out) {
    GOPF_API_BEGIN
    if (!d || !out) throw Error("gopf_pos: NULL argument");
    if (rank == 2) {
        out[1] = (int)(node % d[1]);
        out[0] = (int)(node / d[1]);
    } else if (rank == 3) {
        out[1] = (int)(node % d[1]);
        out[0] = (int)((node / d[1]) % d[0]);
        out[2] = (int)(node / ((int64_t)d[0] * d[1]));
    } else {
        throw Error("util: Domain size has to be either 2 or 3");
    }
    GOPF_API_END
}

// pfutil/fftWrap.go:57-74
int gopf_freq(int rank, const int* n, int64_t i, double* out) {
    GOPF_API_BEGIN
    if (!n || !out) throw Error("gopf_freq: NULL argument");
    if (rank != 2 && rank != 3) throw Error("gopf_freq: rank must be 2 or 3 (fftWrap.go:61 indexes res[1])");
    FreqGeom g;
    g.rank = rank;
    g.d0 = n[0];
    g.d1 = n[1];
    g.d2 = rank > 2 ? n[2] : 1;
    ref_freq(g, i, out);
    GOPF_API_END
}

// pfutil/fftWrap.go:78-95
int gopf_conjugate_node(int rank, const int* n, int64_t i, int64_t* out) {
    GOPF_API_BEGIN
    if (!n || !out) throw Error("gopf_conjugate_node: NULL argument");
    if (rank != 2 && rank != 3) throw Error("gopf_conjugate_node: rank must be 2 or 3");
    const int64_t nr = n[0], nc = n[1];
    const int64_t c = i % nc, r = (i / nc) % nr;
    const int64_t conj_c = (nc - c) % nc, conj_r = (nr - r) % nr;
    int64_t conj_d = 0;
    if (rank == 3) {
        const int64_t d = i / (nr * nc);
        conj_d = (n[2] - d) % n[2];
    }
    *out = conj_d * nr * nc + conj_r * nc + conj_c;
    GOPF_API_END
}

int gopf_fft_plan_create(int rank, const int* n, int device, gopf_fft_plan** out) {
    GOPF_API_BEGIN
    if (!n || !out) throw Error("gopf_fft_plan_create: NULL argument");
    *out = nullptr;
    FftPlan* p = new FftPlan(rank, n, device);
    gopf_fft_plan* h = new gopf_fft_plan;
    h->p = p;
    *out = h;
    GOPF_API_END
}

int gopf_fft_exec(gopf_fft_plan* plan, double* host, int sign) {
    GOPF_API_BEGIN
    if (!plan) throw Error("gopf_fft_exec: plan is NULL");
    plan->p->exec_host(host, sign);
    GOPF_API_END
}

int gopf_fft_exec_device(gopf_fft_plan* plan, void* dev, int sign, void* stream) {
    GOPF_API_BEGIN
    if (!plan || !dev) throw Error("gopf_fft_exec_device: NULL argument");
    plan->p->exec_device(reinterpret_cast<cplx*>(dev), sign, reinterpret_cast<cudaStream_t>(stream));
    GOPF_API_END
}

int gopf_fft_exec_axis_device(gopf_fft_plan* plan, const void* in, void* out, int sign, int axis, int tile_cells,
                              void* stream) {
    GOPF_API_BEGIN
    if (!plan || !in || !out) throw Error("gopf_fft_exec_axis_device: NULL argument");
    FftPlan& p = *plan->p;
    if (axis < 0 || axis > 2 || p.extent(axis) <= 1) throw Error("gopf_fft_exec_axis_device: bad axis");
    if (!p.axis_fast(axis)) throw Error("gopf_fft_exec_axis_device: axis length has no fast kernel");
    if (sign != -1 && sign != 1) throw Error("gopf_fft_exec_axis_device: sign must be -1 or +1");
    p.use_device();
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : p.stream;
    cudaError_t e = launch_pass(p.geom(axis), tile_cells > 0 ? tile_cells : p.tx_want,
                                plain_io(reinterpret_cast<const cplx*>(in), reinterpret_cast<cplx*>(out), sign > 0, 1.0),
                                p.twiddle(axis), s);
    if (e != cudaSuccess) throw Error(strf("axis pass failed: %s", cudaGetErrorString(e)));
    GOPF_API_END
}

int gopf_fft_exec_rows_device(gopf_fft_plan* plan, const void* in, void* out, int sign, int axis, int64_t slabs, int64_t cols,
                              const int64_t* in_map, const int64_t* out_map, int tile_cells, void* stream) {
    GOPF_API_BEGIN
    if (!plan || !in || !out || !in_map || !out_map) throw Error("gopf_fft_exec_rows_device: NULL argument");
    FftPlan& p = *plan->p;
    if (axis < 0 || axis > 2 || p.extent(axis) <= 1) throw Error("gopf_fft_exec_rows_device: bad axis");
    if (!p.axis_fast(axis)) throw Error("gopf_fft_exec_rows_device: axis length has no fast kernel");
    if (sign != -1 && sign != 1) throw Error("gopf_fft_exec_rows_device: sign must be -1 or +1");
    if (slabs < 1 || cols < 2) throw Error("gopf_fft_exec_rows_device: need slabs >= 1 and cols >= 2");
    p.use_device();
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : p.stream;
    PassGeom g = p.geom(axis);
    g.axis = 1;  // a strided pass: the axis number only matters to the k-space kernels
    g.A = slabs;
    g.B = cols;
    g.bcount = g.bw = cols;
    auto to_map = [](const int64_t* m) {
        RowMap r = uniform_rows(m[0], m[1]);
        if (m[3] < 31) {
            r.split_stride = m[2];
            r.split_log = (int)m[3];
            r.split_mask = (1 << r.split_log) - 1;
        }
        if (m[5] < 31) {
            r.a_split_stride = m[4];
            r.a_split_log = (int)m[5];
            r.a_split_mask = (1 << r.a_split_log) - 1;
        }
        return r;
    };
    g.in = to_map(in_map);
    g.out = to_map(out_map);
    cudaError_t e = launch_pass(g, tile_cells > 0 ? tile_cells : p.tx_want,
                                plain_io(reinterpret_cast<const cplx*>(in), reinterpret_cast<cplx*>(out), sign > 0, 1.0),
                                p.twiddle(axis), s);
    if (e != cudaSuccess) throw Error(strf("row pass failed: %s", cudaGetErrorString(e)));
    GOPF_API_END
}

int gopf_fft_freq_device(gopf_fft_plan* plan, const int64_t* nodes, int64_t count, double* out) {
    GOPF_API_BEGIN
    if (!plan || !nodes || !out) throw Error("gopf_fft_freq_device: NULL argument");
    plan->p->freq_device(reinterpret_cast<const long long*>(nodes), count, out);
    GOPF_API_END
}

int gopf_fft_plan_destroy(gopf_fft_plan* plan) {
    GOPF_API_BEGIN
    if (plan) {
        delete plan->p;
        delete plan;
    }
    GOPF_API_END
}

}  // extern "C"
