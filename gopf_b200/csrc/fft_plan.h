// FftPlan: device-side counterpart of pfutil.FFTWWrapper
// (/root/reference/pfutil/fftWrap.go:8-95).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <vector>

#include "cplx.cuh"
#include "fft_kernels.cuh"
#include "host_util.h"

namespace gopf {

class FftPlan {
public:
    FftPlan(int rank, const int* n, int device);
    ~FftPlan();
    FftPlan(const FftPlan&) = delete;
    FftPlan& operator=(const FftPlan&) = delete;

    int rank;
    int dims[3];     // as given by the caller (reference Dimensions)
    int n0, n1, n2;  // normalised to 3 axes, FFTW order (leading extents 1 for rank < 3)
    size_t N;
    int device;
    cudaStream_t stream;
    int tx_want;  // preferred strided-tile width (cells along the fastest axis)

    FreqGeom freq_geom() const;
    // true when Freq components can be read off per-axis tables (rank 2, cubic rank 3)
    bool freq_axis_consistent() const;
    // FFTW axis (0..2 in normalised order) carrying reference Freq component c
    int axis_of_component(int c) const;
    // device table f[idx] = wrap(idx / n[axis]) for normalised axis 0..2
    const double* freq_axis(int axis) const { return d_freq_[axis]; }
    const cplx* twiddle(int axis) const;
    int extent(int axis) const { return axis == 0 ? n0 : (axis == 1 ? n1 : n2); }
    bool axis_fast(int axis) const;
    PassGeom geom(int axis) const { return make_geom(n0, n1, n2, axis); }

    // all axes, in place, unnormalised; sign -1 forward / +1 inverse
    void exec_device(cplx* data, int sign, cudaStream_t s);
    void exec_host(double* host_c128, int sign);
    void freq_device(const long long* nodes, long long count, double* out);
    cplx* scratch();
    void use_device() const;

private:
    std::map<int, cplx*> d_tw_;  // per distinct length
    double* d_freq_[3];
    cplx* d_scratch_;
    cplx* d_buf_;
    void generic_pass(cplx* data, int axis, int sign, cudaStream_t s);
};

}  // namespace gopf
