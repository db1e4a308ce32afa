// Test infrastructure: drives the host-compiled FFT passes (emul_fft.cpp) under a sanitizer.
// drive a few passes under ThreadSanitizer: a missing barrier in the kernels shows up as a race on the shared tile
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
extern "C" int emul_fft_pass(int n0, int n1, int n2, int axis, int inverse, double scale, const double* in, double* out, int tx);
int main() {
    int shapes[][3] = {{1, 6, 64}, {1, 4, 256}, {4, 16, 32}, {8, 64, 8}, {128, 1, 8}, {16, 16, 16}};
    for (auto& sh : shapes) {
        const size_t n = (size_t)sh[0] * sh[1] * sh[2];
        std::vector<double> x(2 * n), y(2 * n);
        for (size_t i = 0; i < 2 * n; ++i) x[i] = (double)rand() / RAND_MAX;
        for (int axis = 0; axis < 3; ++axis) {
            if (sh[axis] <= 1) continue;
            for (int tx : {2, 4, 8}) {
                if (axis == 2 && tx != 2) continue;
                int rc = emul_fft_pass(sh[0], sh[1], sh[2], axis, 0, 1.0, x.data(), y.data(), tx);
                printf("shape %d %d %d axis %d tx %d rc %d\n", sh[0], sh[1], sh[2], axis, tx, rc);
            }
        }
    }
    return 0;
}
