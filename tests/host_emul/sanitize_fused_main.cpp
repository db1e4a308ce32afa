// Test infrastructure: drives the host-compiled fused kernels (emul_fused.cpp) under a sanitizer.
#include <cstdio>
#include <cstdlib>
#include <vector>
extern "C" int emul_fused_steps(int rank, int edge, const void* program, const void* derived, double* S, int nsteps,
                                unsigned long long step0, int tx, int late);
static std::vector<char> slurp(const char* p) { FILE* f = fopen(p, "rb"); std::vector<char> b(1 << 16); size_t n = fread(b.data(), 1, b.size(), f); b.resize(n); fclose(f); return b; }
// argv[1]: directory holding <tag>.prog / <tag>.der written by scripts/host_sanitizers.py
int main(int argc, char** argv) {
    const char* dir = argc > 1 ? argv[1] : ".";
    struct Case { const char* tag; int rank, edge; } cases[] = {{"ch3", 3, 16}, {"ch2", 2, 32}, {"pfc2", 2, 32}, {"kn3", 3, 16}};
    for (auto& c : cases) {
        char a[512], b[512];
        snprintf(a, 512, "%s/%s.prog", dir, c.tag); snprintf(b, 512, "%s/%s.der", dir, c.tag);
        auto prog = slurp(a), der = slurp(b);
        size_t n = 1; for (int i = 0; i < c.rank; ++i) n *= c.edge;
        std::vector<double> S(2 * n);
        for (auto& v : S) v = 0.01 * ((double)rand() / RAND_MAX - 0.5);
        for (int tx : {2, 4}) for (int late : {0, 1}) {
            int rc = emul_fused_steps(c.rank, c.edge, prog.data(), der.data(), S.data(), 2, 0, tx, late);
            printf("%s tx %d late %d rc %d\n", c.tag, tx, late, rc);
        }
    }
}
