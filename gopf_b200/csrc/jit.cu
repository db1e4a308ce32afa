// Run-time specialisation of registered functions: see jit.h.
#include "jit.h"

#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "host_util.h"
#include "jit_embed.inc"  // embed_cplx_cuh, embed_step_program_h, embed_kupdate_cuh (Makefile)

namespace gopf {
namespace jit {

// ---- code generation ---------------------------------------------------------------------
namespace {

// a double as a C literal that reads back bit-identically (hexadecimal floating constant,
// C99 6.4.4.2 / C++17)
std::string literal(double v) {
    if (std::isnan(v)) return "(0.0 / 0.0)";
    if (std::isinf(v)) return v > 0 ? "(1.0 / 0.0)" : "(-1.0 / 0.0)";
    char buf[64];
    snprintf(buf, sizeof(buf), "%a", v);
    return std::signbit(v) ? "(" + std::string(buf) + ")" : std::string(buf);
}

std::string tmp(int id) { return "t" + std::to_string(id); }

// helpers shared by every generated function; the forms are the ones eval_derived uses
// (step_program.h), so both paths round alike up to the compiler's contraction of a*b+c
const char* kPrelude =
    "#ifndef GOPF_FN\n"
    "#define GOPF_FN static inline\n"
    "#endif\n"
    "GOPF_FN double gopf_ipow(double x, int n) {\n"
    "    double r = 1.0;\n"
    "    for (int i = 0; i < n; ++i) r *= x;\n"
    "    return r;\n"
    "}\n";

}  // namespace

std::string expression_source(const DevDerived& D, unsigned* used_mask, unsigned* imag_mask) {
    if (D.kind != DK_RPN) throw Error("jit: not a registered-function program");
    unsigned mask = 0, imask = 0;
    std::string body;
    std::vector<int> st;
    int next = 0;
    auto push = [&](const std::string& rhs) {
        body += "    const double " + tmp(next) + " = " + rhs + ";\n";
        st.push_back(next++);
    };
    auto need = [&](size_t n) {
        if (st.size() < n) throw Error("jit: operand stack underflow in the function program");
    };
    auto unary = [&](const std::string& pre, const std::string& post) {
        need(1);
        const int a = st.back();
        st.pop_back();
        push(pre + tmp(a) + post);
    };
    auto binary = [&](const std::string& op) {
        need(2);
        const int b = st.back();
        st.pop_back();
        const int a = st.back();
        st.pop_back();
        push(tmp(a) + " " + op + " " + tmp(b));
    };
    for (int i = 0; i < D.n_ops; ++i) {
        const double a = D.arg[i];
        switch (D.op[i]) {
            case OP_CONST: push(literal(a)); break;
            case OP_FIELD_RE:
            case OP_FIELD_IM: {
                const int f = (int)a;
                if (f < 0 || f >= GOPF_MAX_FIELDS) throw Error("jit: field index out of range");
                mask |= 1u << f;
                if (D.op[i] == OP_FIELD_IM) imask |= 1u << f;
                push(std::string(D.op[i] == OP_FIELD_RE ? "r" : "i") + std::to_string(f));
                break;
            }
            case OP_ADD: binary("+"); break;
            case OP_SUB: binary("-"); break;
            case OP_MUL: binary("*"); break;
            case OP_DIV: binary("/"); break;
            case OP_NEG: unary("-", ""); break;
            case OP_POWI: unary("gopf_ipow(", ", " + std::to_string((int)a) + ")"); break;
            case OP_POW: {
                need(2);
                const int y = st.back();
                st.pop_back();
                const int x = st.back();
                st.pop_back();
                push("pow(" + tmp(x) + ", " + tmp(y) + ")");
                break;
            }
            case OP_H: {
                need(1);
                const std::string x = tmp(st.back());
                st.pop_back();
                push("3.0 * " + x + " * " + x + " - 2.0 * " + x + " * " + x + " * " + x);
                break;
            }
            case OP_DH: {
                need(1);
                const std::string x = tmp(st.back());
                st.pop_back();
                push("6.0 * " + x + " - 6.0 * " + x + " * " + x);
                break;
            }
            case OP_LANDAU: {
                need(1);
                const std::string x = tmp(st.back());
                st.pop_back();
                push(x + " * " + x + " - 2.0 * " + x + " * " + x + " * " + x + " + " + x + " * " + x + " * " + x + " * " + x);
                break;
            }
            case OP_DLANDAU: {
                need(1);
                const std::string x = tmp(st.back());
                st.pop_back();
                push("2.0 * " + x + " - 6.0 * " + x + " * " + x + " + 4.0 * " + x + " * " + x + " * " + x);
                break;
            }
            case OP_EXP: unary("exp(", ")"); break;
            case OP_LOG: unary("log(", ")"); break;
            case OP_SIN: unary("sin(", ")"); break;
            case OP_COS: unary("cos(", ")"); break;
            case OP_TANH: unary("tanh(", ")"); break;
            case OP_SQRT: unary("sqrt(", ")"); break;
            case OP_ABS: unary("fabs(", ")"); break;
            case OP_NEGPART: unary("fmin(", ", 0.0)"); break;
            default: throw Error(strf("jit: unknown op %d in the function program", (int)D.op[i]));
        }
    }
    std::string src = kPrelude;
    src += "GOPF_FN double gopf_expr(";
    for (int f = 0; f < GOPF_MAX_FIELDS; ++f)
        src += strf("%sdouble r%d, double i%d", f ? ", " : "", f, f);
    src += ") {\n";
    for (int f = 0; f < GOPF_MAX_FIELDS; ++f) src += strf("    (void)r%d; (void)i%d;\n", f, f);
    src += body;
    // eval_derived returns the top of the stack, 0 for an empty program
    src += "    return " + (st.empty() ? std::string("0.0") : tmp(st.back())) + ";\n}\n";
    if (used_mask) *used_mask = mask;
    if (imag_mask) *imag_mask = imask;
    return src;
}

std::string derived_pass_source(const DevDerived& D, int N, std::string* name_expr) {
    unsigned mask = 0, imask = 0;
    std::string src = "#define GOPF_FN static __device__ __forceinline__\n" + expression_source(D, &mask, &imask);
    src +=
        "namespace gopf {\n"
        "struct PassIO;\n"
        "template <int E, class At>\n"
        "__device__ __forceinline__ void gopf_jit_load_line(const PassIO& io, double2 (&v)[E], At at);\n"
        "}\n"
        "#define GOPF_JIT_LOAD_LINE(io, v, at) gopf_jit_load_line(io, v, at);\n"
        "#include \"fft_kernels.cuh\"\n"
        "namespace gopf {\n"
        // batches of at most 8 cells bound the registers the loads hold while they are in flight
        "template <int E, class At>\n"
        "__device__ __forceinline__ void gopf_jit_load_line(const PassIO& io, double2 (&v)[E], At at) {\n"
        "    constexpr int H = E > 8 ? 8 : E;\n"
        "#pragma unroll\n"
        "    for (int m0 = 0; m0 < E; m0 += H) {\n";
    std::string args;
    for (int f = 0; f < GOPF_MAX_FIELDS; ++f) {
        const bool used = (mask >> f) & 1u, im = (imask >> f) & 1u;
        if (used && im) {
            src += strf("        double2 c%d[H];\n#pragma unroll\n        for (int j = 0; j < H; ++j) c%d[j] = io.R.r[%d][at(m0 + j)];\n", f, f, f);
            args += strf("%sc%d[j].x, c%d[j].y", f ? ", " : "", f, f);
        } else if (used) {
            src += strf("        const double* p%d = reinterpret_cast<const double*>(io.R.r[%d]);\n        double r%d[H];\n", f, f, f);
            src += strf("#pragma unroll\n        for (int j = 0; j < H; ++j) r%d[j] = p%d[2 * at(m0 + j)];\n", f, f);
            args += strf("%sr%d[j], 0.0", f ? ", " : "", f);
        } else {
            args += strf("%s0.0, 0.0", f ? ", " : "");
        }
    }
    src += "#pragma unroll\n        for (int j = 0; j < H; ++j) v[m0 + j] = mk(gopf_expr(" + args + "), 0.0);\n";
    src += "    }\n}\n";
    src += strf("template __global__ void k_pass_contig<%d>(const PassGeom g, const __grid_constant__ PassIO io, const cplx* __restrict__ tw);\n", N);
    src += "}  // namespace gopf\n";
    if (name_expr) *name_expr = strf("gopf::k_pass_contig<%d>", N);
    return src;
}

std::string derived_kernel_source(const DevDerived& D, unsigned* used_mask) {
    unsigned mask = 0;
    std::string src = "#define GOPF_FN static __device__ __forceinline__\n" + expression_source(D, &mask);
    // two cells per thread and iteration: 2048 resident threads x 2 x 16 B keeps ~64 KB of loads
    // per SM in flight (the bandwidth-delay product of HBM3e is ~35 KB per SM)
    src +=
        "extern \"C\" __global__ void __launch_bounds__(256)\n"
        "    gopf_jit_derived(const double2* f0, const double2* f1, const double2* f2, const double2* f3,\n"
        "                     double2* out, long long n) {\n"
        "    (void)f0; (void)f1; (void)f2; (void)f3;\n"
        "    const long long stride = (long long)gridDim.x * blockDim.x;\n"
        "    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2 * stride) {\n"
        "        const long long j = i + stride;\n"
        "        const bool two = j < n;\n"
        "        const long long jj = two ? j : i;\n";
    std::string args_a, args_b;
    for (int f = 0; f < GOPF_MAX_FIELDS; ++f) {
        const bool used = (mask >> f) & 1u;
        if (used) {
            src += strf("        const double2 a%d = f%d[i];\n", f, f);
            src += strf("        const double2 b%d = f%d[jj];\n", f, f);
        }
        args_a += used ? strf("%sa%d.x, a%d.y", f ? ", " : "", f, f) : strf("%s0.0, 0.0", f ? ", " : "");
        args_b += used ? strf("%sb%d.x, b%d.y", f ? ", " : "", f, f) : strf("%s0.0, 0.0", f ? ", " : "");
    }
    src += "        const double va = gopf_expr(" + args_a + ");\n";
    src += "        const double vb = gopf_expr(" + args_b + ");\n";
    src +=
        "        out[i] = make_double2(va, 0.0);\n"
        "        if (two) out[j] = make_double2(vb, 0.0);\n"
        "    }\n"
        "}\n";
    if (used_mask) *used_mask = mask;
    return src;
}

std::string kupdate_kernel_source(const DevKProgram& P_in, const FreqGeom& fg, long long n, unsigned tab_mask) {
    static_assert(sizeof(DevKProgram) % 8 == 0, "DevKProgram is laid down as 64-bit words");
    DevKProgram P;
    std::memcpy(&P, &P_in, sizeof(P));
    // a literal NULL behind a dereference is undefined behaviour the compiler may turn into an empty kernel
    for (int i = 0; i < P.n_fields; ++i)
        for (int side = 0; side < 2; ++side) {
            const DevTerm* t = side ? P.eq[i].den : P.eq[i].rhs;
            const int nt = side ? P.eq[i].n_den : P.eq[i].n_rhs;
            for (int j = 0; j < nt; ++j)
                if (t[j].kind == TK_VOLUME_LP && (t[j].param < 0 || t[j].param >= GOPF_MAX_SPECIAL || !P.lp_multiplier[t[j].param]))
                    throw Error("jit: VolumeConservingLP term without a multiplier address");
        }
    std::string src;
    src += strf("#define GOPF_FILTER(P) ((const double*)0x%llxull)\n", (unsigned long long)(uintptr_t)P.filter);
    src += "#define GOPF_LP(P, slot) (";
    for (int i = 0; i < GOPF_MAX_SPECIAL; ++i)
        src += strf("(slot) == %d ? (const double*)0x%llxull : ", i, (unsigned long long)(uintptr_t)P.lp_multiplier[i]);
    src += "(const double*)0)\n";
    src += strf("#define GOPF_TAB_PRESENT(tab, i) (((%uu) >> (i)) & 1u)\n", tab_mask);
    src += "#define GOPF_JIT_UNROLL _Pragma(\"unroll\")\n";
    src += "#include \"kupdate.cuh\"\n";
    // the addresses live in the macros; the constant image carries none
    P.filter = nullptr;
    for (int i = 0; i < GOPF_MAX_SPECIAL; ++i) P.lp_multiplier[i] = nullptr;
    const size_t words = sizeof(DevKProgram) / 8;
    unsigned long long w[sizeof(DevKProgram) / 8];
    std::memcpy(w, &P, sizeof(P));
    src += strf("namespace gopf {\nstatic __device__ const unsigned long long jit_prog_words[%zu] = {", words);
    for (size_t i = 0; i < words; ++i) src += strf("%s0x%llxull,", i % 6 == 0 ? "\n    " : " ", w[i]);
    src += "\n};\n}  // namespace gopf\n";
    src += strf(
        "#define GOPF_JIT_PROGRAM (*reinterpret_cast<const gopf::DevKProgram*>(gopf::jit_prog_words))\n"
        "#define GOPF_JIT_GEOM {%d, %d, %d, %d}\n"
        "#define GOPF_JIT_NODES %lldLL\n",
        fg.rank, fg.d0, fg.d1, fg.d2, n);
    src +=
        "extern \"C\" __global__ void __launch_bounds__(256) gopf_jit_kupdate(gopf::SpectraPtrs sp, gopf::ImplicitTab tab) {\n"
        "    const gopf::FreqGeom fg = GOPF_JIT_GEOM;\n"
        "    gopf::update_all(GOPF_JIT_PROGRAM, sp, tab, fg, GOPF_JIT_NODES);\n"
        "}\n"
        // the RK4 passes of the same program (pf/rk4.go:29-127)
        "extern \"C\" __global__ void __launch_bounds__(256) gopf_jit_rk4_rhs(gopf::SpectraPtrs sp, gopf::SpectraPtrs kout) {\n"
        "    const gopf::FreqGeom fg = GOPF_JIT_GEOM;\n"
        "    gopf::rk4_rhs_all(GOPF_JIT_PROGRAM, sp, kout, fg, GOPF_JIT_NODES);\n"
        "}\n"
        "extern \"C\" __global__ void __launch_bounds__(256) gopf_jit_rk4_point(int mode, double fdt, gopf::SpectraPtrs field,\n"
        "        gopf::SpectraPtrs initial, gopf::SpectraPtrs final_, gopf::SpectraPtrs kf) {\n"
        "    const gopf::FreqGeom fg = GOPF_JIT_GEOM;\n"
        "    gopf::rk4_point_all(GOPF_JIT_PROGRAM, mode, fdt, field, initial, final_, kf, fg, GOPF_JIT_NODES);\n"
        "}\n";
    return src;
}

// ---- NVRTC ---------------------------------------------------------------------------------
namespace {

struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    int (*AddNameExpression)(void*, const char*) = nullptr;
    int (*GetLoweredName)(void*, const char*, const char**) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};

template <class F>
bool sym(void* h, const char* name, F* out, std::string* why) {
    *out = reinterpret_cast<F>(dlsym(h, name));
    if (!*out) {
        *why = std::string("symbol ") + name + " not found";
        return false;
    }
    return true;
}

Nvrtc* nvrtc() {
    static Nvrtc lib;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                               "/usr/local/cuda/targets/x86_64-linux/lib/libnvrtc.so.12"};
        for (const char* n : names) {
            lib.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (lib.h) break;
        }
        if (!lib.h) {
            lib.why = "libnvrtc.so.12 not found";
            return;
        }
        const bool ok = sym(lib.h, "nvrtcCreateProgram", &lib.CreateProgram, &lib.why) &&
                        sym(lib.h, "nvrtcCompileProgram", &lib.CompileProgram, &lib.why) &&
                        sym(lib.h, "nvrtcGetCUBINSize", &lib.GetCUBINSize, &lib.why) &&
                        sym(lib.h, "nvrtcGetCUBIN", &lib.GetCUBIN, &lib.why) &&
                        sym(lib.h, "nvrtcGetProgramLogSize", &lib.GetProgramLogSize, &lib.why) &&
                        sym(lib.h, "nvrtcGetProgramLog", &lib.GetProgramLog, &lib.why) &&
                        sym(lib.h, "nvrtcDestroyProgram", &lib.DestroyProgram, &lib.why) &&
                        sym(lib.h, "nvrtcAddNameExpression", &lib.AddNameExpression, &lib.why) &&
                        sym(lib.h, "nvrtcGetLoweredName", &lib.GetLoweredName, &lib.why) &&
                        sym(lib.h, "nvrtcGetErrorString", &lib.GetErrorString, &lib.why);
        if (!ok) lib.h = nullptr;
    });
    return &lib;
}

}  // namespace

bool compile_cubin(const std::string& source, std::vector<char>* cubin, std::string* log, const std::string* name_expr,
                   std::string* lowered) {
    Nvrtc* rt = nvrtc();
    if (!rt->h) {
        if (log) *log = "NVRTC unavailable: " + rt->why;
        return false;
    }
    void* prog = nullptr;
    const char* headers[] = {embed_cplx_cuh, embed_step_program_h, embed_kupdate_cuh, embed_fft_engine_cuh, embed_fft_kernels_cuh};
    const char* names[] = {"cplx.cuh", "step_program.h", "kupdate.cuh", "fft_engine.cuh", "fft_kernels.cuh"};
    int rc = rt->CreateProgram(&prog, source.c_str(), "gopf_jit.cu", 5, headers, names);
    if (rc != 0) {
        if (log) *log = std::string("nvrtcCreateProgram: ") + rt->GetErrorString(rc);
        return false;
    }
    if (name_expr && (rc = rt->AddNameExpression(prog, name_expr->c_str())) != 0) {
        if (log) *log = std::string("nvrtcAddNameExpression: ") + rt->GetErrorString(rc);
        rt->DestroyProgram(&prog);
        return false;
    }
    // -default-device: the host-side inline helpers of the embedded headers are never called here
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"};
    rc = rt->CompileProgram(prog, 4, opts);
    std::string text;
    size_t log_size = 0;
    if (rt->GetProgramLogSize(prog, &log_size) == 0 && log_size > 1) {
        text.resize(log_size);
        rt->GetProgramLog(prog, &text[0]);
    }
    bool ok = rc == 0;
    if (!ok) text = std::string("nvrtcCompileProgram: ") + rt->GetErrorString(rc) + "\n" + text;
    if (ok) {
        size_t sz = 0;
        ok = rt->GetCUBINSize(prog, &sz) == 0 && sz > 0;
        if (ok) {
            cubin->resize(sz);
            ok = rt->GetCUBIN(prog, cubin->data()) == 0;
        }
        if (!ok) text += "\nnvrtcGetCUBIN failed";
        if (ok && name_expr && lowered) {
            const char* low = nullptr;
            ok = rt->GetLoweredName(prog, name_expr->c_str(), &low) == 0 && low;
            if (ok) *lowered = low;
            else text += "\nnvrtcGetLoweredName failed";
        }
    }
    rt->DestroyProgram(&prog);
    if (log) *log = text;
    if (const char* dir = std::getenv("GOPF_JIT_DUMP")) {  // sources and images for cuobjdump / ncu --import-source
        static int serial = 0;
        const std::string stem = std::string(dir) + "/gopf_jit_" + std::to_string(serial++);
        if (FILE* f = fopen((stem + ".cu").c_str(), "w")) {
            fwrite(source.data(), 1, source.size(), f);
            fclose(f);
        }
        if (ok)
            if (FILE* f = fopen((stem + ".cubin").c_str(), "wb")) {
                fwrite(cubin->data(), 1, cubin->size(), f);
                fclose(f);
            }
    }
    return ok;
}

// ---- driver API: load and launch -------------------------------------------------------------
namespace {

struct Driver {
    void* h = nullptr;
    int (*ModuleLoadData)(void**, const void*) = nullptr;
    int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
    int (*ModuleUnload)(void*) = nullptr;
    int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**,
                        void**) = nullptr;
    int (*GetErrorString)(int, const char**) = nullptr;
    int (*FuncSetAttribute)(void*, int, int) = nullptr;
    std::string why;
};

Driver* driver() {
    static Driver lib;
    static std::once_flag once;
    std::call_once(once, [] {
        lib.h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!lib.h) lib.h = dlopen("libcuda.so", RTLD_NOW | RTLD_LOCAL);
        if (!lib.h) {
            lib.why = "libcuda.so.1 not found";
            return;
        }
        const bool ok = sym(lib.h, "cuModuleLoadData", &lib.ModuleLoadData, &lib.why) &&
                        sym(lib.h, "cuModuleGetFunction", &lib.ModuleGetFunction, &lib.why) &&
                        sym(lib.h, "cuModuleUnload", &lib.ModuleUnload, &lib.why) &&
                        sym(lib.h, "cuLaunchKernel", &lib.LaunchKernel, &lib.why) &&
                        sym(lib.h, "cuGetErrorString", &lib.GetErrorString, &lib.why) &&
                        sym(lib.h, "cuFuncSetAttribute", &lib.FuncSetAttribute, &lib.why);
        if (!ok) lib.h = nullptr;
    });
    return &lib;
}

std::string driver_error(Driver* d, const char* what, int rc) {
    const char* s = nullptr;
    if (d->GetErrorString) d->GetErrorString(rc, &s);
    return strf("%s: %s (CUresult %d)", what, s ? s : "unknown error", rc);
}

}  // namespace

struct Kernel {
    void* module = nullptr;
    void* function = nullptr;
    size_t smem_opted = 0;  // dynamic shared memory the function has been opted in for
};

Kernel* load(const std::vector<char>& cubin, const char* entry, std::string* log) {
    Driver* d = driver();
    if (!d->h) {
        if (log) *log = "CUDA driver unavailable: " + d->why;
        return nullptr;
    }
    // the runtime creates the primary context lazily; make sure it exists and is current here
    if (cudaFree(nullptr) != cudaSuccess) {
        if (log) *log = std::string("no CUDA context: ") + cudaGetErrorString(cudaGetLastError());
        return nullptr;
    }
    Kernel* k = new Kernel();
    int rc = d->ModuleLoadData(&k->module, cubin.data());
    if (rc != 0) {
        if (log) *log = driver_error(d, "cuModuleLoadData", rc);
        delete k;
        return nullptr;
    }
    rc = d->ModuleGetFunction(&k->function, k->module, entry);
    if (rc != 0) {
        if (log) *log = driver_error(d, "cuModuleGetFunction", rc);
        d->ModuleUnload(k->module);
        delete k;
        return nullptr;
    }
    return k;
}

void unload(Kernel* k) {
    if (!k) return;
    Driver* d = driver();
    if (d->h && k->module) d->ModuleUnload(k->module);
    delete k;
}

bool launch(Kernel* k, unsigned grid, unsigned block, void** args, cudaStream_t stream, std::string* log, size_t smem) {
    Driver* d = driver();
    if (!k || !d->h) {
        if (log) *log = "jit kernel not loaded";
        return false;
    }
    if (smem > 48 * 1024 && smem > k->smem_opted) {
        const int CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8;
        const int rc = d->FuncSetAttribute(k->function, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
        if (rc != 0) {
            if (log) *log = driver_error(d, "cuFuncSetAttribute", rc);
            return false;
        }
        k->smem_opted = smem;
    }
    const int rc = d->LaunchKernel(k->function, grid, 1, 1, block, 1, 1, (unsigned)smem, (void*)stream, args, nullptr);
    if (rc != 0) {
        if (log) *log = driver_error(d, "cuLaunchKernel", rc);
        return false;
    }
    return true;
}

// On by default since round 2 (the whole device suite runs green with it); GOPF_JIT=0 keeps the
// interpreter kernels.
bool enabled() {
    const char* e = std::getenv("GOPF_JIT");
    return !(e && e[0] == '0');
}

// On by default since its first device run (round 2: tests/test_zz_jit_gpu.py green, cfg 4 at 512^3
// 25.3 -> 23.1 ms/step); GOPF_JIT_INPASS=0 keeps the pointwise kernel + plain pass.
bool inpass_enabled() {
    const char* e = std::getenv("GOPF_JIT_INPASS");
    return !(e && e[0] == '0');
}

}  // namespace jit
}  // namespace gopf
