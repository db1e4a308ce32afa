// Host-side error plumbing shared by the C ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace gopf {

void set_last_error(const std::string& msg);
const char* get_last_error();

struct Error : public std::runtime_error {
    explicit Error(const std::string& m) : std::runtime_error(m) {}
};

inline std::string strf(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return std::string(buf);
}

#define GOPF_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            throw ::gopf::Error(::gopf::strf("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                             __FILE__, __LINE__));                          \
    } while (0)

// Wraps a C-ABI body: exceptions become status 1 + last-error text.
#define GOPF_API_BEGIN try {
#define GOPF_API_END                                 \
    return 0;                                        \
    }                                                \
    catch (const std::exception& e) {                \
        ::gopf::set_last_error(e.what());            \
        return 1;                                    \
    }                                                \
    catch (...) {                                    \
        ::gopf::set_last_error("unknown C++ exception"); \
        return 1;                                    \
    }

}  // namespace gopf
