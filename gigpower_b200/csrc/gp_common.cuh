// Shared internals of libgigpower_b200: error reporting, launch accounting,
// small complex helpers.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/gigpower_b200.h"

namespace gp {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define GP_CUDA(expr)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess)                                                        \
            return gp::fail(GP_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,     \
                            cudaGetErrorString(e__));                                  \
    } while (0)

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// numpy.nan_to_num on one double (solution_fbs.py:288): NaN -> 0, +-inf -> +-DBL_MAX
__device__ __forceinline__ double nan_to_num(double x) {
    if (x != x) return 0.0;
    if (isinf(x)) return copysign(1.7976931348623157e308, x);
    return x;
}

}  // namespace gp

void gp_nr3_destroy_impl(void* impl);

struct GpFeeder {
    int device = 0;
    int kind = 0;  // 1 = FBS, 2 = NR3
    void* impl = nullptr;
};
