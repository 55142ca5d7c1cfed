// Shared host-side plumbing for the C-ABI translation units.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "soccdpt_b200.h"

namespace soccdpt {

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SOCCDPT_E_CUDA;
    }
    return SOCCDPT_OK;
}

#define SOCCDPT_REQUIRE(cond, ...)             \
    do {                                       \
        if (!(cond)) {                         \
            ::soccdpt::set_error(__VA_ARGS__); \
            return SOCCDPT_E_INVALID;          \
        }                                      \
    } while (0)

#define SOCCDPT_CUDA(call)                                                        \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            ::soccdpt::set_error("%s failed: %s", #call, cudaGetErrorString(e__)); \
            return SOCCDPT_E_CUDA;                                                \
        }                                                                         \
    } while (0)

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

inline cudaStream_t as_stream(soccdpt_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace soccdpt
