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

// The library serves whichever device is current on the calling thread (one process per GPU is the deployment model, but
// nothing here assumes device 0): device properties and per-kernel attributes are cached PER DEVICE.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < kMaxDevices) ? d : 0;
}
inline int sm_count() {
    static int n[kMaxDevices] = {};
    const int d = current_device();
    if (n[d] == 0) {
        cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d);
        if (n[d] <= 0) n[d] = 148;
    }
    return n[d];
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: remembers, per device, the largest size
// it has been raised to.  `if (attr.need(bytes)) cudaFuncSetAttribute(...)`.
struct SmemAttr {
    size_t v[kMaxDevices] = {};
    bool need(size_t bytes) {
        const int d = current_device();
        if (bytes > v[d]) {
            v[d] = bytes;
            return true;
        }
        return false;
    }
};

inline cudaStream_t as_stream(soccdpt_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL)
// Every kernel of a launch plan depends on the one before it, and a forward pass is ~125 of them: with plain stream order
// kernel i+1 is only SCHEDULED once kernel i has drained, so each boundary costs launch latency + CTA rasterisation + the
// prologue (mbarrier init, TMEM allocation, tensor-map fetch, constant staging).  Kernels launched through launch_pdl()
// carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may become resident while the previous kernel is
// still running; they run their prologue and then block in pdl_wait() (griddepcontrol.wait) until the previous grid has
// completed and its memory is visible.  RULE: a kernel launched through launch_pdl() executes pdl_wait() in every thread
// before its first access to memory another kernel may have written (or may still be reading: its own outputs), and before
// any early return; kernels launched the plain way stay fully serialised, so the two kinds mix freely.
// The attribute is switched per kernel family (bit mask; SOCCDPT_PDL=<mask> in the environment or soccdpt_set_pdl()):
// without it the launch is plain stream order and pdl_wait() is a no-op.
enum PdlFamily { PDL_CONV = 1, PDL_ATTENTION = 2, PDL_ELEMENTWISE = 4, PDL_POSTPROCESS = 8 };
bool pdl_enabled(int family);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int family, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled(family) ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#if defined(__CUDACC__)
// blocks until every grid this one depends on has completed and flushed
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the NEXT kernel's CTAs become resident once every CTA of this grid has called it (or exited: the implicit trigger).
// MEASURED (tools/bench_pdl.py, profiles/r1_progress.md): triggering early -- right after pdl_wait() -- parks the next
// kernel's CTAs on the SMs for the whole duration of this one and costs 1-2 % of throughput at B = 64; so only the
// persistent conv kernel triggers explicitly, when a CTA starts its LAST tile, and everything else leaves it implicit.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

}  // namespace soccdpt
