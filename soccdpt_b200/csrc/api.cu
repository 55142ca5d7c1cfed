// Library-wide entry points: version, error string, launch counter, device query.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace soccdpt {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

static std::atomic<int> g_pdl{-1};
static int pdl_mask() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("SOCCDPT_PDL");
        v = e ? (atoi(e) & 15) : SOCCDPT_PDL_DEFAULT;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v;
}
bool pdl_enabled(int family) { return (pdl_mask() & family) != 0; }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace soccdpt

extern "C" {

int soccdpt_abi_version(void) { return SOCCDPT_ABI_VERSION; }
const char *soccdpt_last_error(void) { return soccdpt::g_err; }
long long soccdpt_launch_count(void) { return soccdpt::g_launches.load(); }

int soccdpt_set_pdl(int mask) {
    const int prev = soccdpt::pdl_mask();
    soccdpt::g_pdl.store(mask < 0 ? -1 : (mask & 15));
    return prev;
}

int soccdpt_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    SOCCDPT_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SOCCDPT_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return SOCCDPT_OK;
}
}
