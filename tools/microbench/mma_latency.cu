// Microbenchmark: cost per tcgen05.mma for the shapes of the window-attention pipeline (one CTA per SM, one issuing warp,
// elect-predicated issue).  Prints SM cycles per MMA for dependent chains (same accumulator) and interleaved accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I soccdpt_b200/csrc -I include -o build/mma_latency tools/microbench/mma_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace tc;

constexpr uint32_t DESC_HI64 = 32u | (1u << 14) | (4u << 29);     // SW64, SBO 512
constexpr uint32_t DESC_HI128 = 64u | (1u << 14) | (2u << 29);    // SW128, SBO 1024

// MODE: 0 = TS (A in TMEM), B MN-major SW64 ; 1 = TS, B K-major SW64 ; 2 = SS K-major SW64 both ; 3 = SS SW128 both
// NACC: accumulators used round robin; CE: commit + wait after every CE MMAs (0 = only at the end); fully unrolled groups of 16
template <int MODE, int N, int NACC, int CE>
__global__ void __launch_bounds__(128, 1) bench(int groups, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 65536);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16384; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 1) {
        const uint32_t lo = (smem_u32(smem) >> 4) | (1u << 16);
        const uint32_t idesc = umma_idesc(N) | (MODE == 0 ? (1u << 16) : 0u);
        const uint32_t hi = MODE == 3 ? DESC_HI128 : DESC_HI64;
        uint32_t phase = 0;
        long long best = 1ll << 60;
        for (int rep = 0; rep < 5; ++rep) {
            const long long t0 = clock64();
            for (int gq = 0; gq < groups; ++gq) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t d = tmem + 256 + (uint32_t)((i % NACC) * 64);
                    if (MODE <= 1) umma_ts_lo_elect(d, tmem + (uint32_t)((i & 7) * 8), lo + (uint32_t)((i & 7) * 64), hi, idesc, 1u);
                    else umma_ss_lo_elect(d, lo + (uint32_t)((i & 1) * 2), lo + 1024 + (uint32_t)((i & 1) * 2), hi, idesc, 1u);
                    if (CE > 0 && (i + 1) % CE == 0) {
                        umma_commit_elect(bar);
                        mbar_wait(bar, phase);
                        phase ^= 1;
                    }
                }
            }
            const long long t1 = clock64();
            umma_commit_elect(bar);
            mbar_wait(bar, phase);
            phase ^= 1;
            const long long t2 = clock64();
            if (t2 - t0 < best) {
                best = t2 - t0;
                if (threadIdx.x == 32 && blockIdx.x == 0) {
                    out[0] = t1 - t0;
                    out[1] = t2 - t0;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int MODE, int N, int NACC, int CE>
void run(const char *name, long long *d) {
    long long h[2];
    const int groups = 8;
    cudaFuncSetAttribute(bench<MODE, N, NACC, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    bench<MODE, N, NACC, CE><<<148, 128, 70000>>>(groups, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-58s issue %6.1f cycles/MMA, complete %6.1f cycles/MMA  (%s)\n", name, (double)h[0] / (groups * 16), (double)h[1] / (groups * 16),
           cudaGetErrorString(e));
}

int main() {
    long long *d;
    cudaMalloc(&d, 16);
    run<0, 32, 1, 0>("TS N=32 MN-major B, same D (P V chain)", d);
    run<0, 32, 2, 0>("TS N=32 MN-major B, 2 accumulators", d);
    run<0, 32, 4, 0>("TS N=32 MN-major B, 4 accumulators", d);
    run<1, 32, 1, 0>("TS N=32 K-major B, same D", d);
    run<0, 64, 1, 0>("TS N=64 MN-major B, same D", d);
    run<0, 128, 1, 0>("TS N=128 MN-major B, same D", d);
    run<2, 32, 1, 0>("SS N=32 SW64, same D", d);
    run<2, 64, 1, 0>("SS N=64 SW64, same D", d);
    run<2, 64, 4, 0>("SS N=64 SW64, 4 accumulators", d);
    run<2, 128, 1, 0>("SS N=128 SW64, same D", d);
    run<2, 256, 1, 0>("SS N=256 SW64, same D", d);
    run<3, 256, 1, 0>("SS N=256 SW128, same D", d);
    run<2, 64, 4, 2>("SS N=64 SW64, commit+wait every 2 (S hand-off)", d);
    run<0, 32, 1, 4>("TS N=32 MN-major, commit+wait every 4 (P V hand-off)", d);
    run<0, 32, 1, 8>("TS N=32 MN-major, commit+wait every 8", d);
    run<0, 32, 1, 16>("TS N=32 MN-major, commit+wait every 16", d);
    return 0;
}
