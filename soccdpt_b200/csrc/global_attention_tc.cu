// Global multi-head attention of the ViT-hybrid encoder (timm 0.6.12 vision_transformer.Attention, head dim 64,
// 577 tokens for dpt_hybrid_384) on tcgen05 / TMEM.
//
//   one CTA (256 threads) per (image, head, 128-query tile); keys are processed in blocks of 128:
//     S_j[128x128] = Q K_j^T          tcgen05.mma M128 N128 K64, fp32 accumulators double-buffered in TMEM columns 0..255
//     pass 1: row max over all key blocks (scores are only read, never stored)
//     pass 2: S_j recomputed (4 MMAs, far cheaper than a rescale of O), P_j = exp2((S_j - max) * scale * log2e) -> bf16 pairs
//             written back to TMEM (tcgen05.st, columns 320..383), O[128x64] += P_j V_j with the A operand in TMEM
//             (M128 N64 K128, TMEM columns 256..319): P never touches shared memory
//     out = O / rowsum, bf16
//   Q, K are copied as they are (16-byte chunks) into swizzled rows; V is transposed while staged (the B operand of P V is
//   V^T, K-major).  Two threads per query row (TMEM lane), each owns 64 of the 128 keys of a block and 32 output channels.
//   Padded keys (>= N) have zero K / V rows and are masked out of the max and the exponentials.
#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;
constexpr int GD = 64;                  // head dim
constexpr int GT = 256;                 // threads
constexpr int KB = 128;                 // keys per block
constexpr int MAX_BLOCKS = 6;           // 768 keys: K and V^T (32 KB per 128 keys) next to Q in 227 KB
constexpr int SM_Q = 0;                 // 128 rows x 128 B, SWIZZLE_128B
constexpr int SM_MISC = 16384;          // row max [2][128] | row sum [2][128] | barriers | tmem slot
constexpr int SM_K = SM_MISC + 4096;    // nb x (128 rows x 128 B);  then V^T: 2*nb x (64 rows x 128 B)
constexpr int TM_O = 256;               // TMEM column of O (64 columns)
constexpr int TM_P = 320;               // TMEM columns of P: 128 keys as bf16 pairs = 64 columns (A operand of the P V MMAs)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ uint64_t umma_desc128(uint32_t addr) {   // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t umma_idesc(int n) {   // D=f32, A=B=bf16, K-major, M=128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n"
        "tcgen05.wait::st.sync.aligned;"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31]) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 columns, each 32-bit column holding two consecutive K elements (bf16)
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}
__device__ __forceinline__ uint4 pack8(const float *f) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]);
    u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]);
    u.w = pack_bf16x2(f[6], f[7]);
    return u;
}

__global__ void __launch_bounds__(GT, 1)
global_attention_tc_kernel(const bf16 *__restrict__ qkv, bf16 *__restrict__ out, int N, int heads, float scale_log2e) {
    extern __shared__ uint8_t ga_raw[];
    uint8_t *smem = ga_raw + ((1024u - (smem_u32(ga_raw) & 1023u)) & 1023u);   // offset on the symbol, not a pointer round trip
    float *s_max = reinterpret_cast<float *>(smem + SM_MISC);                  // [2][128]
    float *s_sum = s_max + 256;                                                // [2][128]
    uint64_t *bar_s = reinterpret_cast<uint64_t *>(s_sum + 256);               // [2] S buffers
    uint64_t *bar_pv = bar_s + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_pv + 1);

    const int t = threadIdx.x, warp = t >> 5;
    const int row = t & 127;                       // query row of the tile == TMEM lane
    const int wg = t >> 7;                         // which 64 keys of a block / which 32 output channels
    const int nb = (N + KB - 1) / KB;
    const int b = blockIdx.x / heads, head = blockIdx.x % heads;
    const int C = heads * GD;
    const int q0 = blockIdx.y * 128;
    const bf16 *base = qkv + (long long)b * N * 3 * C + head * GD;
    uint8_t *sK = smem + SM_K, *sVT = sK + nb * 16384;

    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_s[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_s[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_pv)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    soccdpt::pdl_wait();        // qkv is the previous kernel's output
    // ---- stage Q (this tile), K (all keys), V^T (all keys); rows past N are zero
    for (int i = t; i < 128 * 8; i += GT) {
        const int r = i >> 3, c = i & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q0 + r < N) v = *reinterpret_cast<const uint4 *>(base + (long long)(q0 + r) * 3 * C + c * 8);
        *reinterpret_cast<uint4 *>(smem + SM_Q + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    for (int i = t; i < nb * KB * 8; i += GT) {
        const int r = i >> 3, c = i & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < N) v = *reinterpret_cast<const uint4 *>(base + (long long)r * 3 * C + C + c * 8);
        *reinterpret_cast<uint4 *>(sK + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    for (int i = t; i < nb * KB * 8; i += GT) {     // thread = (key, 8 channels): lanes of a warp hold consecutive keys
        const int c = i / (nb * KB), key = i - c * (nb * KB);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (key < N) v = *reinterpret_cast<const uint4 *>(base + (long long)key * 3 * C + 2 * C + c * 8);
        const bf16 *e = reinterpret_cast<const bf16 *>(&v);
        const int kb = key >> 6, col = key & 63;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int d = c * 8 + k;
            *reinterpret_cast<bf16 *>(sVT + kb * 8192 + d * 128 + (((col >> 3) ^ (d & 7)) << 4) + (col & 7) * 2) = e[k];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t idesc_s = umma_idesc(KB), idesc_o = umma_idesc(GD);
    const uint64_t dq = umma_desc128(smem_u32(smem + SM_Q));
    uint32_t ph_s[2] = {0, 0}, ph_pv = 0;

    auto issue_s = [&](int j) {     // S_j = Q K_j^T into buffer j & 1 (thread 0 only)
        const uint64_t dk = umma_desc128(smem_u32(sK + j * 16384));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + (uint32_t)((j & 1) * KB), dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit(&bar_s[j & 1]);
    };

    // ---- pass 1: row max
    if (t == 0) {
        issue_s(0);
        if (nb > 1) issue_s(1);
    }
    float m = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
        mbar_wait(&bar_s[j & 1], ph_s[j & 1]);
        ph_s[j & 1] ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int key0 = j * KB + wg * 64;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            float v[32];
            tmem_ld32(t_row + (uint32_t)((j & 1) * KB + wg * 64 + hh * 32), v);
            if (key0 + hh * 32 + 32 <= N) {
#pragma unroll
                for (int i = 0; i < 32; ++i) m = fmaxf(m, v[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) m = (key0 + hh * 32 + i < N) ? fmaxf(m, v[i]) : m;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                       // buffer j & 1 has been read by everybody
        if (t == 0 && j + 2 < nb) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_s(j + 2);
        }
    }
    s_max[wg * 128 + row] = m;
    __syncthreads();
    m = fmaxf(m, s_max[(wg ^ 1) * 128 + row]);
    const float ml = m * scale_log2e;

    // ---- pass 2: P = exp2(S * c - max * c), O += P V
    if (t == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_s(0);
        if (nb > 1) issue_s(1);
    }
    float l = 0.f;
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
        mbar_wait(&bar_s[j & 1], ph_s[j & 1]);
        ph_s[j & 1] ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int key0 = j * KB + wg * 64;
        uint32_t pk[32];                                       // my 64 probabilities as bf16 pairs (key 2c in the low half)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            float v[32];
            tmem_ld32(t_row + (uint32_t)((j & 1) * KB + wg * 64 + hh * 32), v);
            const bool full = key0 + hh * 32 + 32 <= N;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float p = fast_exp2(fmaf(v[i], scale_log2e, -ml));
                if (!full && key0 + hh * 32 + i >= N) p = 0.f;
                v[i] = p;
                l += p;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[hh * 16 + i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        }
        if (j > 0) {                                           // P V of the previous block must have consumed the P columns
            mbar_wait(bar_pv, ph_pv);
            ph_pv ^= 1;
        }
        tmem_st32(t_row + (uint32_t)(TM_P + wg * 32), pk);     // P never touches shared memory: it is the TMEM A operand of P V
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                       // S_j read, P_j written by everybody
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int kb2 = 0; kb2 < 2; ++kb2) {
                const uint64_t dv = umma_desc128(smem_u32(sVT + (j * 2 + kb2) * 8192));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16_ts(tmem + TM_O, tmem + (uint32_t)(TM_P + (kb2 * 4 + k) * 8), dv + 2 * k, idesc_o, (j | kb2 | k) ? 1u : 0u);
            }
            umma_commit(bar_pv);
            if (j + 2 < nb) issue_s(j + 2);
        }
    }
    s_sum[wg * 128 + row] = l;
    __syncthreads();
    l += s_sum[(wg ^ 1) * 128 + row];
    mbar_wait(bar_pv, ph_pv);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        float o[32];
        tmem_ld32(t_row + (uint32_t)(TM_O + wg * 32), o);
        const int q = q0 + row;
        if (q < N) {
            const float inv = 1.0f / l;
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= inv;
            bf16 *op = out + ((long long)b * N + q) * C + head * GD + wg * 32;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) *reinterpret_cast<uint4 *>(op + c4 * 8) = pack8(o + c4 * 8);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

}  // namespace

namespace soccdpt {
int global_attention_tc_max_tokens() { return MAX_BLOCKS * KB; }

// qkv bf16 [batch, N, 3*heads*64] -> out bf16 [batch, N, heads*64]; N <= 640
int launch_global_attention_tc(const void *qkv, void *out, int batch, int N, int heads, cudaStream_t st) {
    const int nb = (N + KB - 1) / KB;
    const size_t smem = (size_t)SM_K + (size_t)nb * (16384 + 2 * 8192) + 1024;
    static soccdpt::SmemAttr configured;
    if (configured.need(smem)) {
        SOCCDPT_CUDA(cudaFuncSetAttribute(global_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid((unsigned)(batch * heads), (unsigned)((N + 127) / 128));
    SOCCDPT_CUDA(launch_pdl(PDL_ATTENTION, global_attention_tc_kernel, grid, dim3(GT), smem, st, static_cast<const bf16 *>(qkv),
                            static_cast<bf16 *>(out), N, heads, 1.4426950408889634f / sqrtf((float)GD)));
    return check_launch("global_attention_tc_kernel");
}
}  // namespace soccdpt
