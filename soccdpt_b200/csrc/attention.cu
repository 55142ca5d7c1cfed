// SwinV2 window attention (timm 0.6.12 WindowAttention + the window partition / cyclic shift /
// reverse of SwinTransformerBlock._attn), one CTA per (window, head), one thread per query token.
//
//   q,k,v   : slices of the qkv GEMM output (bias already added), head_dim = 32
//   scores  : normalize(q) . normalize(k)^T * exp(min(logit_scale, ln 100)) + cpb_bias[h] + shift_mask
//   softmax : chunked online softmax in fp32, then P @ V
// The cyclic shift is folded into the token index arithmetic (no roll kernels, no window copies);
// the {0,-100} shift mask is regenerated from region ids exactly as timm builds its attn_mask buffer.
//
// Round-1 implementation: K/V of the window staged in shared memory as fp32, CUDA-core FMAs.
// (The tcgen05 version keeps S/P in TMEM; see DESIGN.md "next".)
#include <cstdlib>

#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;
constexpr int D = 32;   // head dim of every SwinV2 config on the path
constexpr int CH = 8;   // keys per online-softmax step

__device__ __forceinline__ void load_head(const bf16 *p, float f[D]) {
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const uint4 u = *reinterpret_cast<const uint4 *>(p + i * 8);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 t = __bfloat1622float2(h[k]);
            f[i * 8 + 2 * k] = t.x;
            f[i * 8 + 2 * k + 1] = t.y;
        }
    }
}

__device__ __forceinline__ int region_of(int p, int size, int ws, int shift) {
    // timm: slices (0,-ws), (-ws,-shift), (-shift,None) over the SHIFTED image
    return p < size - ws ? 0 : (p < size - shift ? 1 : 2);
}

__global__ void window_attention_kernel(const bf16 *__restrict__ qkv, const float *__restrict__ biasT,
                                        const float *__restrict__ scale, bf16 *__restrict__ out, int Hs, int Ws,
                                        int C, int ws, int shift) {
    extern __shared__ float smem[];
    const int N = ws * ws;
    float *Ks = smem;                 // [N][32]
    float *Vs = smem + (size_t)N * D; // [N][32]
    int *reg = reinterpret_cast<int *>(Vs + (size_t)N * D);  // [N]

    const int nwx = Ws / ws, nwy = Hs / ws;
    const int win = blockIdx.x % (nwx * nwy), b = blockIdx.x / (nwx * nwy);
    const int head = blockIdx.y;
    const int i = threadIdx.x;
    const bool live = i < N;

    float q[D];
    long long tok = 0;
    int my_reg = 0;
    if (live) {
        const int ty = i / ws, tx = i % ws;
        const int ys = (win / nwx) * ws + ty, xs = (win % nwx) * ws + tx;   // position in the shifted image
        const int yo = (ys + shift) % Hs, xo = (xs + shift) % Ws;           // roll(-shift): shifted[y] = x[y+shift]
        tok = ((long long)b * Hs + yo) * Ws + xo;
        const bf16 *base = qkv + tok * 3 * C + head * D;
        float k[D], v[D];
        load_head(base, q);
        load_head(base + C, k);
        load_head(base + 2 * C, v);
        float qq = 0.f, kk = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) { qq = fmaf(q[d], q[d], qq); kk = fmaf(k[d], k[d], kk); }
        const float qs = scale[head] / fmaxf(sqrtf(qq), 1e-12f);   // F.normalize eps, logit scale folded in
        const float ks = 1.0f / fmaxf(sqrtf(kk), 1e-12f);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            q[d] *= qs;
            Ks[i * D + d] = k[d] * ks;
            Vs[i * D + d] = v[d];
        }
        my_reg = shift > 0 ? region_of(ys, Hs, ws, shift) * 3 + region_of(xs, Ws, ws, shift) : 0;
        reg[i] = my_reg;
    }
    __syncthreads();
    if (!live) return;

    const float *bias = biasT + (size_t)head * N * N + i;   // biasT[h][j][i]
    float m = -INFINITY, l = 0.f, acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    for (int j0 = 0; j0 < N; j0 += CH) {
        float s[CH];
        float cm = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < CH; ++jj) {
            const int j = j0 + jj;
            const float4 *kp = reinterpret_cast<const float4 *>(Ks + j * D);
            float dot = 0.f;
#pragma unroll
            for (int d4 = 0; d4 < D / 4; ++d4) {
                const float4 kv = kp[d4];
                dot = fmaf(q[d4 * 4 + 0], kv.x, dot);
                dot = fmaf(q[d4 * 4 + 1], kv.y, dot);
                dot = fmaf(q[d4 * 4 + 2], kv.z, dot);
                dot = fmaf(q[d4 * 4 + 3], kv.w, dot);
            }
            dot += __ldg(bias + (size_t)j * N);
            if (reg[j] != my_reg) dot += -100.0f;
            s[jj] = dot;
            cm = fmaxf(cm, dot);
        }
        const float mn = fmaxf(m, cm);
        const float corr = __expf(m - mn);
        l *= corr;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] *= corr;
#pragma unroll
        for (int jj = 0; jj < CH; ++jj) {
            const float p = __expf(s[jj] - mn);
            l += p;
            const float4 *vp = reinterpret_cast<const float4 *>(Vs + (j0 + jj) * D);
#pragma unroll
            for (int d4 = 0; d4 < D / 4; ++d4) {
                const float4 vv = vp[d4];
                acc[d4 * 4 + 0] = fmaf(p, vv.x, acc[d4 * 4 + 0]);
                acc[d4 * 4 + 1] = fmaf(p, vv.y, acc[d4 * 4 + 1]);
                acc[d4 * 4 + 2] = fmaf(p, vv.z, acc[d4 * 4 + 2]);
                acc[d4 * 4 + 3] = fmaf(p, vv.w, acc[d4 * 4 + 3]);
            }
        }
        m = mn;
    }
    const float inv = 1.0f / l;
    bf16 *op = out + tok * C + head * D;
#pragma unroll
    for (int i8 = 0; i8 < D / 8; ++i8) {
        uint4 u;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(acc[i8 * 8 + 2 * k] * inv, acc[i8 * 8 + 2 * k + 1] * inv);
        *reinterpret_cast<uint4 *>(op + i8 * 8) = u;
    }
}


// ============================================================================================
// tcgen05 version for 256-token windows (16x16; stages 0-2 of swinv2_tiny_window16_256).
//   one CTA (128 threads) per (window, head); two CTAs per SM (105 KB smem, 256 TMEM columns each)
//   per 128-query half:   S[128x256] = Qn * Kn^T      tcgen05.mma  M128 N256 K32   (fp32 in TMEM)
//                         softmax, thread = query row = TMEM lane: pass 1 adds bias+mask, finds the row
//                         max and writes the logits back to TMEM; pass 2 exponentiates and writes
//                         P (bf16) into shared memory in the 128B-swizzled K-major layout
//                         O[128x32]  = P * V           tcgen05.mma  M128 N32  K256  (A = P from smem,
//                         B = V^T staged by a register transpose), O normalised by the row sum
// q and k are L2-normalised (and q multiplied by the logit scale) in fp32 while they are staged.
constexpr int TC_N = 256;                 // tokens per window
constexpr int TC_THREADS = 128;
constexpr int TC_SMEM_Q = 0;              // 128 rows x 64 B, SWIZZLE_64B
constexpr int TC_SMEM_K = 8192;           // 256 rows x 64 B, SWIZZLE_64B
constexpr int TC_SMEM_VT = 8192 + 16384;  // 4 k-blocks x (32 rows x 128 B), SWIZZLE_128B
constexpr int TC_SMEM_P = TC_SMEM_VT + 16384;   // 4 k-blocks x (128 rows x 128 B), SWIZZLE_128B
constexpr int TC_SMEM_MISC = TC_SMEM_P + 65536; // region ids (256 B) + mbarrier + tmem slot
constexpr int TC_SMEM_BYTES = TC_SMEM_MISC + 256 + 64 + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// shared-memory matrix descriptor, K-major; swizzle_bytes in {64, 128}
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, int swizzle_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * swizzle_bytes) >> 4) << 32;   // 8-row group pitch
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(swizzle_bytes == 128 ? 2 : 4) << 61;
    return d;
}
__device__ __forceinline__ uint32_t umma_idesc(int n) {
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n"
        "tcgen05.wait::st.sync.aligned;"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
          "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
          "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
          "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
          "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
          "r"(__float_as_uint(v[31])) : "memory");
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

__global__ void __launch_bounds__(TC_THREADS, 2)
window_attention_tc_kernel(const bf16 *__restrict__ qkv, const float *__restrict__ biasT, const float *__restrict__ scale,
                           bf16 *__restrict__ out, int Hs, int Ws, int C, int ws, int shift) {
    extern __shared__ uint8_t tc_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *reg = smem + TC_SMEM_MISC;                                   // [256] region ids
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + TC_SMEM_MISC + 256);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + TC_SMEM_MISC + 256 + 16);

    const int t = threadIdx.x, warp = t >> 5;
    const int nwx = Ws / ws, nwy = Hs / ws;
    const int win = blockIdx.x % (nwx * nwy), b = blockIdx.x / (nwx * nwy);
    const int head = blockIdx.y;

    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // token index (in the un-shifted image) of window row r, and its shift-mask region
    auto token_of = [&](int r, int &region) -> long long {
        const int ty = r / ws, tx = r - ty * ws;
        const int ys = (win / nwx) * ws + ty, xs = (win % nwx) * ws + tx;
        const int yo = (ys + shift) % Hs, xo = (xs + shift) % Ws;
        region = shift > 0 ? region_of(ys, Hs, ws, shift) * 3 + region_of(xs, Ws, ws, shift) : 0;
        return ((long long)b * Hs + yo) * Ws + xo;
    };

    // ---- stage K (normalised, SW64) and V^T (SW128) for both key halves
#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {
        const int r = t + rr * 128;
        int region;
        const long long tok = token_of(r, region);
        reg[r] = (uint8_t)region;
        const bf16 *base = qkv + tok * 3 * C + head * D;
        float k[D], v[D];
        load_head(base + C, k);
        load_head(base + 2 * C, v);
        float kk = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) kk = fmaf(k[d], k[d], kk);
        const float ks = 1.0f / fmaxf(sqrtf(kk), 1e-12f);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {   // 4 x 16-byte chunks of the 64-byte row, Swizzle<2,4,3>
            uint4 u;
            u.x = pack_bf16x2(k[c4 * 8 + 0] * ks, k[c4 * 8 + 1] * ks);
            u.y = pack_bf16x2(k[c4 * 8 + 2] * ks, k[c4 * 8 + 3] * ks);
            u.z = pack_bf16x2(k[c4 * 8 + 4] * ks, k[c4 * 8 + 5] * ks);
            u.w = pack_bf16x2(k[c4 * 8 + 6] * ks, k[c4 * 8 + 7] * ks);
            *reinterpret_cast<uint4 *>(smem + TC_SMEM_K + r * 64 + ((c4 ^ ((r >> 1) & 3)) << 4)) = u;
        }
        // V^T: element (d, key r) -> k-block r/64, row d, column r%64 (128-byte rows, Swizzle<3,4,3>)
        const int kb = r >> 6, col = r & 63;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int off = TC_SMEM_VT + kb * 4096 + d * 128 + (((col >> 3) ^ (d & 7)) << 4) + (col & 7) * 2;
            *reinterpret_cast<bf16 *>(smem + off) = __float2bfloat16_rn(v[d]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
    const float sc = scale[head];
    const float LOG2E = 1.4426950408889634f;
    uint32_t phase = 0;

#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int r = half * 128 + t;             // my query row inside the window
        int my_reg;
        const long long tok = token_of(r, my_reg);
        {   // stage normalised, scaled Q of this half (row t)
            float q[D];
            load_head(qkv + tok * 3 * C + head * D, q);
            float qq = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) qq = fmaf(q[d], q[d], qq);
            const float qs = sc / fmaxf(sqrtf(qq), 1e-12f);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint4 u;
                u.x = pack_bf16x2(q[c4 * 8 + 0] * qs, q[c4 * 8 + 1] * qs);
                u.y = pack_bf16x2(q[c4 * 8 + 2] * qs, q[c4 * 8 + 3] * qs);
                u.z = pack_bf16x2(q[c4 * 8 + 4] * qs, q[c4 * 8 + 5] * qs);
                u.w = pack_bf16x2(q[c4 * 8 + 6] * qs, q[c4 * 8 + 7] * qs);
                *reinterpret_cast<uint4 *>(smem + TC_SMEM_Q + t * 64 + ((c4 ^ ((t >> 1) & 3)) << 4)) = u;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da = umma_desc(smem_u32(smem + TC_SMEM_Q), 64), db = umma_desc(smem_u32(smem + TC_SMEM_K), 64);
            const uint32_t idesc = umma_idesc(TC_N);
            umma_f16(tmem, da, db, idesc, 0u);
            umma_f16(tmem, da + 2, db + 2, idesc, 1u);      // second K step: +32 bytes
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- pass 1: logits = S + bias + mask, row max, logits back to TMEM
        const float *bias = biasT + (size_t)head * TC_N * TC_N + r;     // biasT[h][key j][query i]: coalesced over i
        float m = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < TC_N; c0 += 32) {
            float v[32];
            tmem_ld32(t_row + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = v[j] + __ldg(bias + (size_t)(c0 + j) * TC_N);
                if (reg[c0 + j] != my_reg) x += -100.0f;
                v[j] = x;
                m = fmaxf(m, x);
            }
            tmem_st32(t_row + (uint32_t)c0, v);
        }
        // ---- pass 2: P = exp(logit - max) -> bf16, K-major SW128 rows in smem; row sum
        const float ml = m * LOG2E;
        float l = 0.f;
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
            uint8_t *prow = smem + TC_SMEM_P + kb * 16384 + (t >> 3) * 1024 + (t & 7) * 128;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float v[32];
                tmem_ld32(t_row + (uint32_t)(kb * 64 + hh * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = fast_exp2(fmaf(v[j], LOG2E, -ml));
                    l += v[j];
                }
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    uint4 u;
                    u.x = pack_bf16x2(v[c4 * 8 + 0], v[c4 * 8 + 1]);
                    u.y = pack_bf16x2(v[c4 * 8 + 2], v[c4 * 8 + 3]);
                    u.z = pack_bf16x2(v[c4 * 8 + 4], v[c4 * 8 + 5]);
                    u.w = pack_bf16x2(v[c4 * 8 + 6], v[c4 * 8 + 7]);
                    const int chunk = hh * 4 + c4;
                    *reinterpret_cast<uint4 *>(prow + ((chunk ^ (t & 7)) << 4)) = u;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                       // all S reads done, all P rows written
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = umma_idesc(D);
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) {
                const uint64_t da = umma_desc(smem_u32(smem + TC_SMEM_P + kb * 16384), 128);
                const uint64_t db = umma_desc(smem_u32(smem + TC_SMEM_VT + kb * 4096), 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {   // ---- epilogue: O / l -> bf16 -> out[token, head*32 ...]
            float o[32];
            tmem_ld32(t_row, o);
            const float inv = 1.0f / l;
            bf16 *op = out + tok * C + head * D;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint4 u;
                u.x = pack_bf16x2(o[c4 * 8 + 0] * inv, o[c4 * 8 + 1] * inv);
                u.y = pack_bf16x2(o[c4 * 8 + 2] * inv, o[c4 * 8 + 3] * inv);
                u.z = pack_bf16x2(o[c4 * 8 + 4] * inv, o[c4 * 8 + 5] * inv);
                u.w = pack_bf16x2(o[c4 * 8 + 6] * inv, o[c4 * 8 + 7] * inv);
                *reinterpret_cast<uint4 *>(op + c4 * 8) = u;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                       // O read by everyone before the next half overwrites TMEM / Q
    }
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}

}  // namespace

extern "C" int soccdpt_window_attention_fwd(const void *qkv, const float *bias, const float *scale, void *out,
                                            int batch, int Hs, int Ws, int C, int heads, int ws, int shift,
                                            soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(qkv && bias && scale && out, "window_attention: NULL pointer");
    SOCCDPT_REQUIRE(batch >= 1 && heads >= 1 && C == heads * D, "window_attention: head_dim must be 32 (C=%d heads=%d)", C, heads);
    SOCCDPT_REQUIRE(ws >= 1 && Hs % ws == 0 && Ws % ws == 0, "window_attention: window %d does not tile %dx%d", ws, Hs, Ws);
    SOCCDPT_REQUIRE(shift >= 0 && shift < ws, "window_attention: bad shift %d", shift);
    const int N = ws * ws;
    if (N == TC_N && !getenv("SOCCDPT_ATTN_CUDA_CORE")) {
        static bool tc_configured = false;
        if (!tc_configured) {
            SOCCDPT_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
            tc_configured = true;
        }
        dim3 grid((unsigned)(batch * (Hs / ws) * (Ws / ws)), (unsigned)heads);
        window_attention_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, soccdpt::as_stream(stream)>>>(
            static_cast<const bf16 *>(qkv), bias, scale, static_cast<bf16 *>(out), Hs, Ws, C, ws, shift);
        return soccdpt::check_launch("window_attention_tc_kernel");
    }
    SOCCDPT_REQUIRE(N % CH == 0 && N <= 1024, "window_attention: window tokens must be a multiple of %d and <= 1024 (got %d)", CH, N);
    const int threads = (N + 31) / 32 * 32;
    const size_t smem = (size_t)N * D * 2 * sizeof(float) + (size_t)N * sizeof(int);
    SOCCDPT_REQUIRE(smem <= 227 * 1024, "window_attention: window too large for shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        SOCCDPT_CUDA(cudaFuncSetAttribute(window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)(batch * (Hs / ws) * (Ws / ws)), (unsigned)heads);
    window_attention_kernel<<<grid, threads, smem, soccdpt::as_stream(stream)>>>(
        static_cast<const bf16 *>(qkv), bias, scale, static_cast<bf16 *>(out), Hs, Ws, C, ws, shift);
    return soccdpt::check_launch("window_attention_kernel");
}
