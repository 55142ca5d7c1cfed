// SwinV2 window attention for the small last-stage windows (8x8 of swin2_tiny_256, 12x12 of swin2_base_384), un-shifted
// (a window that covers its whole feature map is never shifted, timm SwinTransformerBlock._calc_window_shift).
//
// A (window, head) item is N = 64 / 144 tokens x d = 32: far too small for a 128-row tcgen05 tile (a four-windows-per-tile
// variant measured 57-61 us against 50 us for the CUDA-core kernel, profiles/r2_progress.md), and the CUDA-core kernel is
// bound by instruction issue (8650 instructions per warp at 16 % occupancy).  Here a warp owns 16 query rows and runs the
// item on warp-level tensor-core MMAs (m16n8k16 bf16, fp32 accumulate) in the flash-attention register layout: S = Q^.K^T
// stays in the accumulator fragments, the softmax reduces across the four lanes of a quad, and the normalised-later P is
// re-packed in registers as the A operand of P.V.  ~600 instructions per warp.
//
//   scores = normalize(q) . normalize(k)^T * scale[h] + cpb_bias[h][rel(i, j)]      (same convention as attention_tc.cu:
//   q^, k^ rounded to bf16 after the fp32 normalisation, P rounded to bf16, 1/l applied to the fp32 output)
#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;
constexpr int D = 32;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(unsigned addr, unsigned (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(unsigned addr, unsigned (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {      // bare MUFU.EX2 (exp2f adds a denormal-range fix-up: 3 more instructions per score)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 1 / max(sqrt(ss), 1e-12) = F.normalize's divisor (MUFU.RSQ: 2 ulp, the result is rounded to bf16 anyway)
__device__ __forceinline__ float inv_norm(float ss) { return fminf(rsqrtf(ss), 1e12f); }
__device__ __forceinline__ unsigned pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const unsigned *>(&h);
}
// 64-byte rows, two per 128-byte line: the 16-byte chunk index is XOR-ed with the line index so that the eight rows of an
// ldmatrix phase fall into eight different 16-byte bank groups
__device__ __forceinline__ unsigned tile_off(int row, int chunk) { return (unsigned)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

template <int WS>
__global__ void __launch_bounds__(WS * WS * 2)
window_attention_small_kernel(const bf16 *__restrict__ qkv, const float *__restrict__ bias_tab, const float *__restrict__ scale,
                              bf16 *__restrict__ out, int Hs, int Ws, int C) {
    constexpr int N = WS * WS, NT = N / 8, KS = N / 16, TW = 2 * WS - 1, THREADS = 2 * N;
    __shared__ __align__(128) unsigned char tiles[3 * N * 64];       // Q^, K^, V: [N][32] bf16, swizzled
    __shared__ float tab[TW * TW];
    const int head = blockIdx.y;
    for (int i = threadIdx.x; i < TW * TW; i += THREADS) tab[i] = bias_tab[(size_t)head * TW * TW + i];   // constants
    soccdpt::pdl_wait();
    const int nwx = Ws / WS, nwy = Hs / WS;
    const int win = blockIdx.x % (nwx * nwy), b = blockIdx.x / (nwx * nwy);
    const int y0 = (win / nwx) * WS, x0 = (win % nwx) * WS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- stage: thread -> (row, 16-byte chunk); the four chunks of a row sit in one quad (row norm by two shuffles)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int idx = it * THREADS + threadIdx.x, row = idx >> 2, ck = idx & 3;
        const size_t tok = ((size_t)b * Hs + y0 + row / WS) * Ws + x0 + row % WS;
        const bf16 *src = qkv + tok * 3 * C + head * D + ck * 8;
        uint4 u[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) u[m] = *reinterpret_cast<const uint4 *>(src + m * C);
#pragma unroll
        for (int m = 0; m < 2; ++m) {                   // q, k: F.normalize (eps 1e-12) in fp32, then bf16
            float f[8];
            const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u[m]);
            float ss = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 t2 = __bfloat1622float2(h[k]);
                f[2 * k] = t2.x; f[2 * k + 1] = t2.y;
                ss = fmaf(t2.x, t2.x, fmaf(t2.y, t2.y, ss));
            }
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            const float inv = inv_norm(ss);
            uint4 o;
            o.x = pack2(f[0] * inv, f[1] * inv); o.y = pack2(f[2] * inv, f[3] * inv);
            o.z = pack2(f[4] * inv, f[5] * inv); o.w = pack2(f[6] * inv, f[7] * inv);
            *reinterpret_cast<uint4 *>(tiles + m * N * 64 + tile_off(row, ck)) = o;
        }
        *reinterpret_cast<uint4 *>(tiles + 2 * N * 64 + tile_off(row, ck)) = u[2];
    }
    __syncthreads();

    const unsigned qb = smem_u32(tiles), kb = qb + N * 64, vb = kb + N * 64;
    const int g = lane >> 2, t = lane & 3, m8 = lane >> 3, r8 = lane & 7;
    // ---- S = Q^ K^T for this warp's 16 query rows
    unsigned aq[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ldsm_x4(qb + tile_off(warp * 16 + (m8 & 1) * 8 + r8, ks * 2 + (m8 >> 1)), aq[ks]);
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
        unsigned bk[4];
        ldsm_x4(kb + tile_off(j * 8 + r8, m8), bk);
        mma16816(s[j], aq[0], bk[0], bk[1]);
        mma16816(s[j], aq[1], bk[2], bk[3]);
    }
    // ---- logits, row max (rows g and g + 8 of the warp tile; a row lives in the four lanes of a quad)
    const float sc = scale[head];
    const int i0 = warp * 16 + g, i1 = i0 + 8;
    const float *t0 = tab + (i0 / WS + WS - 1) * TW + (i0 % WS) + WS - 1;     // bias[i][j] = t[-(ky * TW + kx)]
    const float *t1 = tab + (i1 / WS + WS - 1) * TW + (i1 % WS) + WS - 1;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int key = j * 8 + 2 * t + e, ko = (key / WS) * TW + key % WS;
            s[j][e] = fmaf(s[j][e], sc, t0[-ko]);
            s[j][2 + e] = fmaf(s[j][2 + e], sc, t1[-ko]);
            mx0 = fmaxf(mx0, s[j][e]);
            mx1 = fmaxf(mx1, s[j][2 + e]);
        }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    constexpr float LOG2E = 1.4426950408889634f;
    const float n0 = -mx0 * LOG2E, n1 = -mx1 * LOG2E;
    float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            s[j][e] = ex2(fmaf(s[j][e], LOG2E, n0));
            s[j][2 + e] = ex2(fmaf(s[j][2 + e], LOG2E, n1));
            l0 += s[j][e];
            l1 += s[j][2 + e];
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- O = P V: the score fragments of key tiles 2kk, 2kk + 1 are the A fragment of key step kk
    float o[4][4];
#pragma unroll
    for (int jd = 0; jd < 4; ++jd) o[jd][0] = o[jd][1] = o[jd][2] = o[jd][3] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
        unsigned ap[4];
        ap[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
        ap[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
        ap[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        ap[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int half = 0; half < 2; ++half) {          // d column tiles 2*half, 2*half + 1
            unsigned bv[4];
            ldsm_x4_trans(vb + tile_off(kk * 16 + (m8 & 1) * 8 + r8, half * 2 + (m8 >> 1)), bv);
            mma16816(o[half * 2], ap, bv[0], bv[1]);
            mma16816(o[half * 2 + 1], ap, bv[2], bv[3]);
        }
    }
    const float r0 = 1.0f / l0, r1 = 1.0f / l1;
    const size_t tok0 = ((size_t)b * Hs + y0 + i0 / WS) * Ws + x0 + i0 % WS;
    const size_t tok1 = ((size_t)b * Hs + y0 + i1 / WS) * Ws + x0 + i1 % WS;
    bf16 *o0 = out + tok0 * C + head * D + 2 * t, *o1 = out + tok1 * C + head * D + 2 * t;
#pragma unroll
    for (int jd = 0; jd < 4; ++jd) {
        *reinterpret_cast<unsigned *>(o0 + jd * 8) = pack2(o[jd][0] * r0, o[jd][1] * r0);
        *reinterpret_cast<unsigned *>(o1 + jd * 8) = pack2(o[jd][2] * r1, o[jd][3] * r1);
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// 16x16 windows (256 tokens; every stage-0..2 block of swin2_tiny_256), shifted or not.  One CTA (8 warps) per (window, head):
// Q^, K^, V of the item in shared memory (48 KB), a warp owns two 16-query tiles (= two window rows) one after the other and
// walks the 256 keys in four quarters with an online-softmax rescale in between, so two sets of score fragments (2 x 32 registers)
// plus the output fragments stay under 128 registers and two CTAs share an SM: one stages while the other computes.
// The shift mask is regenerated from the region ids exactly as timm builds attn_mask: only the last window row / column of
// a shifted block has more than one region, and only those CTAs run the masked instantiation of the half step.
struct RowState {          // online softmax of fragment rows g and g + 8 (per-lane partial sums, reduced across the quad at the end)
    float mx0, mx1, l0, l1;
};

// scores of 64 keys (window rows 4q .. 4q+3): 8 column tiles x 2 k-steps
__device__ __forceinline__ void scores_quarter(int q, const unsigned (&aq)[2][4], unsigned kb, int m8, int r8, float (&s)[8][4]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
        unsigned bk[4];
        ldsm_x4(kb + tile_off(q * 64 + j * 8 + r8, m8), bk);
        mma16816(s[j], aq[0], bk[0], bk[1]);
        mma16816(s[j], aq[1], bk[2], bk[3]);
    }
}

template <bool MASKED>
__device__ __forceinline__ void softmax_pv_quarter(int q, float (&s)[8][4], unsigned vb, const float *t0, const float *t1, float sc,
                                                   int m8, int r8, int t, bool qy_hi, int th, bool ymask, unsigned xbad0,
                                                   unsigned xbad1, RowState &st, float (&o)[4][4]) {
    constexpr float LOG2E = 1.4426950408889634f;
    float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ky = q * 4 + (j >> 1);                                    // key row of this column tile (compile-time)
        const bool ybad = MASKED && ymask && (qy_hi != (ky >= th));
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int ko = ky * 31 + (j & 1) * 8 + 2 * t + e;
            float v0 = fmaf(s[j][e], sc, t0[-ko]), v1 = fmaf(s[j][2 + e], sc, t1[-ko]);
            if (MASKED) {
                if (ybad || ((xbad0 >> ((j & 1) * 2 + e)) & 1u)) v0 -= 100.0f;
                if (ybad || ((xbad1 >> ((j & 1) * 2 + e)) & 1u)) v1 -= 100.0f;
            }
            s[j][e] = v0; s[j][2 + e] = v1;
            c0 = fmaxf(c0, v0); c1 = fmaxf(c1, v1);
        }
    }
    c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1)); c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1));
    c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2)); c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
    const float nm0 = fmaxf(st.mx0, c0), nm1 = fmaxf(st.mx1, c1);
    if (q > 0) {                                                            // rescale what the earlier quarters accumulated
        const float f0 = ex2((st.mx0 - nm0) * LOG2E), f1 = ex2((st.mx1 - nm1) * LOG2E);
        st.l0 *= f0; st.l1 *= f1;
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) { o[jd][0] *= f0; o[jd][1] *= f0; o[jd][2] *= f1; o[jd][3] *= f1; }
    }
    st.mx0 = nm0; st.mx1 = nm1;
    const float n0 = -nm0 * LOG2E, n1 = -nm1 * LOG2E;
    float p0[2] = {0.0f, 0.0f}, p1[2] = {0.0f, 0.0f};       // two partial sums per row: the FADD chain was the longest dependency
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            s[j][e] = ex2(fmaf(s[j][e], LOG2E, n0));
            s[j][2 + e] = ex2(fmaf(s[j][2 + e], LOG2E, n1));
            p0[e] += s[j][e];
            p1[e] += s[j][2 + e];
        }
    }
    st.l0 += p0[0] + p0[1];
    st.l1 += p1[0] + p1[1];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        unsigned ap[4];
        ap[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
        ap[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
        ap[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        ap[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int hd = 0; hd < 2; ++hd) {
            unsigned bv[4];
            ldsm_x4_trans(vb + tile_off(q * 64 + kk * 16 + (m8 & 1) * 8 + r8, hd * 2 + (m8 >> 1)), bv);
            mma16816(o[hd * 2], ap, bv[0], bv[1]);
            mma16816(o[hd * 2 + 1], ap, bv[2], bv[3]);
        }
    }
}

// One 16-query tile against the 256 keys in four 64-key quarters.  The score MMAs of quarter q + 1 are issued BEFORE the softmax
// of quarter q (two register sets of 32 accumulators), so inside one warp the tensor pipe works under the FMA / MUFU / LDS work
// of the softmax instead of alternating with it.
template <bool MASKED>
__device__ __forceinline__ void query_tile(const unsigned (&aq)[2][4], unsigned kb, unsigned vb, const float *t0, const float *t1,
                                           float sc, int m8, int r8, int t, bool qy_hi, int th, bool ymask, unsigned xbad0,
                                           unsigned xbad1, RowState &st, float (&o)[4][4]) {
    float sa[8][4], sb[8][4];
    scores_quarter(0, aq, kb, m8, r8, sa);
    scores_quarter(1, aq, kb, m8, r8, sb);
    softmax_pv_quarter<MASKED>(0, sa, vb, t0, t1, sc, m8, r8, t, qy_hi, th, ymask, xbad0, xbad1, st, o);
    scores_quarter(2, aq, kb, m8, r8, sa);
    softmax_pv_quarter<MASKED>(1, sb, vb, t0, t1, sc, m8, r8, t, qy_hi, th, ymask, xbad0, xbad1, st, o);
    scores_quarter(3, aq, kb, m8, r8, sb);
    softmax_pv_quarter<MASKED>(2, sa, vb, t0, t1, sc, m8, r8, t, qy_hi, th, ymask, xbad0, xbad1, st, o);
    softmax_pv_quarter<MASKED>(3, sb, vb, t0, t1, sc, m8, r8, t, qy_hi, th, ymask, xbad0, xbad1, st, o);
}

__global__ void __launch_bounds__(256, 2)
window_attention_mma16_kernel(const bf16 *__restrict__ qkv, const float *__restrict__ bias_tab, const float *__restrict__ scale,
                              bf16 *__restrict__ out, int Hs, int Ws, int C, int shift, int heads) {
    constexpr int WS = 16, N = 256, TW = 31;
    extern __shared__ __align__(128) unsigned char tiles[];          // Q^, K^, V: [256][32] bf16, swizzled; then the bias table
    float *tab = reinterpret_cast<float *>(tiles + 3 * N * 64);
    // CTA order = window-major, heads adjacent (the heads of a window read the same token rows)
    const int widx = blockIdx.x / heads, head = blockIdx.x - widx * heads;
    for (int i = threadIdx.x; i < TW * TW; i += 256) tab[i] = bias_tab[(size_t)head * TW * TW + i];   // constants
    soccdpt::pdl_wait();
    const int nwx = Ws / WS, nwy = Hs / WS;
    const int win = widx % (nwx * nwy), b = widx / (nwx * nwy);
    const int y0 = (win / nwx) * WS, x0 = (win % nwx) * WS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto token_of = [&](int row) -> size_t {        // roll(-shift): shifted[y] = x[(y + shift) % H]
        int yo = y0 + (row >> 4) + shift, xo = x0 + (row & 15) + shift;
        yo -= yo >= Hs ? Hs : 0;
        xo -= xo >= Ws ? Ws : 0;
        return ((size_t)b * Hs + yo) * Ws + xo;
    };
    {
        uint4 u[4][3];                              // all twelve 16-byte loads of this thread in flight at once
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = it * 256 + threadIdx.x, row = idx >> 2, ck = idx & 3;
            const bf16 *src = qkv + token_of(row) * 3 * C + head * D + ck * 8;
#pragma unroll
            for (int m = 0; m < 3; ++m) u[it][m] = __ldg(reinterpret_cast<const uint4 *>(src + m * C));
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = it * 256 + threadIdx.x, row = idx >> 2, ck = idx & 3;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                float f[8];
                const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u[it][m]);
                float ss = 0.0f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 t2 = __bfloat1622float2(h[k]);
                    f[2 * k] = t2.x; f[2 * k + 1] = t2.y;
                    ss = fmaf(t2.x, t2.x, fmaf(t2.y, t2.y, ss));
                }
                ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                const float inv = inv_norm(ss);
                uint4 q4;
                q4.x = pack2(f[0] * inv, f[1] * inv); q4.y = pack2(f[2] * inv, f[3] * inv);
                q4.z = pack2(f[4] * inv, f[5] * inv); q4.w = pack2(f[6] * inv, f[7] * inv);
                *reinterpret_cast<uint4 *>(tiles + m * N * 64 + tile_off(row, ck)) = q4;
            }
            *reinterpret_cast<uint4 *>(tiles + 2 * N * 64 + tile_off(row, ck)) = u[it][2];
        }
    }
    __syncthreads();

    const unsigned qb = smem_u32(tiles), kb = qb + N * 64, vb = kb + N * 64;
    const int g = lane >> 2, t = lane & 3, m8 = lane >> 3, r8 = lane & 7;
    const float sc = scale[head];
    // regions (timm: slices (0,-ws), (-ws,-shift), (-shift,None) of the shifted image): a second region only in the last window row / column
    const int th = WS - shift;
    const bool ymask = shift > 0 && (win / nwx) == nwy - 1, xmask = shift > 0 && (win % nwx) == nwx - 1;
    unsigned xbad0 = 0u, xbad1 = 0u;                // bit (parity * 2 + e): key column (parity*8 + 2t + e) is in the other x region
    if (xmask) {
#pragma unroll
        for (int pe = 0; pe < 4; ++pe) {
            const bool khi = ((pe >> 1) * 8 + 2 * t + (pe & 1)) >= th;
            xbad0 |= (unsigned)(khi != (g >= th)) << pe;
            xbad1 |= (unsigned)(khi != (g + 8 >= th)) << pe;
        }
    }
#pragma unroll 1
    for (int qt = warp; qt < 16; qt += 8) {          // query tile = window row qt; rows g (qx = g) and g + 8 of the fragment
        unsigned aq[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) ldsm_x4(qb + tile_off(qt * 16 + (m8 & 1) * 8 + r8, ks * 2 + (m8 >> 1)), aq[ks]);
        const float *t0 = tab + (qt + 15) * TW + g + 15, *t1 = t0 + 8;
        RowState st = {-INFINITY, -INFINITY, 0.0f, 0.0f};
        float o[4][4];
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) o[jd][0] = o[jd][1] = o[jd][2] = o[jd][3] = 0.0f;
        if (ymask || xmask) query_tile<true>(aq, kb, vb, t0, t1, sc, m8, r8, t, qt >= th, th, ymask, xbad0, xbad1, st, o);
        else query_tile<false>(aq, kb, vb, t0, t1, sc, m8, r8, t, false, th, false, 0u, 0u, st, o);
        float l0 = st.l0, l1 = st.l1;
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float r0 = 1.0f / l0, r1 = 1.0f / l1;
        bf16 *o0 = out + token_of(qt * 16 + g) * C + head * D + 2 * t, *o1 = out + token_of(qt * 16 + g + 8) * C + head * D + 2 * t;
#pragma unroll
        for (int jd = 0; jd < 4; ++jd) {
            *reinterpret_cast<unsigned *>(o0 + jd * 8) = pack2(o[jd][0] * r0, o[jd][1] * r0);
            *reinterpret_cast<unsigned *>(o1 + jd * 8) = pack2(o[jd][2] * r1, o[jd][3] * r1);
        }
    }
}

}  // namespace

namespace soccdpt {
// ws in {8, 12}, shift == 0 (checked by the caller)
int launch_window_attention_small(const void *qkv, const float *biasT, const float *scale, void *out, int batch, int Hs, int Ws,
                                  int C, int heads, int ws, cudaStream_t st) {
    dim3 grid((unsigned)(batch * (Hs / ws) * (Ws / ws)), (unsigned)heads);
    if (ws == 8) {
        SOCCDPT_CUDA(launch_pdl(PDL_ATTENTION, window_attention_small_kernel<8>, grid, dim3(128), 0, st, static_cast<const bf16 *>(qkv),
                                biasT, scale, static_cast<bf16 *>(out), Hs, Ws, C));
    } else {
        SOCCDPT_CUDA(launch_pdl(PDL_ATTENTION, window_attention_small_kernel<12>, grid, dim3(288), 0, st, static_cast<const bf16 *>(qkv),
                                biasT, scale, static_cast<bf16 *>(out), Hs, Ws, C));
    }
    return check_launch("window_attention_small_kernel");
}

// 16x16 windows, any shift in [0, 16)
int launch_window_attention_mma16(const void *qkv, const float *biasT, const float *scale, void *out, int batch, int Hs, int Ws,
                                  int C, int heads, int shift, cudaStream_t st) {
    const int items = batch * (Hs / 16) * (Ws / 16) * heads;
    constexpr size_t smem = 3 * 256 * 64 + 964 * sizeof(float);
    static SmemAttr configured;
    if (configured.need(smem))
        SOCCDPT_CUDA(cudaFuncSetAttribute(window_attention_mma16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SOCCDPT_CUDA(launch_pdl(PDL_ATTENTION, window_attention_mma16_kernel, dim3((unsigned)items), dim3(256), smem, st,
                            static_cast<const bf16 *>(qkv), biasT, scale, static_cast<bf16 *>(out), Hs, Ws, C, shift, heads));
    return check_launch("window_attention_mma16_kernel");
}
}  // namespace soccdpt
