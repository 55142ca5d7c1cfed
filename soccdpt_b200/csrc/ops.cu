// Bandwidth-bound glue kernels of the network path + the CUDA-core reference convolution.
// Layout everywhere: NHWC / token-major bf16 activations, fp32 math inside the kernels.
#include <cstdlib>

#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == SOCCDPT_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == SOCCDPT_ACT_GELU) return gelu_erf(v);
    return v;
}

__device__ __forceinline__ void unpack8(const uint4 &u, float f[8]) {
    const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(p[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float f[8]) {
    uint4 u;
    __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------ reference convolution
// One warp per output pixel; lane l owns output channels l, l+32, ...  Same epilogue contract as
// the tcgen05 kernel (conv_tcgen05.cu) -- used by the tests as an on-device cross-check.
__global__ void __launch_bounds__(256) conv_ref_kernel(const soccdpt_conv_t c) {
    const int lane = threadIdx.x & 31;
    const int st = c.stride > 1 ? c.stride : 1, pad = c.KH / 2 - c.pad_trim;
    const int Ho = (c.H + st - 1) / st, Wo = (c.W + st - 1) / st;
    const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long total = (long long)c.N * Ho * Wo;
    if (pix >= total) return;
    const int w0 = (int)(pix % Wo), h0 = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
    const bf16 *x = static_cast<const bf16 *>(c.x);
    const bf16 *wgt = static_cast<const bf16 *>(c.wgt);
    const int taps = c.KH * c.KW;
    float proj[4] = {0.f, 0.f, 0.f, 0.f};
    for (int co = lane; co < c.Cout; co += 32) {
        float acc = 0.0f;
        for (int kh = 0; kh < c.KH; ++kh) {
            const int hh = h0 * st + kh - pad;
            if (hh < 0 || hh >= c.H) continue;
            for (int kw = 0; kw < c.KW; ++kw) {
                const int ww = w0 * st + kw - pad;
                if (ww < 0 || ww >= c.W) continue;
                const bf16 *xp = x + (((long long)n * c.H + hh) * c.W + ww) * c.Cin;
                const bf16 *wp = wgt + ((long long)co * taps + kh * c.KW + kw) * c.Cin;
                for (int ci = 0; ci < c.Cin; ci += 8) {
                    float a[8], b[8];
                    unpack8(*reinterpret_cast<const uint4 *>(xp + ci), a);
                    unpack8(*reinterpret_cast<const uint4 *>(wp + ci), b);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc = fmaf(a[k], b[k], acc);
                }
            }
        }
        float v = acc + (c.bias ? c.bias[co] : 0.0f);
        v = apply_act(v, c.act);
        if (c.qk_heads > 0 && co < 64 * c.qk_heads) {   // cosine-attention epilogue: the 32 lanes hold one head of q or k
            float inv = 1.0f / fmaxf(sqrtf(warp_sum(v * v)), 1e-12f);
            if (co < 32 * c.qk_heads) inv *= c.qk_scale[co >> 5];
            v *= inv;
        }
        const long long o = pix * c.Cout + co;
        if (c.res1) v += __bfloat162float(static_cast<const bf16 *>(c.res1)[o]);
        if (c.res2) v += __bfloat162float(static_cast<const bf16 *>(c.res2)[o]);
        if (c.up_src) {     // bilinear x2, align_corners=True, of the low-resolution map at this pixel
            const bf16 *u = static_cast<const bf16 *>(c.up_src);
            const float fy = (float)(c.up_h - 1) / (float)(2 * c.up_h - 1) * (float)h0, fx = (float)(c.up_w - 1) / (float)(2 * c.up_w - 1) * (float)w0;
            const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + (y0 < c.up_h - 1 ? 1 : 0), x1 = x0 + (x0 < c.up_w - 1 ? 1 : 0);
            const float ly = fy - (float)y0, lx = fx - (float)x0;
            auto at = [&](int yy, int xx) { return __bfloat162float(u[(((long long)n * c.up_h + yy) * c.up_w + xx) * c.Cout + co]); };
            v += (1.0f - ly) * ((1.0f - lx) * at(y0, x0) + lx * at(y0, x1)) + ly * ((1.0f - lx) * at(y1, x0) + lx * at(y1, x1));
        }
        if (c.y) static_cast<bf16 *>(c.y)[o] = __float2bfloat16_rn(v);
        if (c.y_relu) static_cast<bf16 *>(c.y_relu)[o] = __float2bfloat16_rn(fmaxf(v, 0.0f));
        for (int p = 0; p < c.proj_n; ++p) proj[p] = fmaf(c.proj_w[p * c.Cout + co], v, proj[p]);
    }
    for (int p = 0; p < c.proj_n; ++p) {
        float s = warp_sum(proj[p]);
        if (lane == 0) {
            s += c.proj_b[p];
            if (c.proj_relu) s = fmaxf(s, 0.0f);
            c.proj_out[pix * c.proj_n + p] = s;
        }
    }
}

// ------------------------------------------------------------------ patch embed (conv4x4 s4 + LN)
// Weights live transposed in shared memory (wT[k][e]: lanes read consecutive channels, conflict free);
// every warp handles 4 horizontally adjacent tokens per iteration so each weight read feeds 4 FMAs.
// Lane l owns channels l, l+32, ... (E <= 128).  Two-pass LayerNorm (eps 1e-5) in fp32 on the warp.
constexpr int PE_TOK = 4;
template <int NCH>      // 32-channel slices per lane: 3 for E = 96 (a fourth, all-zero slice cost a quarter of the FMAs), 4 for E = 128
__global__ void __launch_bounds__(256)
patch_embed_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                   const float *__restrict__ g, const float *__restrict__ be, bf16 *__restrict__ out,
                   float *__restrict__ out_f32, int B, int H, int W, int E) {
    extern __shared__ __align__(16) float pe_smem[];
    float4 *wT4 = reinterpret_cast<float4 *>(pe_smem);   // [12][E]: taps 4*k4 .. 4*k4+3 of channel e (one LDS.128 feeds 4 FMAs per token)
    float *stage = pe_smem + 48 * E;           // [8 warps][PE_TOK][48]
    for (int i = threadIdx.x; i < 48 * E; i += 256) {
        const int e = i / 48, k = i - e * 48;
        pe_smem[((k >> 2) * E + e) * 4 + (k & 3)] = w[i];
    }
    soccdpt::pdl_wait();        // the weights above are constants; the frames below may come from the previous kernel
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ph = H / 4, pw = W / 4;
    const int groups_per_row = pw / PE_TOK;
    const unsigned groups = (unsigned)B * (unsigned)ph * (unsigned)groups_per_row;      // < 2^31 (checked on the host)
    float *in = stage + warp * PE_TOK * 48;
    // per-lane constants of the channels this lane owns (loop invariant: they were re-read from L1 for every token)
    float bias_r[NCH], gam_r[NCH], bet_r[NCH];
    bool own[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        own[j] = lane + 32 * j < E;
        bias_r[j] = own[j] ? b[lane + 32 * j] : 0.0f;
        gam_r[j] = own[j] ? g[lane + 32 * j] : 0.0f;
        bet_r[j] = own[j] ? be[lane + 32 * j] : 0.0f;
    }
    // staging slots of this lane's (up to two) input float4s: f = lane, lane + 32 -> r = f >> 2 (= ci*4 + i), q = f & 3
    const int q_tok = lane & 3, r0 = lane >> 2, r1 = (lane + 32) >> 2;
    const size_t plane = (size_t)H * W;
    const size_t off0 = (size_t)(r0 >> 2) * plane + (size_t)(r0 & 3) * W + q_tok * 4;
    const size_t off1 = (size_t)(r1 >> 2) * plane + (size_t)(r1 & 3) * W + q_tok * 4;
    float *d0 = in + q_tok * 48 + r0 * 4, *d1 = in + q_tok * 48 + r1 * 4;      // ci*16 + i*4 == r*4
    for (unsigned gi = blockIdx.x * 8u + warp; gi < groups; gi += gridDim.x * 8u) {
        const unsigned gx = gi % (unsigned)groups_per_row, rowi = gi / (unsigned)groups_per_row;
        const unsigned ty = rowi % (unsigned)ph, n = rowi / (unsigned)ph;
        // 12 image rows (3 channels x 4 rows) x 16 consecutive floats = 48 float4, coalesced 64 B segments
        const float *xg = x + ((size_t)n * 3 * H + (size_t)ty * 4) * W + (size_t)gx * (PE_TOK * 4);
        const float4 v0 = *reinterpret_cast<const float4 *>(xg + off0);
        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < 16) v1 = *reinterpret_cast<const float4 *>(xg + off1);
        *reinterpret_cast<float4 *>(d0) = v0;
        if (lane < 16) *reinterpret_cast<float4 *>(d1) = v1;
        __syncwarp();
        float acc[PE_TOK][NCH];
#pragma unroll
        for (int t = 0; t < PE_TOK; ++t)
#pragma unroll
            for (int j = 0; j < NCH; ++j) acc[t][j] = bias_r[j];
#pragma unroll 2
        for (int k4 = 0; k4 < 12; ++k4) {
            float4 wv[NCH];
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                wv[j] = own[j] ? wT4[k4 * E + lane + 32 * j] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < PE_TOK; ++t) {
                const float4 xv = *reinterpret_cast<const float4 *>(in + t * 48 + k4 * 4);
#pragma unroll
                for (int j = 0; j < NCH; ++j)     // ascending tap order, like the scalar loop it replaces
                    acc[t][j] = fmaf(xv.w, wv[j].w, fmaf(xv.z, wv[j].z, fmaf(xv.y, wv[j].y, fmaf(xv.x, wv[j].x, acc[t][j]))));
            }
        }
        __syncwarp();
        const size_t tok0 = ((size_t)n * ph + ty) * pw + (size_t)gx * PE_TOK;
        bf16 *o16 = out + tok0 * E + lane;
        float *o32 = out_f32 ? out_f32 + tok0 * E + lane : nullptr;
#pragma unroll
        for (int t = 0; t < PE_TOK; ++t) {
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < NCH; ++j) s += own[j] ? acc[t][j] : 0.0f;
            const float mean = warp_sum(s) / (float)E;
            float q = 0.0f;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const float d = acc[t][j] - mean;
                q += own[j] ? d * d : 0.0f;
            }
            const float rstd = rsqrtf(warp_sum(q) / (float)E + 1e-5f);
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                if (own[j]) {
                    const float v = (acc[t][j] - mean) * rstd * gam_r[j] + bet_r[j];
                    o16[t * E + 32 * j] = __float2bfloat16_rn(v);
                    if (o32) o32[t * E + 32 * j] = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ patch embed on the tensor cores
// The CUDA-core kernel above is bound by FMA issue (48 x E FMAs per token: 130 us for 64 frames against a 31 us HBM floor).
// Here a warp owns 16 horizontally adjacent tokens: A (16 tokens x 48 taps) comes straight from the fp32 image in the
// m16n8k16 fragment layout (k-step = input channel, k = ky*4 + kx: each lane reads float2 pieces, a warp-wide load covers two
// full 128-byte row segments), split into bf16 hi + lo; the weights are split the same way at kernel start and kept in shared
// memory as ready-made B fragments.  hi*hi + lo*hi + hi*lo with fp32 accumulation reproduces the fp32 convolution to ~2^-17
// (the dropped lo*lo term), i.e. far below the bf16 rounding of the output.  LayerNorm runs on the accumulator fragments
// (row sums across the four lanes of a quad).
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_bf16x2(float x, float y, unsigned &hi, unsigned &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const unsigned *>(&h);
    lo = *reinterpret_cast<const unsigned *>(&l);
}

template <int NT>      // 8-channel column tiles: E = 8 * NT (12 for swin2_tiny, 16 for swin2_base)
__global__ void __launch_bounds__(256, 2)
patch_embed_mma_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                       const float *__restrict__ g, const float *__restrict__ be, bf16 *__restrict__ out,
                       float *__restrict__ out_f32, int B, int H, int W) {
    constexpr int E = 8 * NT;
    // B fragments [split][channel][column tile][lane] -> {b0, b1}; then bias / gamma / beta
    __shared__ uint2 wfrag[2 * 3 * NT * 32];
    __shared__ float cst[3 * E];
    for (int i = threadIdx.x; i < 3 * NT * 32; i += 256) {
        const int lane = i & 31, j = (i >> 5) % NT, c = i / (32 * NT);
        const int gq = lane >> 2, t = lane & 3;
        const float *wr = w + (size_t)(8 * j + gq) * 48 + c * 16 + 2 * t;       // B[k][n] = w[n][k]
        uint2 hi, lo;
        split_bf16x2(wr[0], wr[1], hi.x, lo.x);
        split_bf16x2(wr[8], wr[9], hi.y, lo.y);
        wfrag[i] = hi;
        wfrag[3 * NT * 32 + i] = lo;
    }
    for (int i = threadIdx.x; i < E; i += 256) { cst[i] = b[i]; cst[E + i] = g[i]; cst[2 * E + i] = be[i]; }
    soccdpt::pdl_wait();        // the weights above are constants; the frames below may come from the previous kernel
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, t = lane & 3;
    const int ph = H / 4, pw = W / 4, tpr = pw / 16;                     // 16-token tiles per token row
    const unsigned tiles = (unsigned)B * (unsigned)ph * (unsigned)tpr;   // < 2^31 (checked on the host)
    const size_t plane = (size_t)H * W;
    // this lane's pieces of a tile: rows (tokens) gq and gq + 8, taps (ky = t>>1 [+2], kx = (t&1)*2 .. +1)
    const size_t lane_off = (size_t)(t >> 1) * W + (size_t)gq * 4 + (size_t)(t & 1) * 2;
    auto tile_base = [&](unsigned ti) {
        const unsigned tx = ti % (unsigned)tpr, rowi = ti / (unsigned)tpr;
        const unsigned ty = rowi % (unsigned)ph, n = rowi / (unsigned)ph;
        return x + ((size_t)n * 3 * H + (size_t)ty * 4) * W + (size_t)tx * 64 + lane_off;
    };
    auto load_tile = [&](const float *p, float2 (&v)[3][4]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *pc = p + c * plane;
            v[c][0] = __ldcs(reinterpret_cast<const float2 *>(pc));                       // token gq,     ky
            v[c][1] = __ldcs(reinterpret_cast<const float2 *>(pc + 32));                  // token gq + 8, ky
            v[c][2] = __ldcs(reinterpret_cast<const float2 *>(pc + 2 * (size_t)W));       // token gq,     ky + 2
            v[c][3] = __ldcs(reinterpret_cast<const float2 *>(pc + 2 * (size_t)W + 32));  // token gq + 8, ky + 2
        }
    };
    unsigned ti = blockIdx.x * 8u + warp;
    float2 nxt[3][4];
    if (ti < tiles) load_tile(tile_base(ti), nxt);
    for (; ti < tiles; ti += gridDim.x * 8u) {
        unsigned ah[3][4], al[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) split_bf16x2(nxt[c][r].x, nxt[c][r].y, ah[c][r], al[c][r]);
        const unsigned tn = ti + gridDim.x * 8u;
        if (tn < tiles) load_tile(tile_base(tn), nxt);        // in flight under the MMAs and the LayerNorm of this tile
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float b0 = cst[8 * j + 2 * t], b1 = cst[8 * j + 2 * t + 1];
            acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b0; acc[j][3] = b1;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const uint2 wh = wfrag[(c * NT + j) * 32 + lane], wl = wfrag[((3 + c) * NT + j) * 32 + lane];
                mma_bf16_16816(acc[j], ah[c], wl.x, wl.y);      // small terms first
                mma_bf16_16816(acc[j], al[c], wh.x, wh.y);
                mma_bf16_16816(acc[j], ah[c], wh.x, wh.y);
            }
        // LayerNorm (eps 1e-5) of rows gq (acc[.][0..1]) and gq + 8 (acc[.][2..3]); a row lives in the 4 lanes of a quad
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int j = 0; j < NT; ++j) { s0 += acc[j][0] + acc[j][1]; s1 += acc[j][2] + acc[j][3]; }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float m0 = s0 / (float)E, m1 = s1 / (float)E;
        float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            acc[j][0] -= m0; acc[j][1] -= m0; acc[j][2] -= m1; acc[j][3] -= m1;
            q0 += acc[j][0] * acc[j][0] + acc[j][1] * acc[j][1];
            q1 += acc[j][2] * acc[j][2] + acc[j][3] * acc[j][3];
        }
        q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
        q0 += __shfl_xor_sync(0xffffffffu, q0, 2); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
        const float r0 = rsqrtf(q0 / (float)E + 1e-5f), r1 = rsqrtf(q1 / (float)E + 1e-5f);
        const size_t tok0 = (size_t)ti * 16;          // tiles are numbered in token order
        bf16 *o16a = out + (tok0 + gq) * E + 2 * t, *o16b = o16a + 8 * E;
        float *o32a = out_f32 ? out_f32 + (tok0 + gq) * E + 2 * t : nullptr, *o32b = out_f32 ? o32a + 8 * E : nullptr;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float ga = cst[E + 8 * j + 2 * t], gb = cst[E + 8 * j + 2 * t + 1];
            const float ba = cst[2 * E + 8 * j + 2 * t], bb = cst[2 * E + 8 * j + 2 * t + 1];
            const float v0 = acc[j][0] * r0 * ga + ba, v1 = acc[j][1] * r0 * gb + bb;
            const float v2 = acc[j][2] * r1 * ga + ba, v3 = acc[j][3] * r1 * gb + bb;
            *reinterpret_cast<__nv_bfloat162 *>(o16a + 8 * j) = __floats2bfloat162_rn(v0, v1);
            *reinterpret_cast<__nv_bfloat162 *>(o16b + 8 * j) = __floats2bfloat162_rn(v2, v3);
            if (o32a) {
                *reinterpret_cast<float2 *>(o32a + 8 * j) = make_float2(v0, v1);
                *reinterpret_cast<float2 *>(o32b + 8 * j) = make_float2(v2, v3);
            }
        }
    }
}

// ------------------------------------------------------------------ y = res + LayerNorm(t)
// A group of G lanes (16 or 32) owns one row; every lane keeps its ITERS x 8 elements in registers, so the
// row is read once (16-byte bf16 loads).  C = 96 uses half-warps (12 of 16 lanes busy instead of 12 of 32).
// Two residual flavours: `res` (bf16) or `master` (fp32 residual stream, updated in place; `accumulate`
// = 0 overwrites it).  The Swin residual stream is kept in fp32 so that 24 post-norm additions do not
// each round to bf16; y is the bf16 copy the next GEMM consumes.
template <int G, int ITERS>
__global__ void __launch_bounds__(256)
layernorm_kernel(const bf16 *__restrict__ t, const bf16 *__restrict__ res, float *__restrict__ master, int accumulate,
                 const float *__restrict__ gamma, const float *__restrict__ beta, bf16 *__restrict__ y, long long rows,
                 int C, float eps) {
    soccdpt::pdl_wait();
    const int gl = threadIdx.x % G;                              // lane inside the group
    const long long row = (long long)blockIdx.x * (256 / G) + threadIdx.x / G;
    const bool live = row < rows;                                // keep dead groups in the shuffles
    const int chunks = C / 8;
    const long long base = (live ? row : 0) * C;
    const uint4 *tp = reinterpret_cast<const uint4 *>(t + base);
    const uint4 *rp = res ? reinterpret_cast<const uint4 *>(res + base) : nullptr;
    float4 *mp = master ? reinterpret_cast<float4 *>(master + base) : nullptr;
    const bool add_master = mp && accumulate;
    // every load of the row -- branch output AND residual -- is issued before the first reduction: one memory round trip
    // per thread instead of two (the kernel is a pure HBM stream; bytes in flight per SM are what sets its speed)
    uint4 traw[ITERS], rraw[ITERS];
    float4 ma[ITERS], mb[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int k = gl + it * G;
        if (k < chunks) {
            traw[it] = tp[k];
            if (rp) rraw[it] = rp[k];
            if (add_master) { ma[it] = mp[2 * k]; mb[it] = mp[2 * k + 1]; }
        }
    }
    float f[ITERS][8];
    float s = 0.0f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        if (gl + it * G < chunks) {
            unpack8(traw[it], f[it]);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += f[it][i];
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.0f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        if (gl + it * G < chunks) {
#pragma unroll
            for (int i = 0; i < 8; ++i) q += (f[it][i] - mean) * (f[it][i] - mean);
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
    if (!live) return;
    uint4 *yp = reinterpret_cast<uint4 *>(y + base);
    const bool add = rp || add_master;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int k = gl + it * G;
        if (k < chunks) {
            float r[8], o8[8];
            if (rp) unpack8(rraw[it], r);
            if (add_master) {
                const float4 a = ma[it], b2 = mb[it];
                r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b2.x; r[5] = b2.y; r[6] = b2.z; r[7] = b2.w;
            }
            const float4 g0 = *reinterpret_cast<const float4 *>(gamma + k * 8), g1 = *reinterpret_cast<const float4 *>(gamma + k * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4 *>(beta + k * 8), b1 = *reinterpret_cast<const float4 *>(beta + k * 8 + 4);
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float v = (f[it][i] - mean) * rstd * gg[i] + bb[i];
                o8[i] = add ? r[i] + v : v;
            }
            if (mp) {
                mp[2 * k] = make_float4(o8[0], o8[1], o8[2], o8[3]);
                mp[2 * k + 1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
            }
            yp[k] = pack8(o8);
        }
    }
}

int launch_layernorm(const bf16 *t, const bf16 *res, float *master, int accumulate, const float *gamma, const float *beta,
                     bf16 *y, long long rows, int C, float eps, cudaStream_t st) {
#define SOCC_LN(G, IT)                                                                                              \
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, layernorm_kernel<G, IT>, dim3((unsigned)((rows + (256 / G) - 1) / (256 / G))), dim3(256), 0, \
                                     st, t, res, master, accumulate, gamma, beta, y, rows, C, eps))
    // (C = 3 * 2^k as C / 24 lanes x 3 chunks -- every lane busy, three chunks of loads in flight per thread -- was measured
    // 3 % SLOWER than a quarter of the lanes idle: 104 registers halve the resident warps; profiles/r1_progress.md step 15)
    if (C <= 128) SOCC_LN(16, 1);
    else if (C <= 256) SOCC_LN(32, 1);
    else if (C <= 512) SOCC_LN(32, 2);
    else if (C <= 1024) SOCC_LN(32, 4);
    else if (C <= 2048) SOCC_LN(32, 8);
    else {
        soccdpt::set_error("layernorm: C = %d > 2048 is not supported", C);
        return SOCCDPT_E_INVALID;
    }
#undef SOCC_LN
    return soccdpt::check_launch("layernorm_kernel");
}

// ------------------------------------------------------------------ PatchMerging gather
__global__ void __launch_bounds__(256)
patch_merge_gather_kernel(const bf16 *__restrict__ x, bf16 *__restrict__ y, int B, int H, int W, int C) {
    const int chunks = C / 8;
    const long long total = (long long)B * (H / 2) * (W / 2) * 4 * chunks;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int ck = (int)(i % chunks);
        const int q = (int)((i / chunks) % 4);
        const long long o = i / (4 * chunks);
        const int ox = (int)(o % (W / 2)), oy = (int)((o / (W / 2)) % (H / 2)), n = (int)(o / ((long long)(W / 2) * (H / 2)));
        const int dh = q & 1, dw = q >> 1;  // x0:(0,0) x1:(1,0) x2:(0,1) x3:(1,1)
        const uint4 v = *reinterpret_cast<const uint4 *>(x + (((long long)n * H + oy * 2 + dh) * W + ox * 2 + dw) * C + ck * 8);
        *reinterpret_cast<uint4 *>(y + (o * 4 + q) * C + ck * 8) = v;
    }
}

// ------------------------------------------------------------------ bilinear, align_corners=True, NHWC bf16
// blockIdx.y = output row (n, Y): the vertical taps / weights are block-uniform; a thread owns 8 channels of one output
// pixel.  32-bit index math only (the first version spent five emulated 64-bit divisions per 16 output bytes).
__global__ void __launch_bounds__(256)
upsample_bilinear_kernel(const bf16 *__restrict__ x, bf16 *__restrict__ y, unsigned rows, int h, int w, int H, int W, int C) {
    soccdpt::pdl_wait();
    const unsigned chunks = (unsigned)C / 8u;
    for (unsigned row = blockIdx.y; row < rows; row += gridDim.y) {
    const unsigned n = row / (unsigned)H, Y = row - n * (unsigned)H;
    const float sh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f;
    const float sw = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.0f;
    const float fy = sh * (float)Y;
    const int y0 = (int)fy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
    const float ly = fy - (float)y0;
    const bf16 *r0 = x + ((size_t)n * h + y0) * w * C, *r1 = x + ((size_t)n * h + y1) * w * C;
    bf16 *yr = y + (size_t)row * W * C;
    const unsigned items = (unsigned)W * chunks;
#pragma unroll 4
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < items; i += gridDim.x * 256u) {
        const unsigned X = i / chunks, ck = i - X * chunks;
        const float fx = sw * (float)X;
        const int x0 = (int)fx, x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float lx = fx - (float)x0;
        const unsigned o0 = (unsigned)x0 * (unsigned)C + ck * 8u, o1 = (unsigned)x1 * (unsigned)C + ck * 8u;
        float a[8], b[8], c2[8], d[8], r[8];
        unpack8(__ldg(reinterpret_cast<const uint4 *>(r0 + o0)), a);
        unpack8(__ldg(reinterpret_cast<const uint4 *>(r0 + o1)), b);
        unpack8(__ldg(reinterpret_cast<const uint4 *>(r1 + o0)), c2);
        unpack8(__ldg(reinterpret_cast<const uint4 *>(r1 + o1)), d);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            r[k] = (1.0f - ly) * ((1.0f - lx) * a[k] + lx * b[k]) + ly * ((1.0f - lx) * c2[k] + lx * d[k]);
        *reinterpret_cast<uint4 *>(yr + (size_t)i * 8u) = pack8(r);
    }
    }
}

// Exact x2 case (every FeatureFusion upsample of the decoder): with align_corners=True the output rows {2k+1, 2k+2} share
// their two source rows (k, k+1) -- likewise the columns -- so the four source chunks of a 2x2 output block are loaded ONCE
// (one 16-byte load per 16-byte store instead of four).  Row / column groups: {0}, {1,2}, ..., {2h-3, 2h-2}, {2h-1}: h + 1 of them.
// A thread owns one (column group, 16-byte channel chunk) and MARCHES down kUpSeg row groups: the bottom source row of a group
// is the top row of the next, so every group costs two loads (issued before the arithmetic of the previous group), and the
// column weights, the chunk pointers and the unpacked top row stay in registers.  The first version (one thread per group item)
// spent 60 % of its 456 instructions per thread on index arithmetic (issue active 66 %) at 181 us for the 64 -> 128 level; this
// one needs ~190 and is latency-bound instead: 0.278 -> 0.248 ms per step over the four levels.
constexpr int kUpSeg = 8;

struct UpRow { float l[8], r[8]; };      // the left / right source chunks of one source row, unpacked

__device__ __forceinline__ void up_emit(const UpRow &top, const UpRow &bot, const float (&wx)[2][2], float sh, int Y0, int Y1,
                                        bf16 *__restrict__ ycol, size_t row_pitch, size_t col_step, bool two_cols) {
    // four corner weights per output (1 mul + 3 fma per value); same expressions as the general kernel
#pragma unroll
    for (int jy = 0; jy < 2; ++jy) {
        if (jy == 1 && Y1 == Y0) break;
        const int Y = jy == 0 ? Y0 : Y1;
        const float fy = sh * (float)Y, ly = fy - (float)(int)fy;
        const float wy0 = 1.0f - ly, wy1 = ly;
        bf16 *yr = ycol + (size_t)Y * row_pitch;
#pragma unroll
        for (int jx = 0; jx < 2; ++jx) {
            if (jx == 1 && !two_cols) break;
            const float w00 = wy0 * wx[jx][0], w01 = wy0 * wx[jx][1], w10 = wy1 * wx[jx][0], w11 = wy1 * wx[jx][1];
            float r[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = fmaf(bot.r[k], w11, fmaf(bot.l[k], w10, fmaf(top.r[k], w01, top.l[k] * w00)));
            __stcs(reinterpret_cast<uint4 *>(yr + (size_t)jx * col_step), pack8(r));
        }
    }
}

__global__ void __launch_bounds__(256, 3)      // 80 registers: 3 CTAs per SM measured best (2: 0.270, 3: 0.248, 4: 0.268 ms per step)
upsample2x_kernel(const bf16 *__restrict__ x, bf16 *__restrict__ y, unsigned N, int h, int w, int C) {
    soccdpt::pdl_wait();
    const int H = 2 * h, W = 2 * w;
    const unsigned chunks = (unsigned)C / 8u, gw = (unsigned)w + 1u, gh = (unsigned)h + 1u;
    const unsigned segs = (gh + kUpSeg - 1) / kUpSeg;
    const float sh = (float)(h - 1) / (float)(H - 1), sw = (float)(w - 1) / (float)(W - 1);
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i >= gw * chunks) return;
    const unsigned gx = i / chunks, ck = i - gx * chunks;
    const int X0 = gx == 0 ? 0 : 2 * (int)gx - 1, X1 = min(2 * (int)gx, W - 1);
    const int x0 = gx == 0 ? 0 : (int)gx - 1, x1 = min(x0 + 1, w - 1);
    float wx[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float fx = sw * (float)(j == 0 ? X0 : X1), lx = fx - (float)(int)fx;
        wx[j][0] = 1.0f - lx; wx[j][1] = lx;
    }
    const size_t row_pitch = (size_t)W * C, col_step = (size_t)(X1 - X0) * C;
    for (unsigned job = blockIdx.y; job < N * segs; job += gridDim.y) {
        const unsigned n = job / segs, sg = job - n * segs;
        const int g0 = (int)(sg * kUpSeg), g1 = min(g0 + kUpSeg, (int)gh);
        const bf16 *xl = x + (size_t)n * h * w * C + (size_t)x0 * C + ck * 8u, *xr = x + (size_t)n * h * w * C + (size_t)x1 * C + ck * 8u;
        bf16 *ycol = y + (size_t)n * H * row_pitch + (size_t)X0 * C + ck * 8u;
        const size_t src_pitch = (size_t)w * C;
        auto src_row = [&](int gy, bool bottom) { const int y0 = gy == 0 ? 0 : gy - 1; return bottom ? min(y0 + 1, h - 1) : y0; };
        UpRow P, Q;
        {
            const int rt = src_row(g0, false), rb = src_row(g0, true);
            unpack8(__ldg(reinterpret_cast<const uint4 *>(xl + rt * src_pitch)), P.l);
            unpack8(__ldg(reinterpret_cast<const uint4 *>(xr + rt * src_pitch)), P.r);
            unpack8(__ldg(reinterpret_cast<const uint4 *>(xl + rb * src_pitch)), Q.l);
            unpack8(__ldg(reinterpret_cast<const uint4 *>(xr + rb * src_pitch)), Q.r);
        }
        // group gy uses (top, bottom) = source rows (gy - 1, min(gy, h - 1)) for gy >= 1, (0, 1) for gy = 0: after group 0 the
        // top row stays, after any other group the bottom row becomes the top row and one new row is fetched
        int gy = g0;
        bool p_is_top = true;
        while (gy < g1) {
            const int nb = src_row(gy + 1, true);                       // bottom row of the next group
            const bool more = gy + 1 < g1, keep_top = gy == 0;
            uint4 rl = make_uint4(0, 0, 0, 0), rr = rl;
            if (more) {                                                 // in flight under this group's arithmetic
                rl = __ldg(reinterpret_cast<const uint4 *>(xl + nb * src_pitch));
                rr = __ldg(reinterpret_cast<const uint4 *>(xr + nb * src_pitch));
            }
            const int Y0 = gy == 0 ? 0 : 2 * gy - 1, Y1 = min(2 * gy, H - 1);
            if (p_is_top) up_emit(P, Q, wx, sh, Y0, Y1, ycol, row_pitch, col_step, X1 != X0);
            else up_emit(Q, P, wx, sh, Y0, Y1, ycol, row_pitch, col_step, X1 != X0);
            if (more) {
                // keep_top: the new row replaces the bottom; otherwise it replaces the old top, and the roles swap
                UpRow &dst = (p_is_top != keep_top) ? P : Q;
                unpack8(rl, dst.l);
                unpack8(rr, dst.r);
                if (!keep_top) p_is_top = !p_is_top;
            }
            ++gy;
        }
    }
}

// ------------------------------------------------------------------ seg head tail
// logits f32 [N,h,w,P] -> bilinear x2 (align_corners=True) -> sigmoid | 0.5*tanh+0.5 -> f32 NCHW
// One thread per 4 consecutive output pixels of a row (W = 2w is a multiple of 4 whenever w is even; odd w falls back to the
// scalar tail): their bilinear taps fall into at most 4 consecutive source columns, which are loaded once per source row and
// channel, and every class plane gets one float4 store.
__global__ void __launch_bounds__(256)
seg_finish_kernel(const float *__restrict__ lg, float *__restrict__ seg, int N, int h, int w, int P, int act) {
    soccdpt::pdl_wait();
    const int H = 2 * h, W = 2 * w;
    const int W4 = W / 4;
    const long long total = (long long)N * H * W4;
    const float sh = (float)(h - 1) / (float)(H - 1), sw = (float)(w - 1) / (float)(W - 1);
    for (long long o = (long long)blockIdx.x * 256 + threadIdx.x; o < total; o += (long long)gridDim.x * 256) {
        const int X4 = (int)(o % W4) * 4, Y = (int)((o / W4) % H), n = (int)(o / ((long long)W4 * H));
        const float fy = sh * (float)Y;
        const int y0 = (int)fy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
        const float ly = fy - (float)y0;
        int xo[4];
        float lx[4];
        const int xb = (int)(sw * (float)X4);                       // first source column of the group
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float fx = sw * (float)(X4 + i);
            const int x0 = (int)fx;
            xo[i] = x0 - xb;                                        // 0..2 (x2 up-scaling: 4 outputs span < 2 source columns)
            lx[i] = fx - (float)x0;
        }
        const float *r0 = lg + (((long long)n * h + y0) * w) * P, *r1 = lg + (((long long)n * h + y1) * w) * P;
        for (int p = 0; p < P; ++p) {
            float t[4];                                             // vertical lerp of source columns xb .. xb+3 (clamped)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xs = min(xb + j, w - 1);
                const float a = __ldg(r0 + (long long)xs * P + p), c = __ldg(r1 + (long long)xs * P + p);
                t[j] = (1.0f - ly) * a + ly * c;
            }
            float r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float a = xo[i] == 0 ? t[0] : (xo[i] == 1 ? t[1] : t[2]);
                const float b = xo[i] == 0 ? t[1] : (xo[i] == 1 ? t[2] : t[3]);
                // same association as (1-ly)*((1-lx)*a + lx*b) + ly*((1-lx)*c + lx*d) up to fp32 rounding
                const float v = (1.0f - lx[i]) * a + lx[i] * b;
                r[i] = act == 0 ? 1.0f / (1.0f + expf(-v)) : 0.5f * tanhf(v) + 0.5f;
            }
            *reinterpret_cast<float4 *>(seg + (((long long)n * P + p) * H + Y) * W + X4) = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float *__restrict__ x, bf16 *__restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void __launch_bounds__(256) bf16_to_f32_kernel(const bf16 *__restrict__ x, float *__restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        y[i] = __bfloat162float(x[i]);
}

int grid_for(long long work_items) {
    long long blocks = (work_items + 255) / 256;
    const long long cap = (long long)soccdpt::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

namespace soccdpt {
int validate_conv(const soccdpt_conv_t *c) {
    SOCCDPT_REQUIRE(c != nullptr, "conv descriptor is NULL");
    SOCCDPT_REQUIRE(c->x && c->wgt, "conv: x / wgt is NULL");
    SOCCDPT_REQUIRE(c->N >= 1 && c->H >= 1 && c->W >= 1, "conv: bad image dims %dx%dx%d", c->N, c->H, c->W);
    SOCCDPT_REQUIRE(c->Cin >= 8 && c->Cin % 8 == 0, "conv: Cin must be a multiple of 8 (got %d)", c->Cin);
    SOCCDPT_REQUIRE(c->Cout >= 8 && c->Cout % 8 == 0, "conv: Cout must be a multiple of 8 (got %d)", c->Cout);
    SOCCDPT_REQUIRE((c->KH == 1 || c->KH == 3 || (c->KH == 2 && c->stride == 2 && c->pad_trim == 1)) && c->KW == c->KH,
                    "conv: kernel must be 1x1, 3x3, or 2x2 with stride 2 and no padding (patch merging)");
    SOCCDPT_REQUIRE(c->act >= 0 && c->act <= 2, "conv: bad activation %d", c->act);
    SOCCDPT_REQUIRE(c->stride >= 0 && c->stride <= 2, "conv: stride must be 1 or 2 (got %d)", c->stride);
    SOCCDPT_REQUIRE(c->pad_trim >= 0 && c->pad_trim <= c->KH / 2, "conv: pad_trim must be in [0, KH/2] (got %d)", c->pad_trim);
    SOCCDPT_REQUIRE(c->proj_n >= 0 && c->proj_n <= 4, "conv: proj_n must be in [0,4]");
    if (c->proj_n > 0) {
        SOCCDPT_REQUIRE(c->proj_w && c->proj_b && c->proj_out, "conv: projection pointers are NULL");
        SOCCDPT_REQUIRE(c->Cout <= 256, "conv: fused projection needs Cout <= 256");
    }
    SOCCDPT_REQUIRE(c->y || c->y_relu || c->proj_n > 0, "conv: no output requested");
    if (c->qk_heads != 0) {
        SOCCDPT_REQUIRE(c->qk_heads > 0 && c->qk_scale && c->Cout == 96 * c->qk_heads,
                        "conv: the cosine-attention epilogue needs qk_scale and Cout == 3 * 32 * qk_heads (Cout=%d, qk_heads=%d)", c->Cout, c->qk_heads);
        SOCCDPT_REQUIRE(c->y && !c->y_relu && !c->res1 && !c->res2 && c->proj_n == 0 && c->act == SOCCDPT_ACT_NONE,
                        "conv: the cosine-attention epilogue writes y only (no activation, residual, ReLU copy or projection)");
    }
    if (c->up_src) {
        SOCCDPT_REQUIRE(c->stride <= 1 && c->proj_n == 0 && c->qk_heads == 0 && c->Cout % 32 == 0,
                        "conv: the up-sampled residual needs stride 1, Cout %% 32 == 0 and no projection (Cout=%d)", c->Cout);
        SOCCDPT_REQUIRE(c->up_h >= 1 && c->up_w >= 1 && c->H == 2 * c->up_h && c->W == 2 * c->up_w,
                        "conv: the up-sampled residual is an exact x2 (%dx%d -> %dx%d)", c->up_h, c->up_w, c->H, c->W);
    }
    return SOCCDPT_OK;
}
}  // namespace soccdpt

extern "C" {

int soccdpt_conv_ref_fwd(const soccdpt_conv_t *c, soccdpt_stream_t stream) {
    int rc = soccdpt::validate_conv(c);
    if (rc) return rc;
    const int st = c->stride > 1 ? c->stride : 1;
    const long long pix = (long long)c->N * ((c->H + st - 1) / st) * ((c->W + st - 1) / st);
    conv_ref_kernel<<<(unsigned)((pix + 7) / 8), 256, 0, soccdpt::as_stream(stream)>>>(*c);
    return soccdpt::check_launch("conv_ref_kernel");
}

int soccdpt_patch_embed_fwd(const float *x, const float *w, const float *b, const float *ln_w, const float *ln_b,
                            void *tokens, float *tokens_f32, int batch, int H, int W, int E, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(x && w && b && ln_w && ln_b && tokens, "patch_embed: NULL pointer");
    SOCCDPT_REQUIRE(batch >= 1 && H % 4 == 0 && W % 16 == 0 && E >= 32 && E <= 128, "patch_embed: need W %% 16 == 0 and 32 <= E <= 128");
    const long long groups = (long long)batch * (H / 4) * (W / 16);
    SOCCDPT_REQUIRE(groups < (1ll << 31), "patch_embed: batch too large for one call");
    static const bool fp32_path = getenv("SOCCDPT_PATCH_EMBED_FP32") && getenv("SOCCDPT_PATCH_EMBED_FP32")[0] == '1';
    if ((E == 96 || E == 128) && W % 64 == 0 && !fp32_path) {
        const long long tiles = (long long)batch * (H / 4) * (W / 64);
        long long nb = (tiles + 7) / 8;
        const long long capm = (long long)soccdpt::sm_count() * 2 * 4;
        if (nb > capm) nb = capm;
#define SOCC_PEM(NT)                                                                                                        \
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, patch_embed_mma_kernel<NT>, dim3((unsigned)nb), dim3(256), 0, \
                                     soccdpt::as_stream(stream), x, w, b, ln_w, ln_b, static_cast<bf16 *>(tokens), tokens_f32, \
                                     batch, H, W))
        if (E == 96) SOCC_PEM(12);
        else SOCC_PEM(16);
#undef SOCC_PEM
        return soccdpt::check_launch("patch_embed_mma_kernel");
    }
    long long blocks = (groups + 7) / 8;
    const long long cap = (long long)soccdpt::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)(48 * E + 8 * PE_TOK * 48) * sizeof(float);
#define SOCC_PE(NCH)                                                                                                     \
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, patch_embed_kernel<NCH>, dim3((unsigned)blocks), dim3(256), smem, \
                                     soccdpt::as_stream(stream), x, w, b, ln_w, ln_b, static_cast<bf16 *>(tokens), tokens_f32, \
                                     batch, H, W, E))
    if (E <= 32) SOCC_PE(1);
    else if (E <= 64) SOCC_PE(2);
    else if (E <= 96) SOCC_PE(3);
    else SOCC_PE(4);
#undef SOCC_PE
    return soccdpt::check_launch("patch_embed_kernel");
}

int soccdpt_layernorm_fwd(const void *t, const void *res, const float *gamma, const float *beta, void *y,
                          long long rows, int C, float eps, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(t && gamma && beta && y, "layernorm: NULL pointer");
    SOCCDPT_REQUIRE(rows >= 1 && C >= 8 && C % 8 == 0, "layernorm: C must be a multiple of 8 (got %d)", C);
    return launch_layernorm(static_cast<const bf16 *>(t), static_cast<const bf16 *>(res), nullptr, 0, gamma, beta,
                            static_cast<bf16 *>(y), rows, C, eps, soccdpt::as_stream(stream));
}

int soccdpt_layernorm_master_fwd(const void *t, float *master, int accumulate, const float *gamma, const float *beta,
                                 void *y, long long rows, int C, float eps, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(t && master && gamma && beta && y, "layernorm_master: NULL pointer");
    SOCCDPT_REQUIRE(rows >= 1 && C >= 8 && C % 8 == 0, "layernorm_master: C must be a multiple of 8 (got %d)", C);
    return launch_layernorm(static_cast<const bf16 *>(t), nullptr, master, accumulate, gamma, beta, static_cast<bf16 *>(y),
                            rows, C, eps, soccdpt::as_stream(stream));
}

int soccdpt_patch_merge_gather_fwd(const void *x, void *y, int batch, int H, int W, int C, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(x && y && batch >= 1 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "patch_merge: bad arguments");
    const long long items = (long long)batch * (H / 2) * (W / 2) * 4 * (C / 8);
    patch_merge_gather_kernel<<<grid_for(items), 256, 0, soccdpt::as_stream(stream)>>>(
        static_cast<const bf16 *>(x), static_cast<bf16 *>(y), batch, H, W, C);
    return soccdpt::check_launch("patch_merge_gather_kernel");
}

int soccdpt_upsample_bilinear_fwd(const void *x, void *y, int N, int h, int w, int H, int W, int C,
                                  soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(x && y && N >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1 && C % 8 == 0, "upsample: bad arguments");
    SOCCDPT_REQUIRE((long long)N * H < (1ll << 31) && (long long)W * C < (1ll << 31), "upsample: tensor too large");
    if (H == 2 * h && W == 2 * w && h >= 2 && w >= 2) {
        const unsigned per_group = (unsigned)(((long long)(w + 1) * (C / 8) + 255) / 256);
        const unsigned jobs = (unsigned)N * (unsigned)((h + 1 + kUpSeg - 1) / kUpSeg);
        dim3 grid(per_group, jobs < 65535u ? jobs : 65535u);
        SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, upsample2x_kernel, grid, dim3(256), 0, soccdpt::as_stream(stream), static_cast<const bf16 *>(x),
                                         static_cast<bf16 *>(y), (unsigned)N, h, w, C));
        return soccdpt::check_launch("upsample2x_kernel");
    }
    const unsigned per_row = (unsigned)(((long long)W * (C / 8) + 255) / 256), rows = (unsigned)N * (unsigned)H;
    const unsigned gx = (per_row + 3u) / 4u;            // ~4 items per thread: more loads in flight, setup amortised
    dim3 grid(gx < 64u ? gx : 64u, rows < 65535u ? rows : 65535u);
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, upsample_bilinear_kernel, grid, dim3(256), 0, soccdpt::as_stream(stream),
                                     static_cast<const bf16 *>(x), static_cast<bf16 *>(y), rows, h, w, H, W, C));
    return soccdpt::check_launch("upsample_bilinear_kernel");
}

int soccdpt_seg_finish_fwd(const float *logits, float *seg, int N, int h, int w, int P, int act,
                           soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(logits && seg && N >= 1 && h >= 2 && w >= 2 && P >= 1 && P <= 4 && (act == 0 || act == 1),
                    "seg_finish: bad arguments");
    SOCCDPT_REQUIRE(w % 2 == 0 && (reinterpret_cast<uintptr_t>(seg) & 15) == 0, "seg_finish: w must be even and seg 16-byte aligned");
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, seg_finish_kernel, dim3(grid_for((long long)N * h * w)), dim3(256), 0, soccdpt::as_stream(stream),
                                     logits, seg, N, h, w, P, act));
    return soccdpt::check_launch("seg_finish_kernel");
}

int soccdpt_f32_to_bf16(const float *x, void *y, long long n, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(x && y && n >= 1, "f32_to_bf16: bad arguments");
    f32_to_bf16_kernel<<<grid_for(n), 256, 0, soccdpt::as_stream(stream)>>>(x, static_cast<bf16 *>(y), n);
    return soccdpt::check_launch("f32_to_bf16_kernel");
}
int soccdpt_bf16_to_f32(const void *x, float *y, long long n, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(x && y && n >= 1, "bf16_to_f32: bad arguments");
    bf16_to_f32_kernel<<<grid_for(n), 256, 0, soccdpt::as_stream(stream)>>>(static_cast<const bf16 *>(x), y, n);
    return soccdpt::check_launch("bf16_to_f32_kernel");
}
}
