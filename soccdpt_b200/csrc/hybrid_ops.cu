// Kernels that only the ViT-hybrid encoder (dpt_hybrid_384: timm vit_base_resnet50_384) needs:
//   stem conv 7x7 stride 2 with TF-"SAME" padding (StdConv2dSame, weights standardised at pack time),
//   GroupNorm(32) [+ shortcut add] [+ ReLU] on NHWC bf16, MaxPool2dSame 3x3 stride 2, ViT token assembly
//   (cls token + position embedding), ProjectReadout concat, and global multi-head attention (577 tokens, d = 64).
// Reference call sites: SOccDPT/model/backbones/vit.py:44-85 (forward_flex), :147-242 (post-processing),
// SOccDPT/model/backbones/utils.py:27-40 (ProjectReadout); the arithmetic itself is timm 0.6.12's.
#include "common.cuh"

namespace soccdpt {
int global_attention_tc_max_tokens();
int launch_global_attention_tc(const void *qkv, void *out, int batch, int N, int heads, cudaStream_t st);
}  // namespace soccdpt

namespace {

using bf16 = __nv_bfloat16;

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

// ------------------------------------------------------------------ stem: conv 7x7 s2, 3 -> 64, SAME padding
// x f32 NCHW [B,3,H,W] -> y bf16 NHWC [B,H/2,W/2,64].  One CTA = 64 consecutive output pixels of one output row x 64
// channels; the 3 x 7 x 133 input patch and the standardised weights ([147][64]) live in shared memory, every thread
// owns 4 pixels x 4 channels (4 patch loads + one float4 weight load per 16 FMAs).
constexpr int ST_PX = 64, ST_COLS = ST_PX * 2 + 5;
__global__ void __launch_bounds__(256)
stem_conv7_kernel(const float *__restrict__ x, const float *__restrict__ w, bf16 *__restrict__ y, int B, int H, int W) {
    __shared__ __align__(16) float sw[147 * 64];
    __shared__ float patch[21][ST_COLS + 1];
    for (int i = threadIdx.x; i < 147 * 64; i += 256) {
        const int co = i / 147, k = i - co * 147;       // w is [64][3][7][7]
        sw[k * 64 + co] = w[i];
    }
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const int pad_h = max((Ho - 1) * 2 + 7 - H, 0) / 2, pad_w = max((Wo - 1) * 2 + 7 - W, 0) / 2;   // SAME: floor(total/2) first
    const int segs = (Wo + ST_PX - 1) / ST_PX;
    const int seg = blockIdx.x % segs, oy = (blockIdx.x / segs) % Ho, n = blockIdx.x / (segs * Ho);
    const int ox0 = seg * ST_PX;
    for (int i = threadIdx.x; i < 21 * ST_COLS; i += 256) {
        const int r = i / ST_COLS, c = i - r * ST_COLS;     // r = ci * 7 + ky
        const int ci = r / 7, ky = r - ci * 7;
        const int iy = oy * 2 + ky - pad_h, ix = ox0 * 2 + c - pad_w;
        patch[r][c] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[(((long long)n * 3 + ci) * H + iy) * W + ix] : 0.0f;
    }
    __syncthreads();
    const int pg = threadIdx.x >> 4, cg = threadIdx.x & 15;  // pixels pg*4 .. pg*4+3, channels cg*4 .. cg*4+3
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 1
    for (int r = 0; r < 21; ++r) {
        const float *pr = &patch[r][pg * 8];
        const float *wr = sw + r * 7 * 64 + cg * 4;
        float in[13];
#pragma unroll
        for (int c = 0; c < 13; ++c) in[c] = pr[c];          // 4 pixels x stride 2 + 7 taps - 2
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
            const float4 wv = *reinterpret_cast<const float4 *>(wr + kx * 64);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = in[2 * i + kx];
                acc[i][0] = fmaf(v, wv.x, acc[i][0]);
                acc[i][1] = fmaf(v, wv.y, acc[i][1]);
                acc[i][2] = fmaf(v, wv.z, acc[i][2]);
                acc[i][3] = fmaf(v, wv.w, acc[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ox = ox0 + pg * 4 + i;
        if (ox < Wo) {
            uint2 o;
            __nv_bfloat162 lo = __floats2bfloat162_rn(acc[i][0], acc[i][1]), hi = __floats2bfloat162_rn(acc[i][2], acc[i][3]);
            o.x = *reinterpret_cast<uint32_t *>(&lo);
            o.y = *reinterpret_cast<uint32_t *>(&hi);
            *reinterpret_cast<uint2 *>(y + (((long long)n * Ho + oy) * Wo + ox) * 64 + cg * 4) = o;
        }
    }
}

// ------------------------------------------------------------------ GroupNorm(32 groups), NHWC bf16
// pass 1: per-(image, group) sum / sum of squares.  C / 8 divides 256 for every width of the model, so a thread meets the
// same 8 channels on every step of its stride-256 loop and keeps their (at most 4) group sums in registers; one
// shared-memory atomic per thread and group at the end, then one double atomicAdd per (block, group, statistic).
// pass 2: normalise (+ shortcut) (+ ReLU); mean / rstd are derived once per 16-byte chunk, not per element.
template <int SUB>   // channels of one group inside a 16-byte chunk: 8 (cpg >= 8), 4 or 2
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const bf16 *__restrict__ x, double *__restrict__ stats, int HW, int C, int slabs) {
    __shared__ float s_sum[32], s_sq[32];
    if (threadIdx.x < 32) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
    soccdpt::pdl_wait();
    __syncthreads();
    const int n = blockIdx.x / slabs, slab = blockIdx.x % slabs;
    const int chunks = C / 8, cpg = C / 32;               // channels per group (>= 2)
    const long long items = (long long)HW * chunks;
    long long per = (items + slabs - 1) / slabs;
    per = (per + 255) / 256 * 256;                         // slab starts stay multiples of 256 (and so of `chunks`)
    const long long i0 = slab * per, i1 = min(items, i0 + per);
    const bf16 *xn = x + (long long)n * HW * C;
    constexpr int NG = 8 / SUB;
    if (256 % chunks == 0) {
        const int ck = threadIdx.x % chunks;
        float s[NG], q[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) s[k] = q[k] = 0.f;
        // pointer walk: every step advances 256 / chunks pixels
        const bf16 *xp = xn + ((i0 + threadIdx.x) / chunks) * C + ck * 8;
        const long long step = (long long)(256 / chunks) * C;
        const int iters = i1 > i0 + threadIdx.x ? (int)((i1 - i0 - threadIdx.x + 255) / 256) : 0;
#pragma unroll 4
        for (int it = 0; it < iters; ++it, xp += step) {
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4 *>(xp)), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { s[k / SUB] += f[k]; q[k / SUB] = fmaf(f[k], f[k], q[k / SUB]); }
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            atomicAdd(&s_sum[(ck * 8 + k * SUB) / cpg], s[k]);
            atomicAdd(&s_sq[(ck * 8 + k * SUB) / cpg], q[k]);
        }
    } else {                                               // generic widths: per-chunk shared-memory atomics
        for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
            const int ck = (int)(i % chunks);
            float f[8];
            unpack8(*reinterpret_cast<const uint4 *>(xn + (i / chunks) * C + ck * 8), f);
#pragma unroll
            for (int k0 = 0; k0 < 8; k0 += SUB) {
                float s = 0.f, q = 0.f;
#pragma unroll
                for (int k = k0; k < k0 + SUB; ++k) { s += f[k]; q = fmaf(f[k], f[k], q); }
                atomicAdd(&s_sum[(ck * 8 + k0) / cpg], s);
                atomicAdd(&s_sq[(ck * 8 + k0) / cpg], q);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        atomicAdd(&stats[((long long)n * 32 + threadIdx.x) * 2 + 0], (double)s_sum[threadIdx.x]);
        atomicAdd(&stats[((long long)n * 32 + threadIdx.x) * 2 + 1], (double)s_sq[threadIdx.x]);
    }
}

template <int SUB>
__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(const bf16 *__restrict__ x, const double *__restrict__ stats, const float *__restrict__ gamma,
                       const float *__restrict__ beta, const bf16 *__restrict__ shortcut, bf16 *__restrict__ y, long long total_chunks,
                       int HW, int C, float eps, int relu) {
    soccdpt::pdl_wait();
    const int chunks = C / 8, cpg = C / 32;
    const double inv_cnt = 1.0 / ((double)HW * cpg);
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < (unsigned)total_chunks; i += gridDim.x * 256u) {   // 32-bit index math
        const unsigned upix = i / (unsigned)chunks;
        const int ck = (int)(i - upix * (unsigned)chunks);
        const long long pix = upix;
        const int n = (int)(upix / (unsigned)HW);
        float f[8], r[8];
        unpack8(__ldg(reinterpret_cast<const uint4 *>(x + pix * C + ck * 8)), f);
        if (shortcut) unpack8(__ldg(reinterpret_cast<const uint4 *>(shortcut + pix * C + ck * 8)), r);
        const float4 g0 = *reinterpret_cast<const float4 *>(gamma + ck * 8), g1 = *reinterpret_cast<const float4 *>(gamma + ck * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4 *>(beta + ck * 8), b1 = *reinterpret_cast<const float4 *>(beta + ck * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += SUB) {
            const double2 st = *reinterpret_cast<const double2 *>(stats + ((long long)n * 32 + (ck * 8 + k0) / cpg) * 2);
            const double md = st.x * inv_cnt;
            const float m = (float)md, rstd = rsqrtf(fmaxf((float)(st.y * inv_cnt - md * md), 0.0f) + eps);
#pragma unroll
            for (int k = k0; k < k0 + SUB; ++k) {
                float v = (f[k] - m) * rstd * gg[k] + bb[k];
                if (shortcut) v += r[k];
                f[k] = relu ? fmaxf(v, 0.0f) : v;
            }
        }
        *reinterpret_cast<uint4 *>(y + pix * C + ck * 8) = pack8(f);
    }
}

// ------------------------------------------------------------------ MaxPool2dSame 3x3 s2, NHWC bf16 (pad value -inf)
__global__ void __launch_bounds__(256)
maxpool3_s2_kernel(const bf16 *__restrict__ x, bf16 *__restrict__ y, int B, int H, int W, int C) {
    soccdpt::pdl_wait();
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, chunks = C / 8;
    const int pad_h = max((Ho - 1) * 2 + 3 - H, 0) / 2, pad_w = max((Wo - 1) * 2 + 3 - W, 0) / 2;
    const long long total = (long long)B * Ho * Wo * chunks;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int ck = (int)(i % chunks);
        const long long o = i / chunks;
        const int ox = (int)(o % Wo), oy = (int)((o / Wo) % Ho), n = (int)(o / ((long long)Wo * Ho));
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * 2 + ky - pad_h;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = ox * 2 + kx - pad_w;
                if (ix < 0 || ix >= W) continue;
                float f[8];
                unpack8(*reinterpret_cast<const uint4 *>(x + (((long long)n * H + iy) * W + ix) * C + ck * 8), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
            }
        }
        *reinterpret_cast<uint4 *>(y + o * C + ck * 8) = pack8(m);
    }
}

// ------------------------------------------------------------------ ViT tokens: [cls | patch tokens] + pos_embed
// patches bf16 [B,L,D] (patch_embed.proj output), cls f32 [D], pos f32 [1+L][D] -> tokens bf16 [B,1+L,D]
__global__ void __launch_bounds__(256)
vit_tokens_kernel(const bf16 *__restrict__ patches, const float *__restrict__ cls, const float *__restrict__ pos,
                  bf16 *__restrict__ tokens, float *__restrict__ tokens_f32, int B, int L, int D) {
    soccdpt::pdl_wait();
    const long long total = (long long)B * (L + 1) * D;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int d = (int)(i % D);
        const int t = (int)((i / D) % (L + 1));
        const int b = (int)(i / ((long long)D * (L + 1)));
        const float v = t == 0 ? cls[d] : __bfloat162float(patches[((long long)b * L + t - 1) * D + d]);
        const float o = v + pos[(long long)t * D + d];
        tokens[i] = __float2bfloat16_rn(o);
        if (tokens_f32) tokens_f32[i] = o;
    }
}

// ------------------------------------------------------------------ ProjectReadout input: cat(tok[1:], cls.expand)
// tokens bf16 [B,1+L,D] -> feats bf16 [B,L,2D]
__global__ void __launch_bounds__(256)
readout_concat_kernel(const bf16 *__restrict__ tokens, bf16 *__restrict__ feats, int B, int L, int D) {
    soccdpt::pdl_wait();
    const int chunks = D / 8;
    const long long total = (long long)B * L * 2 * chunks;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int ck = (int)(i % (2 * chunks));
        const long long row = i / (2 * chunks);
        const int l = (int)(row % L), b = (int)(row / L);
        const bf16 *src = ck < chunks ? tokens + ((long long)b * (L + 1) + 1 + l) * D + ck * 8
                                      : tokens + ((long long)b * (L + 1)) * D + (ck - chunks) * 8;
        *reinterpret_cast<uint4 *>(feats + row * 2 * D + ck * 8) = *reinterpret_cast<const uint4 *>(src);
    }
}

// ------------------------------------------------------------------ global multi-head attention, head dim 64
// Fallback for N > 640 tokens (the tcgen05 kernel in global_attention_tc.cu covers the model's 577).
// qkv bf16 [B,N,3*H*64] (q|k|v), out bf16 [B,N,H*64]; softmax(q k^T / 8) v.  One CTA per (image, head, 128 queries);
// K and V of the head are staged in shared memory as bf16; one thread per query, chunked online softmax in fp32.
constexpr int GD = 64, GCH = 8;
__global__ void __launch_bounds__(128)
global_attention_kernel(const bf16 *__restrict__ qkv, bf16 *__restrict__ out, int N, int heads, float scale) {
    extern __shared__ __align__(16) uint8_t ga_smem[];
    bf16 *Ks = reinterpret_cast<bf16 *>(ga_smem);            // [Npad][64]
    const int Npad = (N + GCH - 1) / GCH * GCH;
    bf16 *Vs = Ks + (size_t)Npad * GD;
    const int b = blockIdx.x / heads, head = blockIdx.x % heads;
    const int C = heads * GD;
    const bf16 *base = qkv + (long long)b * N * 3 * C + head * GD;
    for (int i = threadIdx.x; i < Npad * (GD / 8); i += 128) {
        const int r = i / (GD / 8), ck = i % (GD / 8);
        uint4 k4 = make_uint4(0, 0, 0, 0), v4 = make_uint4(0, 0, 0, 0);
        if (r < N) {
            k4 = *reinterpret_cast<const uint4 *>(base + (long long)r * 3 * C + C + ck * 8);
            v4 = *reinterpret_cast<const uint4 *>(base + (long long)r * 3 * C + 2 * C + ck * 8);
        }
        *reinterpret_cast<uint4 *>(Ks + r * GD + ck * 8) = k4;
        *reinterpret_cast<uint4 *>(Vs + r * GD + ck * 8) = v4;
    }
    __syncthreads();
    const int qi = blockIdx.y * 128 + threadIdx.x;
    if (qi >= N) return;
    float q[GD], acc[GD];
#pragma unroll
    for (int ck = 0; ck < GD / 8; ++ck) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4 *>(base + (long long)qi * 3 * C + ck * 8), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { q[ck * 8 + k] = f[k] * scale; acc[ck * 8 + k] = 0.f; }
    }
    float m = -INFINITY, l = 0.f;
    for (int j0 = 0; j0 < Npad; j0 += GCH) {
        float s[GCH];
        float cm = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < GCH; ++jj) {
            const uint4 *kp = reinterpret_cast<const uint4 *>(Ks + (j0 + jj) * GD);
            float dot = 0.f;
#pragma unroll
            for (int ck = 0; ck < GD / 8; ++ck) {
                float f[8];
                unpack8(kp[ck], f);
#pragma unroll
                for (int k = 0; k < 8; ++k) dot = fmaf(q[ck * 8 + k], f[k], dot);
            }
            s[jj] = (j0 + jj < N) ? dot : -INFINITY;           // padded keys never contribute
            cm = fmaxf(cm, s[jj]);
        }
        const float mn = fmaxf(m, cm);
        const float corr = __expf(m - mn);
        l *= corr;
#pragma unroll
        for (int d = 0; d < GD; ++d) acc[d] *= corr;
#pragma unroll
        for (int jj = 0; jj < GCH; ++jj) {
            const float p = __expf(s[jj] - mn);
            l += p;
            const uint4 *vp = reinterpret_cast<const uint4 *>(Vs + (j0 + jj) * GD);
#pragma unroll
            for (int ck = 0; ck < GD / 8; ++ck) {
                float f[8];
                unpack8(vp[ck], f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[ck * 8 + k] = fmaf(p, f[k], acc[ck * 8 + k]);
            }
        }
        m = mn;
    }
    const float inv = 1.0f / l;
    bf16 *op = out + ((long long)b * N + qi) * C + head * GD;
#pragma unroll
    for (int ck = 0; ck < GD / 8; ++ck) {
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = acc[ck * 8 + k] * inv;
        *reinterpret_cast<uint4 *>(op + ck * 8) = pack8(f);
    }
}

// ------------------------------------------------------------------ pre-norm residual step of a ViT block
// master (fp32 residual stream) += t (bf16 branch output, optional);  y = LayerNorm(master) (optional);
// stream_bf16 = bf16(master) (optional: the hooked block outputs).  One warp per row, the row stays in registers.
template <int ITERS>
__global__ void __launch_bounds__(256)
prenorm_kernel(const bf16 *__restrict__ t, float *__restrict__ master, const float *__restrict__ gamma,
               const float *__restrict__ beta, bf16 *__restrict__ y, bf16 *__restrict__ stream_bf16, long long rows, int C, float eps) {
    soccdpt::pdl_wait();        // before the early return: every thread of a PDL-launched grid waits
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;                                    // whole warps leave together
    const int chunks = C / 8;
    float4 *mp = reinterpret_cast<float4 *>(master + row * C);
    float f[ITERS][8];
    float s = 0.0f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int k = lane + it * 32;
        if (k < chunks) {
            const float4 a = mp[2 * k], b = mp[2 * k + 1];
            f[it][0] = a.x; f[it][1] = a.y; f[it][2] = a.z; f[it][3] = a.w;
            f[it][4] = b.x; f[it][5] = b.y; f[it][6] = b.z; f[it][7] = b.w;
            if (t) {
                float r[8];
                unpack8(*reinterpret_cast<const uint4 *>(t + row * C + k * 8), r);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[it][i] += r[i];
                mp[2 * k] = make_float4(f[it][0], f[it][1], f[it][2], f[it][3]);
                mp[2 * k + 1] = make_float4(f[it][4], f[it][5], f[it][6], f[it][7]);
            }
            if (stream_bf16) *reinterpret_cast<uint4 *>(stream_bf16 + row * C + k * 8) = pack8(f[it]);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += f[it][i];
        }
    }
    if (!y) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.0f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
        if (lane + it * 32 < chunks) {
#pragma unroll
            for (int i = 0; i < 8; ++i) q += (f[it][i] - mean) * (f[it][i] - mean);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int k = lane + it * 32;
        if (k < chunks) {
            float o8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o8[i] = (f[it][i] - mean) * rstd * gamma[k * 8 + i] + beta[k * 8 + i];
            *reinterpret_cast<uint4 *>(y + row * C + k * 8) = pack8(o8);
        }
    }
}

int grid_for(long long items) {
    long long blocks = (items + 255) / 256;
    const long long cap = (long long)soccdpt::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" {

int soccdpt_stem_conv7_fwd(const float *x, const float *w, void *y, int batch, int H, int W, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(x && w && y && batch >= 1 && H >= 7 && W >= 7, "stem_conv7: bad arguments");
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const long long blocks = (long long)batch * Ho * ((Wo + ST_PX - 1) / ST_PX);
    SOCCDPT_REQUIRE(blocks < (1ll << 31), "stem_conv7: too many tiles");
    stem_conv7_kernel<<<(unsigned)blocks, 256, 0, soccdpt::as_stream(stream)>>>(x, w, static_cast<bf16 *>(y), batch, H, W);
    return soccdpt::check_launch("stem_conv7_kernel");
}

int soccdpt_groupnorm_fwd(const void *x, const float *gamma, const float *beta, const void *shortcut, void *y, int batch,
                          int HW, int C, float eps, int relu, void *stats_scratch, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(x && gamma && beta && y && stats_scratch, "groupnorm: NULL pointer");
    SOCCDPT_REQUIRE(batch >= 1 && HW >= 1 && (C == 64 || C == 128 || (C >= 256 && C % 256 == 0)),
                    "groupnorm: 32 groups of 2, 4 or a multiple of 8 channels (C = 64, 128 or a multiple of 256), got %d", C);
    cudaStream_t st = soccdpt::as_stream(stream);
    double *stats = static_cast<double *>(stats_scratch);           // [batch][32][2]
    SOCCDPT_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * batch * 64, st));
    int slabs = (int)(((long long)HW * (C / 8) + 256 * 32 - 1) / (256 * 32));   // ~32 chunks per thread
    if (slabs < 1) slabs = 1;
    if (slabs > 512) slabs = 512;
    const bf16 *xp = static_cast<const bf16 *>(x), *sp = static_cast<const bf16 *>(shortcut);
    bf16 *yp = static_cast<bf16 *>(y);
    const long long total = (long long)batch * HW * (C / 8);
    SOCCDPT_REQUIRE(total < (1ll << 31), "groupnorm: tensor too large for one call (%lld chunks)", total);
    const unsigned g1 = (unsigned)(batch * slabs), g2 = (unsigned)grid_for(total);
#define SOCC_GN(SUB)                                                                                          \
    do {                                                                                                      \
        SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, groupnorm_stats_kernel<SUB>, dim3(g1), dim3(256), 0, st, xp, stats, \
                                         HW, C, slabs));                                                      \
        int rc = soccdpt::check_launch("groupnorm_stats_kernel");                                             \
        if (rc) return rc;                                                                                    \
        SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, groupnorm_apply_kernel<SUB>, dim3(g2), dim3(256), 0, st, xp, stats, \
                                         gamma, beta, sp, yp, total, HW, C, eps, relu));                      \
    } while (0)
    if (C == 64) SOCC_GN(2);
    else if (C == 128) SOCC_GN(4);
    else SOCC_GN(8);
#undef SOCC_GN
    return soccdpt::check_launch("groupnorm_apply_kernel");
}

int soccdpt_maxpool3s2_fwd(const void *x, void *y, int batch, int H, int W, int C, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(x && y && batch >= 1 && H >= 2 && W >= 2 && C % 8 == 0, "maxpool: bad arguments");
    const long long items = (long long)batch * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, maxpool3_s2_kernel, dim3(grid_for(items)), dim3(256), 0,
                                     soccdpt::as_stream(stream), static_cast<const bf16 *>(x), static_cast<bf16 *>(y), batch, H, W, C));
    return soccdpt::check_launch("maxpool3_s2_kernel");
}

int soccdpt_vit_tokens_fwd(const void *patches, const float *cls, const float *pos, void *tokens, float *tokens_f32, int batch,
                           int L, int D, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(patches && cls && pos && tokens && batch >= 1 && L >= 1 && D >= 8, "vit_tokens: bad arguments");
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, vit_tokens_kernel, dim3(grid_for((long long)batch * (L + 1) * D)), dim3(256),
                                     0, soccdpt::as_stream(stream), static_cast<const bf16 *>(patches), cls, pos,
                                     static_cast<bf16 *>(tokens), tokens_f32, batch, L, D));
    return soccdpt::check_launch("vit_tokens_kernel");
}

int soccdpt_readout_concat_fwd(const void *tokens, void *feats, int batch, int L, int D, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(tokens && feats && batch >= 1 && L >= 1 && D % 8 == 0, "readout_concat: bad arguments");
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, readout_concat_kernel, dim3(grid_for((long long)batch * L * 2 * (D / 8))),
                                     dim3(256), 0, soccdpt::as_stream(stream), static_cast<const bf16 *>(tokens),
                                     static_cast<bf16 *>(feats), batch, L, D));
    return soccdpt::check_launch("readout_concat_kernel");
}

int soccdpt_global_attention_fwd(const void *qkv, void *out, int batch, int N, int heads, int head_dim,
                                 soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(qkv && out && batch >= 1 && N >= 1 && heads >= 1, "global_attention: bad arguments");
    SOCCDPT_REQUIRE(head_dim == GD, "global_attention: head_dim must be 64 (got %d)", head_dim);
    if (N <= soccdpt::global_attention_tc_max_tokens())      // tcgen05 kernel (global_attention_tc.cu)
        return soccdpt::launch_global_attention_tc(qkv, out, batch, N, heads, soccdpt::as_stream(stream));
    // longer sequences: CUDA-core kernel with K / V of one head resident in shared memory
    const int Npad = (N + GCH - 1) / GCH * GCH;
    const size_t smem = (size_t)Npad * GD * 2 * sizeof(bf16);
    SOCCDPT_REQUIRE(smem <= 220 * 1024, "global_attention: %d tokens do not fit shared memory", N);
    static soccdpt::SmemAttr configured;
    if (configured.need(smem)) {
        SOCCDPT_CUDA(cudaFuncSetAttribute(global_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid((unsigned)(batch * heads), (unsigned)((N + 127) / 128));
    global_attention_kernel<<<grid, 128, smem, soccdpt::as_stream(stream)>>>(static_cast<const bf16 *>(qkv), static_cast<bf16 *>(out),
                                                                            N, heads, 1.0f / sqrtf((float)head_dim));
    return soccdpt::check_launch("global_attention_kernel");
}

int soccdpt_prenorm_fwd(const void *t, float *master, const float *gamma, const float *beta, void *y, void *stream_bf16,
                        long long rows, int C, float eps, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(master && rows >= 1 && C >= 8 && C % 8 == 0 && C <= 1024, "prenorm: bad arguments (C = %d)", C);
    SOCCDPT_REQUIRE(!y || (gamma && beta), "prenorm: LayerNorm output requested without gamma/beta");
    cudaStream_t st = soccdpt::as_stream(stream);
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    const bf16 *tp = static_cast<const bf16 *>(t);
    bf16 *yp = static_cast<bf16 *>(y), *sp = static_cast<bf16 *>(stream_bf16);
    if (C <= 256) SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, prenorm_kernel<1>, dim3(blocks), dim3(256), 0, st, tp, master, gamma, beta, yp, sp, rows, C, eps));
    else if (C <= 512) SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, prenorm_kernel<2>, dim3(blocks), dim3(256), 0, st, tp, master, gamma, beta, yp, sp, rows, C, eps));
    else if (C <= 768) SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, prenorm_kernel<3>, dim3(blocks), dim3(256), 0, st, tp, master, gamma, beta, yp, sp, rows, C, eps));
    else SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, prenorm_kernel<4>, dim3(blocks), dim3(256), 0, st, tp, master, gamma, beta, yp, sp, rows, C, eps));
    return soccdpt::check_launch("prenorm_kernel");
}
}
