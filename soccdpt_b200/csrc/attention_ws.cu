// SwinV2 window attention for 256-token windows, warp-specialised and persistent (round 2).
//
// Same arithmetic as attention_tc.cu (timm 0.6.12 WindowAttention + SwinTransformerBlock._attn: cosine attention, clamped
// logit scale, relative-position bias from the baked table, cyclic shift in the index math, {0,-100} shift mask), but the
// round-1 kernel ran one (window, head) per CTA as a SERIAL chain -- global loads, normalise / transpose into operand tiles,
// S = Q K^T, softmax, P V, store -- and its profile (profiles/r1g_voxeliser_attention.ncu-rep) put the softmax, the only
// phase that is bound by a pipe (MUFU), at 28 % of the samples; load latency, staging and CTA barriers took the rest.
//
// Here ONE persistent CTA per SM walks over (window, head) items with three kinds of warps:
//   warps 0..7   producers: each thread owns one token row of the window: loads its q / k / v head slices (64 B each),
//                L2-normalises q and k in fp32 (q also carries the clamped logit scale), writes the K-major SW64 operand
//                rows of Q and K and scatters V transposed (V^T, SW128), plus the head's bias table (log2 domain, shifted by
//                the analytic logit bound) and the shift-mask region ids -- into a 3-deep ring of 48 KB stages, so the
//                producers run up to two items ahead of the consumers and their global-load latency is never exposed.
//   warps 8, 17  one lane each issues the tcgen05.mma of one softmax stream: S = Q_h K^T for a 128-query half into the
//                stream's 256-column half of TMEM, then O = P V with the A operand in TMEM as soon as the stream has
//                written P back over the score columns it consumed (tcgen05.st).
//   warps 9..16  two softmax streams of four warps (thread = query row, all 256 keys), each taking every other item: one pass
//                against the analytic bound (exact row-max pre-pass only for huge logit scales), exp2 on the MUFU pipe, bf16
//                pairs back into TMEM, epilogue O / rowsum -> bf16 -> global.  A stream's MMA / epilogue bubble is covered by
//                the other stream's softmax (first version: all eight warps on one item, in lockstep: MUFU idle 40 %).
// Hand-offs are mbarriers only; there is no CTA-wide barrier in the loop.
#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using bf16 = __nv_bfloat16;
using namespace tc;

constexpr int D = 32;                      // head dim
constexpr int NTOK = 256;                  // tokens per window (16 x 16)
constexpr int WS = 16;
constexpr int STAGES = 3;
constexpr int PROD_WARPS = 4, SM_WARPS = 16;                   // softmax: 2 streams x 8 warps
constexpr int SM_WARP0 = PROD_WARPS + 2;                       // warps: 0..3 producers, 4 / 5 MMA issuers, 6..21 softmax
constexpr int THREADS = 32 * (PROD_WARPS + 2 + SM_WARPS);      // 704
constexpr int TAB = 31 * 31;               // relative-position table of one head
constexpr int TS = 48;                     // its row stride in shared memory ((TS - 16) % 32 == 0: conflict-free reads)
// stage layout (bytes)
constexpr int OFF_Q = 0;                   // 256 rows x 64 B, SWIZZLE_64B (rows 0..127: half 0, 128..255: half 1)
constexpr int OFF_K = 16384;               // 256 rows x 64 B, SWIZZLE_64B
constexpr int OFF_VT = 32768;              // 4 key blocks x (32 rows x 128 B), SWIZZLE_128B
constexpr int OFF_VRAW = 49152;            // 256 rows x 64 B as they arrive (cp.async), transposed into OFF_VT by the producers
constexpr int OFF_TAB = 65536;             // [31][TS] f32
constexpr int OFF_REG = OFF_TAB + 31 * TS * 4;   // [256] u8 region ids
constexpr int OFF_SCALE = OFF_REG + 256;   // f32 logit scale of the item's head (+ padding)
constexpr int STAGE_BYTES = (OFF_SCALE + 16 + 1023) / 1024 * 1024;    // 72704
constexpr int OFF_SUM = STAGES * STAGE_BYTES;                          // [2 streams][2 step parities][2 key halves][128] f32
constexpr int OFF_MAX = OFF_SUM + 2 * 2 * 2 * 128 * 4;                 // [2 streams][2 key halves][128] f32 (row-max fallback)
constexpr int OFF_BAR = OFF_MAX + 2 * 2 * 128 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int TM_O = 64;                   // O columns inside a stream's TMEM half; P (bf16 pairs) = columns 0..63 (keys 0..127)
                                           // and 128..191 (keys 128..255): score columns their writers have consumed

struct Bars {
    uint64_t full[STAGES], empty[STAGES];
    uint64_t s_full[2], p_full[2], o_full[2], tm_free[2];      // per softmax stream
    uint32_t tmem_slot;
};

struct Args {
    const bf16 *qkv;
    const float *bias_tab, *scale;
    bf16 *out;
    int Hs, Ws, C, heads, shift, items, nwx, nwy;
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void unpack8(const uint4 &u, float *f) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 t = __bfloat1622float2(h[k]);
        f[2 * k] = t.x;
        f[2 * k + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8_scaled(const float *f, float s) {
    uint4 u;
    u.x = pack_bf16x2(f[0] * s, f[1] * s);
    u.y = pack_bf16x2(f[2] * s, f[3] * s);
    u.z = pack_bf16x2(f[4] * s, f[5] * s);
    u.w = pack_bf16x2(f[6] * s, f[7] * s);
    return u;
}
__device__ __forceinline__ float sumsq(const uint4 (&raw)[4]) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float f[8];
        unpack8(raw[i], f);
#pragma unroll
        for (int d = 0; d < 8; ++d) t = fmaf(f[d], f[d], t);
    }
    return t;
}
__device__ __forceinline__ int region_of(int p, int size, int shift) {
    // timm: slices (0,-ws), (-ws,-shift), (-shift,None) over the SHIFTED image
    return p < size - WS ? 0 : (p < size - shift ? 1 : 2);
}
__device__ __forceinline__ void tmem_ld32w(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    tmem_ld32_nowait(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <bool MASK>
__global__ void __launch_bounds__(THREADS, 1)
window_attention_ws_kernel(const Args a) {
    extern __shared__ uint8_t ws_raw[];
    uint8_t *smem = ws_raw + ((1024u - (smem_u32(ws_raw) & 1023u)) & 1023u);
    Bars *bars = reinterpret_cast<Bars *>(smem + OFF_BAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_items = (int)blockIdx.x < a.items ? (a.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const float LOG2E = 1.4426950408889634f;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->full[s], 32 * PROD_WARPS);
            mbar_init(&bars->empty[s], 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&bars->s_full[g], 1);
            mbar_init(&bars->p_full[g], 256);
            mbar_init(&bars->o_full[g], 1);
            mbar_init(&bars->tm_free[g], 256);
        }
        fence_barrier_init();
    }
    if (warp == PROD_WARPS) tmem_alloc(&bars->tmem_slot, 512);
    float *s_sum = reinterpret_cast<float *>(smem + OFF_SUM);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_slot;

    // (window, head) item -> token index (in the un-shifted image) of window row r, and its shift-mask region
    auto decode = [&](int item, int &head, int &win, int &b) {
        head = item % a.heads;
        const int wb = item / a.heads;
        win = wb % (a.nwx * a.nwy);
        b = wb / (a.nwx * a.nwy);
    };
    auto token_of = [&](int win, int b, int r, int &region) -> long long {
        const int ty = r >> 4, tx = r & 15;
        const int ys = (win / a.nwx) * WS + ty, xs = (win % a.nwx) * WS + tx;
        const int yo = (ys + a.shift) % a.Hs, xo = (xs + a.shift) % a.Ws;
        region = MASK ? region_of(ys, a.Hs, a.shift) * 3 + region_of(xs, a.Ws, a.shift) : 0;
        return ((long long)b * a.Hs + yo) * a.Ws + xo;
    };

    if (warp < PROD_WARPS) {
        // ============================== producers: one token row per thread ==============================
        // cp.async brings the row's q / k / v head slices straight into the stage TWO items ahead (q and k at their final
        // swizzled operand position, v into a scratch row); the thread later normalises q / k in place and scatters v
        // transposed.  Nothing of a row is held in registers while it is in flight.  128 threads, two rows each.
        soccdpt::pdl_wait();                            // qkv is the previous kernel's output
        auto request = [&](int n) {                     // rows of item n -> stage n % STAGES (an empty group beyond the end)
            if (n < my_items) {
                const int stage = n % STAGES;
                mbar_wait(&bars->empty[stage], (uint32_t)((n / STAGES) & 1) ^ 1u);
                int head, win, b, region;
                decode((int)blockIdx.x + n * (int)gridDim.x, head, win, b);
                uint8_t *st = smem + stage * STAGE_BYTES;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = (int)threadIdx.x + rr * 128;          // window row (a key AND a query)
                    const int swz = (r >> 1) & 3;
                    const long long tok = token_of(win, b, r, region);
                    const bf16 *p = a.qkv + tok * 3 * a.C + head * D;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        cp_async16(st + OFF_Q + r * 64 + ((c4 ^ swz) << 4), p + c4 * 8);
                        cp_async16(st + OFF_K + r * 64 + ((c4 ^ swz) << 4), p + a.C + c4 * 8);
                        cp_async16(st + OFF_VRAW + r * 64 + (c4 << 4), p + 2 * a.C + c4 * 8);
                    }
                    st[OFF_REG + r] = (uint8_t)region;
                }
            }
            cp_async_commit();
        };
        request(0);
        request(1);
        for (int n = 0; n < my_items; ++n) {
            cp_async_wait<1>();                         // this thread's copies of item n have landed (it reads back only its own rows)
            const int stage = n % STAGES;
            uint8_t *st = smem + stage * STAGE_BYTES;
            int head, win, b;
            decode((int)blockIdx.x + n * (int)gridDim.x, head, win, b);
            const float sc = a.scale[head];
            const bool one_pass = 2.01f * sc + 16.0f < 80.0f;
            const float tab_shift = one_pass ? (1.01f * sc + 16.0f) * LOG2E : 0.0f;
            float tabv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = (int)threadIdx.x + i * 128;
                tabv[i] = e < TAB ? a.bias_tab[(size_t)head * TAB + e] : 0.f;
            }
#pragma unroll 1
            for (int rr = 0; rr < 2; ++rr) {
                const int r = (int)threadIdx.x + rr * 128;
                const int swz = (r >> 1) & 3;
                {   // K and Q: L2-normalised in place (q also carries the clamped logit scale)
                    uint4 raw[4];
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) raw[c4] = *reinterpret_cast<const uint4 *>(st + OFF_K + r * 64 + ((c4 ^ swz) << 4));
                    const float ks = 1.0f / fmaxf(sqrtf(sumsq(raw)), 1e-12f);      // F.normalize eps
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        float f[8];
                        unpack8(raw[c4], f);
                        *reinterpret_cast<uint4 *>(st + OFF_K + r * 64 + ((c4 ^ swz) << 4)) = pack8_scaled(f, ks);
                    }
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) raw[c4] = *reinterpret_cast<const uint4 *>(st + OFF_Q + r * 64 + ((c4 ^ swz) << 4));
                    const float qs = sc / fmaxf(sqrtf(sumsq(raw)), 1e-12f);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        float f[8];
                        unpack8(raw[c4], f);
                        *reinterpret_cast<uint4 *>(st + OFF_Q + r * 64 + ((c4 ^ swz) << 4)) = pack8_scaled(f, qs);
                    }
                }
                {   // V^T: element (d, key r) -> key block r / 64, row d, column r % 64 (128-byte rows, Swizzle<3,4,3>)
                    const int kb = r >> 6, col = r & 63;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 u = *reinterpret_cast<const uint4 *>(st + OFF_VRAW + r * 64 + (i << 4));
                        const bf16 *e = reinterpret_cast<const bf16 *>(&u);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int d = i * 8 + j;
                            *reinterpret_cast<bf16 *>(st + OFF_VT + kb * 4096 + d * 128 + (((col >> 3) ^ (d & 7)) << 4) + (col & 7) * 2) = e[j];
                        }
                    }
                }
            }
            float *s_tab = reinterpret_cast<float *>(st + OFF_TAB);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = (int)threadIdx.x + i * 128;
                if (e < TAB) s_tab[(e / 31) * TS + e % 31] = fmaf(tabv[i], LOG2E, -tab_shift);
            }
            if (threadIdx.x == 0) *reinterpret_cast<float *>(st + OFF_SCALE) = sc;
            fence_proxy_async();                      // generic-proxy writes -> tensor core
            mbar_arrive(&bars->full[stage]);
            // only now the request for item n + 2: it waits for the stage of item n - 1, which may still be in use -- issued
            // before the staging of item n it serialised the two softmax streams (item n could not start before n - 1 ended)
            request(n + 2);
        }
        cp_async_wait<0>();
    } else if (warp == PROD_WARPS || warp == PROD_WARPS + 1) {
        // ============================== MMA issuers: one thread per softmax stream ==============================
        // stream g (softmax warps 9 + 4 g .. 12 + 4 g, TMEM columns [256 g, +256)) takes this CTA's items g, g + 2, ...; per
        // item and 128-query half: S = Q_h K^T, then (once the stream has written P) O = P V.  The streams drift apart, so one
        // stream's MMA / epilogue bubble is covered by the other stream's softmax: the MUFU pipe stays fed.  Two issuing
        // threads (warps 8 and 17, different sub-partitions) with blocking waits: one thread polling both streams either burnt
        // its sub-partition's issue slots or (with nanosleep) answered every hand-off ~1 us late.
        if (lane == 0) {
            const int g = warp - PROD_WARPS;
            const uint32_t idesc_s = umma_idesc(NTOK), idesc_o = umma_idesc(D);
            const uint32_t hi64 = (uint32_t)(umma_desc(0, 64) >> 32), hi128 = (uint32_t)(umma_desc(0, 128) >> 32);
            const uint32_t base = tmem + (uint32_t)(g * 256);
            uint32_t cnt = 0;               // (item, half) steps issued so far: barrier phases
            for (int n = g; n < my_items; n += 2) {
                const int stage = n % STAGES;
                mbar_wait(&bars->full[stage], (uint32_t)((n / STAGES) & 1));
                const uint32_t sbase = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t k_lo = (uint32_t)umma_desc(sbase + OFF_K, 64);
                const uint32_t v_lo = (uint32_t)umma_desc(sbase + OFF_VT, 128);
#pragma unroll 1
                for (int h = 0; h < 2; ++h, ++cnt) {
                    mbar_wait(&bars->tm_free[g], (cnt & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t q_lo = (uint32_t)umma_desc(sbase + OFF_Q + h * 8192, 64);
                    umma_ss_lo(base, q_lo, k_lo, hi64, idesc_s, 0u);
                    umma_ss_lo(base, q_lo + 2, k_lo + 2, hi64, idesc_s, 1u);
                    umma_commit(&bars->s_full[g]);
                    mbar_wait(&bars->p_full[g], cnt & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {    // 64 keys = 32 P columns per key block; keys 128.. live at column 128..
                        const uint32_t pa = base + (uint32_t)((kb >> 1) * 128 + (kb & 1) * 32);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_ts_lo(base + (uint32_t)TM_O, pa + 8 * k, v_lo + (uint32_t)(kb * 256 + 2 * k), hi128, idesc_o,
                                       (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&bars->o_full[g]);
                }
                umma_commit(&bars->empty[stage]);       // the item's last MMA: the stage is free once it retires (the stream
                                                        // finished with the stage's table / region ids before p_full)
            }
        }
    } else {
        // ============================== softmax streams: thread = (query row, 128 of the 256 keys) ==============================
        // 2 streams x 8 warps: four softmax warps per scheduler -- with two, every warp was latency bound (IPC ~0.25) and the
        // MUFU pipe sat at 40 %.  16-column steps keep the thread at 80 registers.
        const int sw = warp - SM_WARP0;                              // 0..15
        const int g = sw >> 3;                                       // stream
        const int wg = (sw >> 2) & 1;                                // which 128 keys (and which 16 output channels)
        const int row = (warp & 3) * 32 + lane;                      // query row inside the half == TMEM lane
        const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * 256);
        const float MASKED = -100.0f * LOG2E;
        float *sum_g = s_sum + g * 512;
        uint32_t cnt = 0;
        for (int n = g; n < my_items; n += 2) {
            int head, win, b;
            decode((int)blockIdx.x + n * (int)gridDim.x, head, win, b);
            const uint8_t *st = smem + (n % STAGES) * STAGE_BYTES;
            const float *s_tab = reinterpret_cast<const float *>(st + OFF_TAB);
            const uint8_t *reg = st + OFF_REG;
#pragma unroll 1
            for (int half = 0; half < 2; ++half, ++cnt) {
                mbar_wait(&bars->s_full[g], cnt & 1u);
                tc_fence_after();
                const float sc = *reinterpret_cast<const float *>(st + OFF_SCALE);
                const bool one_pass = 2.01f * sc + 16.0f < 80.0f;     // uniform per item
                const int r = half * 128 + row;                      // my query row inside the window
                const int my_reg = MASK ? (int)reg[r] : 0;
                // cpb bias[i][j] = table[(qy - ky + 15) * 31 + (qx - kx + 15)]: query part in a register, key part is a
                // per-step constant (one key row ky per 16-key step) plus a compile-time offset -> one LDS per logit
                const float *tab_q = s_tab + ((r >> 4) + 15) * TS + (r & 15) + 15;
                float ml = 0.0f;
                if (!one_pass) {        // exact row-max pre-pass (huge logit scales only); the scores stay in TMEM
                    float m = -INFINITY;
#pragma unroll 1
                    for (int c0 = wg * 128; c0 < wg * 128 + 128; c0 += 16) {
                        uint32_t raw[16];
                        tmem_ld16_nowait(t_row + (uint32_t)c0, raw);
                        tmem_ld_wait();
                        const float *tab = tab_q - (c0 >> 4) * TS;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float x = fmaf(__uint_as_float(raw[j]), LOG2E, tab[-j]);
                            if (MASK) x += ((int)reg[c0 + j] != my_reg) ? MASKED : 0.0f;
                            m = fmaxf(m, x);
                        }
                    }
                    float *mx = reinterpret_cast<float *>(smem + OFF_MAX) + g * 256;
                    mx[wg * 128 + row] = m;
                    named_bar(2 + g, 256);                            // the stream's 8 warps (the branch is uniform per item)
                    ml = fmaxf(m, mx[(wg ^ 1) * 128 + row]);
                    named_bar(2 + g, 256);
                }
                float l = 0.f;
                auto softmax_steps = [&](auto one_pass_c) {
                    constexpr bool ONE = decltype(one_pass_c)::value;
#pragma unroll 1
                    for (int stp = 0; stp < 8; ++stp) {
                        const int c0 = wg * 128 + stp * 16;                         // first key of this step (one key row)
                        uint32_t raw[16];
#ifdef SOCCDPT_EXP_NO_LDTM
#pragma unroll
                        for (int j = 0; j < 16; ++j) raw[j] = __float_as_uint(-1.0f - (float)(c0 + j) * 1e-3f * (float)row);
#else
                        tmem_ld16_nowait(t_row + (uint32_t)c0, raw);
                        tmem_ld_wait();
#endif
                        const float *tab = tab_q - (c0 >> 4) * TS;
                        uint32_t pk[8];
                        float l2 = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
#ifdef SOCCDPT_EXP_NO_LDS
                            float x0 = fmaf(__uint_as_float(raw[j]), LOG2E, -20.0f);
                            float x1 = fmaf(__uint_as_float(raw[j + 1]), LOG2E, -20.0f);
#else
                            float x0 = fmaf(__uint_as_float(raw[j]), LOG2E, tab[-j]);
                            float x1 = fmaf(__uint_as_float(raw[j + 1]), LOG2E, tab[-(j + 1)]);
#endif
                            if (MASK) {
                                x0 += ((int)reg[c0 + j] != my_reg) ? MASKED : 0.0f;
                                x1 += ((int)reg[c0 + j + 1] != my_reg) ? MASKED : 0.0f;
                            }
#ifdef SOCCDPT_EXP_NO_MUFU
                            x0 = fmaf(ONE ? x0 : x0 - ml, 1e-3f, 1.0f);
                            x1 = fmaf(ONE ? x1 : x1 - ml, 1e-3f, 1.0f);
#else
                            x0 = fast_exp2(ONE ? x0 : x0 - ml);
                            x1 = fast_exp2(ONE ? x1 : x1 - ml);
#endif
                            l += x0;
                            l2 += x1;
                            pk[j >> 1] = pack_bf16x2(x0, x1);
                        }
                        l += l2;
                        // P (bf16 pairs, key 2c in the low half) over score columns this thread has already consumed
                        tmem_st8(t_row + (uint32_t)(wg * 128 + stp * 8), pk);
                    }
                };
                if (one_pass) softmax_steps(std::true_type{});
                else softmax_steps(std::false_type{});
                tmem_st_wait();
                float *sums = sum_g + (cnt & 1u) * 256;
                sums[wg * 128 + row] = l;
                tc_fence_before();
                mbar_arrive(&bars->p_full[g]);
                // ---- epilogue: my 16 output channels: O / rowsum -> bf16 -> out[token, head * 32 + wg * 16 ...]
                int dummy;
                const long long tok = token_of(win, b, r, dummy);
                mbar_wait(&bars->o_full[g], cnt & 1u);
                tc_fence_after();
                uint32_t raw[16];
                tmem_ld16_nowait(t_row + (uint32_t)(TM_O + wg * 16), raw);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars->tm_free[g]);         // the stream's TMEM half is free for its next scores
                const float inv = 1.0f / (l + sums[(wg ^ 1) * 128 + row]);
                float o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(raw[i]);
                uint4 *op = reinterpret_cast<uint4 *>(a.out + tok * a.C + head * D + wg * 16);
                op[0] = pack8_scaled(o, inv);
                op[1] = pack8_scaled(o + 8, inv);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == PROD_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace

namespace soccdpt {
// qkv bf16 [B, Hs*Ws, 3C]; bias_tab f32 [heads][31*31] relative-position table; 16x16 windows
int launch_window_attention_ws(const void *qkv, const float *bias_tab, const float *scale, void *out, int batch, int Hs, int Ws,
                               int C, int heads, int shift, cudaStream_t st) {
    static SmemAttr configured;
    if (configured.need(SMEM_BYTES)) {
        SOCCDPT_CUDA(cudaFuncSetAttribute(window_attention_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SOCCDPT_CUDA(cudaFuncSetAttribute(window_attention_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    Args a;
    a.qkv = static_cast<const bf16 *>(qkv);
    a.bias_tab = bias_tab;
    a.scale = scale;
    a.out = static_cast<bf16 *>(out);
    a.Hs = Hs; a.Ws = Ws; a.C = C; a.heads = heads; a.shift = shift;
    a.nwx = Ws / WS; a.nwy = Hs / WS;
    a.items = batch * a.nwx * a.nwy * heads;
    const int grid = a.items < sm_count() ? a.items : sm_count();
    if (shift > 0)
        SOCCDPT_CUDA(launch_pdl(PDL_ATTENTION, window_attention_ws_kernel<true>, dim3(grid), dim3(THREADS), SMEM_BYTES, st, a));
    else
        SOCCDPT_CUDA(launch_pdl(PDL_ATTENTION, window_attention_ws_kernel<false>, dim3(grid), dim3(THREADS), SMEM_BYTES, st, a));
    return check_launch("window_attention_ws_kernel");
}
}  // namespace soccdpt
