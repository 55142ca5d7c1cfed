// SwinV2 window attention for 16x16 windows, TMA-fed and pipelined across (window, head) items (round 2).
// timm 0.6.12 WindowAttention + SwinTransformerBlock._attn (roll, window partition, cosine attention, cpb bias, shift mask,
// softmax, P V, window reverse, roll back) -- reference call site SOccDPT/model/backbones/swin_common.py:16-27.
//
// What the round-1 kernel (attention_tc.cu: one CTA per (window, head)) spent its time on was not the softmax: ~45 % of its
// 2900 warp instructions staged the operands (LDG -> L2-normalise q and k -> transpose V -> STS), and a CTA's load / stage /
// MMA / epilogue phases are serial, so the MUFU pipe (the true bound: 65 536 ex2 per item) was busy 30 % of the time.
// Here NOTHING is staged by threads:
//   * the qkv GEMM's epilogue (conv_tcgen05.cu, soccdpt_conv_t.qk_heads) already wrote q^ * scale * log2(e) | k^ | v,
//     normalised from the fp32 accumulator, so the tiles are plain copies of global memory;
//   * a window is fetched as 2 x 2 TMA boxes of 8 x 8 tokens x 32 channels (64-byte rows, SWIZZLE_64B) for each of q, k, v.
//     With shift = ws/2 = 8 a wrapped window splits exactly on the box grid, so the cyclic shift is the box coordinate
//     (x0 + shift) mod W and the kernel is the same for shifted and un-shifted blocks.  Tokens of a window sit in
//     box-major order in shared memory: attention is permutation-equivariant, the bias lookup and the mask follow the order;
//   * V is the B operand of P V in MN-major form (its [key][32 ch] rows as they are): no transpose;
//   * the shift mask is block structured in this order: the regions of timm's attn_mask are the 8 x 8 boxes of the windows in the
//     last window row / column, so a masked (query box, key box) pair is a warp-uniform "P = 0" (the reference adds -100 to those
//     logits: a factor e^-100 < 2^-126 / e^-80 below every un-masked term the one-pass bound admits, i.e. exactly 0 in bf16).
// One persistent 832-thread CTA per SM; all items of a CTA belong to ONE head (grid = a multiple of the head count), so the
// relative-position bias table is staged once.  Work unit = (item, 128-query half h, 64-key block kb = one 8 x 8 box of keys):
//   warp 24     TMA producer, 3 stages x 48 KB, up to three items ahead; it also decodes the item (frame, window, mask case) once
//   warp 25     MMA issuer:  S_u = Q^_h K^_kb^T (M128 N64 K32) into TMEM buffer u % 7, SIX units ahead of
//               O_h (+)= P_u V_kb (A operand = P in TMEM, M128 N32 K64), O in one of two 32-column accumulators.  The loop is
//               unrolled over the 8 units of an item and issues in warp-convergent, elect-predicated form (tc_ptx.cuh)
//   warps 0-19  five softmax groups of 128 threads (thread = query row = TMEM lane); unit u goes to group u % 5:
//               x = S + bias (log2 domain, pre-shifted by the analytic logit bound: ONE pass, no row max), ex2, row sum, bf16
//               pairs written back over the consumed score columns (tcgen05.st).  The bias comes as two LDS.128 per key row
//               from four shifted copies of the x-reversed table; adds and sums are packed fp32x2
//   warps 20-23 epilogue: O / (l_0 + .. + l_3) -> bf16 -> the token's un-shifted position
// Seven S buffers for five groups leave two units of slack for the P -> (P V, next S) -> group hand-off (~800 cycles through
// the single issuing warp); measured structure (tools/trace_attention.py, profiles/r2b_*):
//   * 3 groups x 128-key units on 3 buffers: the groups waited 29-47 % of the time for their next S tile;
//   * issue inside `if (lane == 0)` cost ~88 cycles per MMA on the issuing thread (ELECT / BRA.U.ANY loop per instruction);
//   * tensor time is irrelevant (tools/microbench/mma_latency.cu: 17 cycles per N = 32 TS MMA, 48 per N = 64 SS MMA);
//   * inside the softmax loops the MUFU pipe is ~85 % busy; what is left is hand-off slack and the per-launch ramp.
// Precondition (checked by the engine at pack time): every head's logit scale satisfies 2.01 * scale + 16 < 80.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using bf16 = __nv_bfloat16;
using namespace tc;

#ifndef SOCCDPT_ATTN_GROUPS
#define SOCCDPT_ATTN_GROUPS 5
#endif
constexpr int A_GROUPS = SOCCDPT_ATTN_GROUPS;      // softmax groups of 128 threads; unit u goes to group u % A_GROUPS
constexpr int A_SM_WARPS = 4 * A_GROUPS;
constexpr int A_THREADS = (A_SM_WARPS + 6) * 32;   // + 4 epilogue warps + TMA producer + MMA issuer
constexpr int A_STAGES = 3;
constexpr int A_TILE = 16384;                     // q^, k^ or v of one (window, head): 256 rows x 64 B
constexpr int A_STAGE_BYTES = 3 * A_TILE;
constexpr int A_TS = 40;                          // bias-table row stride (floats); copy stride 31 * 40 = 24 (mod 32): see the table staging
constexpr int A_SUM_SLOTS = 8;                    // row-sum slots per half (see the hazard note at s_sum)
constexpr int A_SMEM_TAB = A_STAGES * A_STAGE_BYTES;
constexpr int A_SMEM_SUM = A_SMEM_TAB + 4 * 31 * A_TS * 4;   // four shifted copies of the bias table (19 840 B)
constexpr int A_SMEM_BAR = A_SMEM_SUM + A_SUM_SLOTS * 4 * 128 * 4;
constexpr int A_SMEM_TRACE = A_SMEM_BAR + 512;
constexpr int A_BUFS = 7;                         // S / P buffers of 64 columns (one 64-key block = one 8 x 8 box of keys)
constexpr int A_LAG = A_BUFS - 1;                 // the scores run six units ahead of P V
constexpr int A_TM_O = A_BUFS * 64;               // TMEM: S/P buffers 7 x 64 columns, O accumulators 2 x 32 columns
constexpr uint32_t A_DESC_HI = 32u | (1u << 14) | (4u << 29);   // SBO = 512 B, descriptor version 1, SWIZZLE_64B

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Item {
    int b, wy, wx;
};

// -DSOCCDPT_ATTN_TRACE (debug builds only, tools/trace_attention.py): CTA 0 logs (event, unit, SM clock) triples
#ifdef SOCCDPT_ATTN_TRACE
constexpr int A_TRACE_IDS = 384, A_TRACE_KINDS = 11;
__device__ unsigned int g_attn_trace[A_TRACE_KINDS * A_TRACE_IDS];
#define A_TRACE_BYTES (A_TRACE_KINDS * A_TRACE_IDS * 4)
#define ATRACE(kind, id)                                                                                    \
    do {                                                                                                    \
        if ((id) < A_TRACE_IDS) s_trace[(kind) * A_TRACE_IDS + (id)] = (unsigned int)clock64();             \
    } while (0)
#else
#define ATRACE(kind, id) do { } while (0)
#define A_TRACE_BYTES 0
#endif

constexpr int A_SMEM_BYTES = A_SMEM_TRACE + A_TRACE_BYTES + 1024;   // + alignment slack

__global__ void __launch_bounds__(A_THREADS, 1)
window_attention_tma_kernel(const __grid_constant__ CUtensorMap mqkv, const float *__restrict__ bias_tab,
                            const float *__restrict__ scale, bf16 *__restrict__ out, int Hs, int Ws, int C, int shift, int nheads,
                            int n_items) {
    extern __shared__ uint8_t at_raw[];
    uint8_t *smem = at_raw + ((1024u - (smem_u32(at_raw) & 1023u)) & 1023u);
    float *s_tab = reinterpret_cast<float *>(smem + A_SMEM_TAB);
    // partial row sums [slot][kb][128]: written by the softmax thread of (half, kb), read by the epilogue of the half.  Slot =
    // half % 8: the scores of half H + 8 are issued after P V of half H + 4 (kb 0), which waited for the epilogue of half H.
    float *s_sum = reinterpret_cast<float *>(smem + A_SMEM_SUM);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + A_SMEM_BAR);
    uint64_t *full = bars, *empty = bars + 3, *s_full = bars + 6, *p_ready = bars + 14, *o_full = bars + 22, *o_empty = bars + 24;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 26);
    int4 *s_item = reinterpret_cast<int4 *>(bars + 28);      // [8] decoded items {b, wy, wx, mask flags}, written by the producer
#ifdef SOCCDPT_ATTN_TRACE
    unsigned int *s_trace = reinterpret_cast<unsigned int *>(smem + A_SMEM_TRACE);
    for (int i = threadIdx.x; i < A_TRACE_KINDS * A_TRACE_IDS; i += A_THREADS) s_trace[i] = 0u;
#endif

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.x % nheads;           // gridDim.x % nheads == 0: item = blockIdx.x + n * gridDim.x keeps its head
    const int n_my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int U = 8 * n_my;                          // units of this CTA: item x 2 query halves x 4 key boxes
    const int nwx = Ws >> 4, nwy = Hs >> 4, nw = nwx * nwy;
    auto item_of = [&](int n) {
        const int widx = ((int)blockIdx.x + n * (int)gridDim.x) / nheads;
        const int win = widx % nw;
        return Item{widx / nw, win / nwx, win % nwx};
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < A_BUFS; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_ready[i], 128);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 128);
        }
        fence_barrier_init();
        prefetch_tensormap(&mqkv);
    }
    if (warp == A_SM_WARPS + 5) tmem_alloc(tmem_slot, 512);
    {   // relative-position bias of this head (weights: no dependence on the previous kernel), log2 domain, shifted by the
        // analytic bound of the logits: |S| <= ~1.004 * scale (bf16-rounded unit vectors), bias in (0, 16), masks only subtract.
        // A thread needs, per key row, the 8 consecutive entries T[qy - ky + 15][qx - kx + 15], kx = 8 bx .. 8 bx + 7: stored
        // REVERSED in x (ascending kx = ascending address) and in four copies shifted by a = 0..3 floats, so that the thread with
        // (15 - qx) % 4 == a reads them as two aligned LDS.128 from copy a: copy_a[dy][m] = T[dy][30 - m - a].  Four LDS.128
        // instead of 32 LDS per 32 scores (shared-memory loads and MUFU share the MIO issue port: it was the bound).  The 8 lanes
        // of an LDS.128 phase (one query row of a box) hit copies 3,2,1,0,3,2,1,0 at m0 = 12,12,12,12,8,8,8,8 (+ const): with
        // the copy stride = 24 (mod 32) floats they cover the 32 banks exactly once.
        const float LOG2E = 1.4426950408889634f;
        const float tab_shift = (1.01f * scale[head] + 16.0f) * LOG2E;
        for (int e = threadIdx.x; e < 4 * 31 * A_TS; e += A_THREADS) {
            const int a = e / (31 * A_TS), r = e - a * (31 * A_TS), dy = r / A_TS, m = r - dy * A_TS;
            const int dx = 30 - m - a;
            s_tab[e] = (dx >= 0 && dx <= 30) ? fmaf(bias_tab[(size_t)head * 961 + dy * 31 + dx], LOG2E, -tab_shift) : 0.0f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    soccdpt::pdl_wait();        // qkv is the previous kernel's output; `out` may still be read by the one before

    if (warp == A_SM_WARPS + 4) {
        // ===================== TMA producer (warp-convergent, elect-predicated issue)
        int st = 0;
        uint32_t eph = 1;                                        // first pass over the ring: the slots are free
        for (int n = 0; n < n_my; ++n) {
            mbar_wait(&empty[st], eph);
            ATRACE(1, n);
            const Item it = item_of(n);
            // decoded once here; the softmax / epilogue threads read it after the full -> s_full / o_full barrier chain
            // (ring of 8: the epilogue lags the producer by at most 3 stages + 1 item of O accumulators)
            if (lane == 0)
                s_item[n & 7] = make_int4(it.b, it.wy, it.wx, (shift > 0 && it.wy == nwy - 1 ? 1 : 0) | (shift > 0 && it.wx == nwx - 1 ? 2 : 0));
            __syncwarp();
            uint8_t *stg = smem + st * A_STAGE_BYTES;
            mbar_expect_tx_elect(&full[st], A_STAGE_BYTES);
#pragma unroll
            for (int box = 0; box < 4; ++box) {
                const int x0 = (it.wx * 16 + (box & 1) * 8 + shift) % Ws, y0 = (it.wy * 16 + (box >> 1) * 8 + shift) % Hs;
#pragma unroll
                for (int part = 0; part < 3; ++part)
                    tma_load_4d_elect(stg + part * A_TILE + box * 4096, &mqkv, &full[st], part * C + head * 32, x0, y0, it.b);
            }
            if (++st == A_STAGES) {
                st = 0;
                eph ^= 1;
            }
        }
    } else if (warp == A_SM_WARPS + 5) {
        // ===================== MMA issuer: the whole warp runs the loop in uniform control flow, elect.sync picks the issuing
        // lane inside each instruction (tc_ptx.cuh: a bare UTCHMMA per MMA instead of an ELECT loop).  This warp is the pacemaker
        // of the pipeline and shares its scheduler with softmax warps: the loop is unrolled over the 8 units of an item (h, kb and
        // the P V unit six behind are compile-time), buffers / stages / phases advance by increments, no division anywhere.
        static_assert(A_BUFS == 7, "the unrolled issue loop pairs unit j of an item with the P V of unit j - 6");
        const uint32_t idesc_s = umma_idesc(64);
        const uint32_t idesc_pv = umma_idesc(32) | (1u << 16);      // B (= V) MN-major
        const uint32_t smem_lo = (smem_u32(smem) >> 4) | (1u << 16);   // descriptor low word of the stage base (LBO = 1)
        int bs = 0, bp = 0;              // S buffer of the next score unit, P buffer of the next P V unit (unit % 7)
        uint32_t pmask = 0;              // phase bit of p_ready[b]
        int hh_pv = 0;                   // half (item x 2 + h) of the next P V unit
        int u_s = 0, u_pv = 0;           // trace only
        auto issue_s = [&](uint32_t stage_lo, int h, int kb) {
            // buffer bs held P of the unit seven back: its P V was issued just before, the tensor pipe runs in order
            const uint32_t a_lo = stage_lo + (uint32_t)((h * 8192) >> 4), b_lo = stage_lo + (uint32_t)((A_TILE + kb * 4096) >> 4);
            umma_ss_lo_elect(tmem + bs * 64, a_lo, b_lo, A_DESC_HI, idesc_s, 0u);
            umma_ss_lo_elect(tmem + bs * 64, a_lo + 2, b_lo + 2, A_DESC_HI, idesc_s, 1u);
            umma_commit_elect(&s_full[bs]);
            ATRACE(4, u_s);
            ++u_s;
            bs = bs == A_BUFS - 1 ? 0 : bs + 1;
        };
        auto issue_pv = [&](uint32_t stage_lo, int kb, uint64_t *stage_done) {
            mbar_wait(&p_ready[bp], (pmask >> bp) & 1u);
            pmask ^= 1u << bp;
            ATRACE(5, u_pv);
            const int o = hh_pv & 1;
            if (kb == 0) mbar_wait(&o_empty[o], ((hh_pv >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t v_lo = stage_lo + (uint32_t)((2 * A_TILE + kb * 4096) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)      // 16 keys per step: 8 P columns, 16 V rows = 1024 B
                umma_ts_lo_elect(tmem + A_TM_O + o * 32, tmem + bp * 64 + 8 * k, v_lo + 64 * k, A_DESC_HI, idesc_pv, (kb | k) != 0 ? 1u : 0u);
            if (kb == 3) {
                umma_commit_elect(&o_full[o]);
                ++hh_pv;
            }
            if (stage_done) umma_commit_elect(stage_done);     // every MMA that reads this stage has been issued
            ATRACE(6, u_pv);
            ++u_pv;
            bp = bp == A_BUFS - 1 ? 0 : bp + 1;
        };
        int st = 0, prev_st = 0;
        uint32_t fph = 0, prev_lo = 0;
        for (int n = 0; n < n_my; ++n) {
            ATRACE(2, n);
            mbar_wait(&full[st], fph);
            tc_fence_after();
            ATRACE(3, n);
            const uint32_t cur_lo = smem_lo + (uint32_t)((st * A_STAGE_BYTES) >> 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                issue_s(cur_lo, j >> 2, j & 3);
                if (j >= 6) issue_pv(cur_lo, (j - 6) & 3, nullptr);                                   // units 0, 1 of this item
                else if (n > 0) issue_pv(prev_lo, (j + 2) & 3, j == 5 ? &empty[prev_st] : nullptr);   // units 2 .. 7 of the previous one
            }
            prev_lo = cur_lo;
            prev_st = st;
            if (++st == A_STAGES) {
                st = 0;
                fph ^= 1;
            }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) issue_pv(prev_lo, (j + 2) & 3, j == 5 ? &empty[prev_st] : nullptr);
    } else if (warp < A_SM_WARPS) {
        // ===================== softmax groups: group g works through the units u = g (mod A_GROUPS); unit u lives in buffer u % 7
        const int g = warp >> 2;
        const int row = (warp & 3) * 32 + lane;                 // query row of the half == TMEM lane
        const int my_bx = row >> 6;                              // box column of my query
        const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int qx = my_bx * 8 + (row & 7), cpy = (15 - qx) & 3;
        const float *tab_row = s_tab + (cpy * 31 + ((row >> 3) & 7) + 15) * A_TS + (15 - qx - cpy);      // 16-byte aligned
        int b = g % A_BUFS;
        uint32_t bph = (uint32_t)(g / A_BUFS) & 1u;               // (u / A_BUFS) & 1, advanced incrementally
        for (int u = g; u < U; u += A_GROUPS) {
            const int n = u >> 3, h = (u >> 2) & 1, kb = u & 3;
            const uint32_t t_row = t_lane + (uint32_t)(b * 64);
            if ((threadIdx.x & 127) == 0) ATRACE(7, u);
            mbar_wait(&s_full[b], bph);
            tc_fence_after();
            if ((threadIdx.x & 127) == 0) ATRACE(8, u);
            // timm's mask regions inside a window of the last window row are its upper / lower 8 rows = the box rows (same for
            // the box columns of the last window column): a query sees a key box only if both box coordinates match there
            const int flags = s_item[n & 7].w;
            const bool masked = ((flags & 1) && h != (kb >> 1)) || ((flags & 2) && my_bx != (kb & 1));     // warp uniform
            // key (ky, kx) of column j in chunk ch: ky = (kb>>1)*8 + ch*4 + (j>>3), kx = (kb&1)*8 + (j&7)
            // bias = tab[(qy - ky + 15)][(qx - kx + 15)], qy = h*8 + ((row>>3)&7)
            const float *tab_u = tab_row + (h - (kb >> 1)) * 8 * A_TS + (kb & 1) * 8;
            f32x2 l2 = mk2(0.f, 0.f);
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t pk[16];
                if (masked) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = 0u;
                } else {
                    uint32_t v[32];
                    tmem_ld32_nowait(t_row + (uint32_t)(ch * 32), v);
                    tmem_ld_wait();
                    const float *tab = tab_u - (ch * 4) * A_TS;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {           // key row r of the chunk: 8 consecutive kx
                        const float4 ta = *reinterpret_cast<const float4 *>(tab - r * A_TS);
                        const float4 tb = *reinterpret_cast<const float4 *>(tab - r * A_TS + 4);
                        const float t[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {
                            float x0, x1;
                            un2(fadd2(mk2(__uint_as_float(v[8 * r + j]), __uint_as_float(v[8 * r + j + 1])), mk2(t[j], t[j + 1])), x0, x1);
                            const float e0 = ex2(x0), e1 = ex2(x1);
                            l2 = fadd2(l2, mk2(e0, e1));
#ifdef SOCCDPT_EXP_PACK      // experiment: bf16 pairs by integer rounding (ALU pipe) instead of F2FP
                            pk[(8 * r + j) >> 1] = __byte_perm(__float_as_uint(e0) + 0x8000u, __float_as_uint(e1) + 0x8000u, 0x7632);
#else
                            pk[(8 * r + j) >> 1] = pack_bf16x2(e0, e1);
#endif
                        }
                    }
                }
                // P (bf16 pairs, key 2c in the low half) over score columns this thread has already consumed
                tmem_st16(t_row + (uint32_t)(ch * 16), pk);
            }
            tmem_st_wait();
            float l0, l1;
            un2(l2, l0, l1);
            s_sum[((u >> 2) & (A_SUM_SLOTS - 1)) * 512 + kb * 128 + row] = l0 + l1;
            tc_fence_before();
            mbar_arrive(&p_ready[b]);
            if ((threadIdx.x & 127) == 0) ATRACE(9, u);
            b += A_GROUPS;
            if (b >= A_BUFS) {
                b -= A_BUFS;
                bph ^= 1u;
            }
            static_assert(A_GROUPS <= A_BUFS, "one wrap per step");
        }
    } else if (warp < A_SM_WARPS + 4) {
        // ===================== epilogue
        const int row = (warp & 3) * 32 + lane;
        const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)A_TM_O;
        for (int hh = 0; hh < 2 * n_my; ++hh) {
            const int o = hh & 1, h = hh & 1;
            mbar_wait(&o_full[o], (hh >> 1) & 1);
            tc_fence_after();
            if ((threadIdx.x & 127) == 0) ATRACE(10, hh);
            uint32_t v[32];
            tmem_ld32_nowait(t_row + (uint32_t)(o * 32), v);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&o_empty[o]);
            const float *sl = s_sum + (hh & (A_SUM_SLOTS - 1)) * 512 + row;
            const float inv = 1.0f / ((sl[0] + sl[128]) + (sl[256] + sl[384]));
            const int4 iti = s_item[(hh >> 1) & 7];
            const Item it{iti.x, iti.y, iti.z};
            const int ty = h * 8 + ((row >> 3) & 7), tx = (row >> 6) * 8 + (row & 7);
            const int yo = (it.wy * 16 + ty + shift) % Hs, xo = (it.wx * 16 + tx + shift) % Ws;
            uint4 *op = reinterpret_cast<uint4 *>(out + (((long long)it.b * Hs + yo) * Ws + xo) * C + head * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 w;
                w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
                w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
                w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
                w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
                op[q] = w;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
#ifdef SOCCDPT_ATTN_TRACE
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < A_TRACE_KINDS * A_TRACE_IDS; i += A_THREADS) g_attn_trace[i] = s_trace[i];
#endif
    if (warp == A_SM_WARPS + 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace

#ifdef SOCCDPT_ATTN_TRACE
extern "C" int soccdpt_debug_attention_trace(unsigned int *host) {
    return (int)cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(unsigned int) * A_TRACE_KINDS * A_TRACE_IDS);
}
#endif

/* Window attention on PRE-NORMALISED operands (soccdpt_conv_t.qk_heads epilogue of the qkv linear). */
extern "C" int soccdpt_window_attention_normed_fwd(const void *qkvn, const float *bias, const float *scale, void *out, int batch,
                                                   int Hs, int Ws, int C, int heads, int shift, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(qkvn && bias && scale && out, "window_attention_normed: NULL pointer");
    SOCCDPT_REQUIRE(batch >= 1 && heads >= 1 && C == heads * 32, "window_attention_normed: head_dim must be 32 (C=%d heads=%d)", C, heads);
    SOCCDPT_REQUIRE(Hs >= 16 && Ws >= 16 && Hs % 16 == 0 && Ws % 16 == 0, "window_attention_normed: 16x16 windows must tile %dx%d", Hs, Ws);
    SOCCDPT_REQUIRE(shift == 0 || shift == 8, "window_attention_normed: shift must be 0 or 8 (got %d)", shift);
    SOCCDPT_REQUIRE((reinterpret_cast<uintptr_t>(qkvn) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                    "window_attention_normed: qkv / out must be 16-byte aligned");
    tc::EncodeTiledFn encode = tc::encode_fn();
    SOCCDPT_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap map;
    {
        cuuint64_t dims[4] = {(cuuint64_t)3 * C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)batch};
        cuuint64_t strides[3] = {(cuuint64_t)3 * C * 2, (cuuint64_t)Ws * 3 * C * 2, (cuuint64_t)Hs * Ws * 3 * C * 2};
        cuuint32_t box[4] = {32, 8, 8, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(qkvn), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "window_attention_normed: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    const long long items = (long long)batch * (Hs / 16) * (Ws / 16) * heads;
    SOCCDPT_REQUIRE(items < (1ll << 29), "window_attention_normed: batch too large for one call");
    SOCCDPT_REQUIRE(heads <= soccdpt::sm_count(), "window_attention_normed: more heads than SMs");
    long long grid = soccdpt::sm_count() / heads * heads;       // a multiple of the head count: one head per CTA
    if (grid > items) grid = items;
    static soccdpt::SmemAttr configured;
    if (configured.need(A_SMEM_BYTES)) {
        SOCCDPT_CUDA(cudaFuncSetAttribute(window_attention_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM_BYTES));
    }
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ATTENTION, window_attention_tma_kernel, dim3((unsigned)grid), dim3(A_THREADS),
                                     (size_t)A_SMEM_BYTES, soccdpt::as_stream(stream), map, bias, scale, static_cast<bf16 *>(out), Hs, Ws, C,
                                     shift, heads, (int)items));
    return soccdpt::check_launch("window_attention_tma_kernel");
}
