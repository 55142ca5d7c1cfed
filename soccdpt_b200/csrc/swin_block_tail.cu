// Fused tail of a SwinV2 (res-post-norm) block branch on tcgen05 / TMEM / TMA (sm_100a only):
//
//   two-GEMM mode (the MLP branch, timm SwinTransformerV2Block: x = x + norm2(mlp(x))):
//       master += LayerNorm( fc2( GELU( fc1(x) + b1 ) ) + b2 ) ;  y = bf16(master)
//   one-GEMM mode (the attention branch after the window attention: x = x + norm1(proj(attn))):
//       master += LayerNorm( x @ w2^T + b2 ) ;  y = bf16(master)
//
// What the fusion removes (dpt_swin2_tiny_256, B = 64, stage 0: 262 144 tokens x 96 channels): the 4C-wide hidden
// activation (200 MB written by fc1 and read back by fc2), the bf16 branch output (50 MB written, 50 MB read) and the
// separate LayerNorm + residual pass: 1.0 GB of HBM traffic per MLP branch becomes 0.3 GB (x in, master in/out, y out).
//
// One persistent CTA (512 threads) per SM, 128 token rows per tile:
//   * warp 0 (one lane): TMA producer -- the tile's rows [128 x K1] (K chunks of 64 channels in the 128B-swizzled K-major
//     layout, a 32-channel tail chunk in the 64B-swizzled layout: C = 96 = 64 + 32), and the weight tiles of fc1 ([128 hidden
//     x 64 K], ring 1) and fc2 ([C x 64 K], ring 2).  When all weight tiles of one M tile fit (stage 0: 144 KB) they are
//     loaded ONCE per CTA, before griddepcontrol.wait (weights are constants), and stay resident.
//   * warp 1 (one lane): MMA issuer.  The hidden dimension is processed in chunks of 128 columns:
//       fc1(j): D1[j % 2] (TMEM, 128 fp32 columns) = A(smem) x W1_j(smem)                    tcgen05.mma SS, N = 128
//       fc2(j): D2 (TMEM, C fp32 columns) += H_j (TMEM, bf16 pairs) x W2_j(smem)              tcgen05.mma TS, N = C
//     H_j is written by the GELU warps OVER the D1 columns they have just consumed (tcgen05.st), so the hidden activation
//     never touches shared memory: no staging buffer, no proxy fence, and the narrow-N fc2 MMA is not bound by an A-operand
//     read from shared memory.  Issue order fc1(j+1), fc2(j): the tensor pipe works ahead of the GELU warps.
//   * warps 4..11 (two-GEMM mode): GELU, thread = (token row, half of the chunk's columns): D1 -> + b1 -> erf GELU (packed
//     fp32x2 math, one MUFU per element) -> bf16 pairs -> TMEM.  Issue / FP32-pipe bound.
//   * warps 12..15 (two-GEMM mode) / warps 4..11 on alternating tiles (one-GEMM mode): LayerNorm + residual, thread = token
//     row: D2 -> + b2 -> mean / variance -> normalise -> transposed through the warp's private scratch so that one
//     instruction touches whole 128-byte row segments -> master (fp32, read-modify-write) and y (bf16).  HBM bound.  The
//     residual rows are streamed in by cp.async two 32-column blocks ahead, across tile boundaries.
//     GELU of tile i + 1 runs while the LayerNorm pass of tile i streams: with one set of warps doing both, every CTA of
//     the grid alternated between the two phases in lockstep and HBM idled half of the time (profiles/r2_progress.md).
//   * TMEM: D1 double-buffered (columns 0..255), D2 at column 256 (double-buffered while 2 C <= 256).
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

// -DSOCCDPT_TAIL_TRACE (debug builds only, tools/trace_block_tail.py): CTA 0 logs SM-clock stamps of its hand-offs (one lane of the MMA
// warp, of the first GELU warp and of the first LayerNorm warp) into a global buffer read back by soccdpt_block_tail_trace_read()
#ifdef SOCCDPT_TAIL_TRACE
constexpr int BT_EVENTS = 24, BT_IDS = 64;
__device__ unsigned long long g_bt_trace[BT_EVENTS * BT_IDS];
#define BTRACE(ev, id)                                                                                                   \
    do {                                                                                                                 \
        if (blockIdx.x == 0 && (id) < BT_IDS) g_bt_trace[(ev) * BT_IDS + (id)] = (unsigned long long)clock64();          \
    } while (0)
#else
#define BTRACE(ev, id) do { } while (0)
#endif

namespace {

using bf16 = __nv_bfloat16;
using namespace tc;

constexpr int BM = 128;                 // token rows per tile
constexpr int HC = 128;                 // hidden columns per chunk
constexpr int NT = 512;                 // 4 control warps + 8 GELU warps + 4 LayerNorm warps
constexpr int MAX_SLOTS = 12;
constexpr int W1_SLOT = 16384;          // [128 hidden rows x 64 K] bf16
constexpr int SMEM_MAX = 227 * 1024;
constexpr int LN_RING = 2;              // 32-column blocks of the residual stream in flight per LayerNorm warp
constexpr int LN_WARP_BYTES = 4096 * (1 + LN_RING);   // transposition scratch + cp.async ring

struct Params {
    soccdpt_block_tail_t a;
    int tiles;
    int kc64, ktail;                    // K1 = 64 * kc64 + 32 * ktail
    int nch;                            // HID / 128; 0 in one-GEMM mode
    int ns1, ns2;                       // ring depths (slots)
    int resident;                       // every weight tile of an M tile has its own slot: loaded once per CTA
    int stream_a;                       // one-GEMM mode, large K1 / C > 256: activations AND weights stream through ring 2 by K chunk
                                        // (slot = [128 x 64] activation chunk + [C x 64] weight chunk); N = C as 256 + (C - 256)
    int d2_bufs;
    uint32_t a_bytes;                   // bytes of one activation tile (= its TMA transaction count)
    uint32_t w2_slot;                   // bytes per ring-2 slot (>= C * 128, multiple of 1024)
    uint32_t w1_chunk;                  // resident mode: bytes of one hidden chunk's fc1 tiles, packed (kc64 * 16 KB + ktail * 8 KB)
    uint32_t off_w1, off_w2, off_ln, off_misc;   // shared-memory offsets (the activation tile sits at 0)
    uint32_t off_b2, off_g, off_be, off_bar;     // inside misc: b1 at 0, then b2 / gamma / beta / barriers
};

struct Bars {
    uint64_t a_full, a_empty;
    uint64_t w1_full[MAX_SLOTS], w1_empty[MAX_SLOTS];
    uint64_t w2_full[MAX_SLOTS], w2_empty[MAX_SLOTS];
    uint64_t d1_full[2], h_full[2];
    uint64_t d2_full[2], d2_empty[2];
    uint32_t tmem_slot;
};

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm + residual by one warp (32 token rows = its 32 TMEM lanes, all C columns), a stream of 32-column blocks
// q = (tile of this warp group, block) whose residual values arrive through a cp.async ring.
struct LnWarp {
    const soccdpt_block_tail_t *a;
    const float *b2, *g, *be;       // shared memory, [C]
    float *scr;                     // 4 KB transposition scratch
    float *ring;                    // LN_RING x 4 KB
    int lane, sub;                  // sub = warp % 4: rows [32 * sub, +32) of a tile
    int nb;                         // 32-column blocks per tile (C / 32)
    int tile0, tile_step, n_tiles;  // this warp group's tiles: tile0 + k * tile_step, k < n_tiles

    __device__ __forceinline__ long long row0_of(int k) const { return (long long)(tile0 + k * tile_step) * BM + sub * 32; }
    // requests the next block of the stream (beyond its end only the (empty) group is committed); ring slot = request order
    int fk, fb, fslot;              // fetch cursor: (tile k, block b) and the ring slot it goes to
    // Per-lane addressing of a tile, set once per tile (the per-row 64-bit products and bound checks of the first version made a
    // 32-column block ~700 instructions of one latency-bound warp: tools/trace_block_tail.py): lane = (row rr + 4 i, 16-byte chunk ch);
    // element offset of (row rr, chunk ch) from the tile's first row, the step of 4 rows, and the number of valid rows - rr.
    int lane_off, rstep;            // rr * C + ch * 4, 4 * C
    const float *fsrc;              // fetch cursor: master + row0_of(fk) * C + lane_off
    int frows;                      // valid rows of the fetch tile minus rr (row rr + 4 i is valid iff 4 i < frows)
    __device__ __forceinline__ void set_fetch_tile() {
        if (fk < n_tiles) {
            const long long r0 = row0_of(fk);
            fsrc = a->master + r0 * a->C + lane_off;
            const long long left = a->M - r0;
            frows = (int)(left < 32 ? left : 32) - (lane >> 3);
        }
    }
    __device__ __forceinline__ void fetch_next() {
        if (fk < n_tiles) {
            float *dst = ring + fslot * 1024 + lane * 4;
            const float *srcp = fsrc + fb * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (4 * i < frows) cp_async16(dst + i * 128, srcp);
                srcp += rstep;
            }
        }
        cp_async_commit();
        if (++fb == nb) { fb = 0; ++fk; set_fetch_tile(); }
        if (++fslot == LN_RING) fslot = 0;
    }
    // block q of the stream: TMEM columns [c, c + 32) of this warp's rows at taddr
    __device__ __forceinline__ void block(int slot, uint32_t taddr, int c, float rstd, float nmr, float *mrow, bf16 *yrow, int rows, int tk = -1) const {
        uint32_t raw[32];
        if (tk >= 0) BTRACE(14, tk);
        tmem_ld32_nowait(taddr + (uint32_t)c, raw);
        tmem_ld_wait();
        if (tk >= 0) BTRACE(15, tk);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 bb = *reinterpret_cast<const float4 *>(b2 + c + j * 4);
            const float4 gg = *reinterpret_cast<const float4 *>(g + c + j * 4);
            const float4 ee = *reinterpret_cast<const float4 *>(be + c + j * 4);
            float4 v;
            v.x = fmaf((__uint_as_float(raw[j * 4 + 0]) + bb.x) * rstd + nmr, gg.x, ee.x);
            v.y = fmaf((__uint_as_float(raw[j * 4 + 1]) + bb.y) * rstd + nmr, gg.y, ee.y);
            v.z = fmaf((__uint_as_float(raw[j * 4 + 2]) + bb.z) * rstd + nmr, gg.z, ee.z);
            v.w = fmaf((__uint_as_float(raw[j * 4 + 3]) + bb.w) * rstd + nmr, gg.w, ee.w);
            *reinterpret_cast<float4 *>(scr + lane * 32 + ((j ^ (lane & 7)) << 2)) = v;
        }
        if (tk >= 0) BTRACE(16, tk);
        cp_async_wait<LN_RING - 1>();          // this thread's copies of block q have landed (it reads back only its own)
        __syncwarp();
        if (tk >= 0) BTRACE(17, tk);
        const int rr = lane >> 3, ch = lane & 7;
        const float *src = ring + slot * 1024 + lane * 4;
        float *mp = mrow + c;               // (row rr, chunk ch) of the block; + rstep per 4 rows
        bf16 *yp = yrow + c;
        // loads of four rows first, then their (predicated) stores: with the row-bound test around the whole body every row was its
        // own branch region and paid the shared-memory latency serially (~120 cycles per row: tools/trace_block_tail.py)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = (h * 4 + i) * 4 + rr;
                const float4 v = *reinterpret_cast<const float4 *>(scr + rl * 32 + ((ch ^ (rl & 7)) << 2));
                const float4 m = *reinterpret_cast<const float4 *>(src + (h * 4 + i) * 128);      // stale ring contents for rows past the end
                o[i].x = m.x + v.x; o[i].y = m.y + v.y; o[i].z = m.z + v.z; o[i].w = m.w + v.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint2 u;
                u.x = pack_bf16x2(o[i].x, o[i].y);
                u.y = pack_bf16x2(o[i].z, o[i].w);
                if (4 * (h * 4 + i) < rows) {
                    *reinterpret_cast<float4 *>(mp) = o[i];
                    *reinterpret_cast<uint2 *>(yp) = u;
                }
                mp += rstep;
                yp += rstep;
            }
        }
        __syncwarp();
        if (tk >= 0) BTRACE(18, tk);
    }
    // one tile: statistics in one TMEM pass (sums shifted by the row's first value), then the blocks
    int cslot;                      // ring slot of the next block to consume
    __device__ __forceinline__ void tile(int k, uint32_t d2) {
        const int C = a->C;
        float s1 = 0.f, s2 = 0.f, shift = 0.f;
        for (int c = 0; c < C; c += 32) {
            uint32_t r[32];
            tmem_ld32_nowait(d2 + (uint32_t)c, r);
            tmem_ld_wait();
            const float4 *b4 = reinterpret_cast<const float4 *>(b2 + c);
            if (c == 0) shift = __uint_as_float(r[0]) + b2[0];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = b4[i];
                const float d0 = __uint_as_float(r[4 * i]) + b.x - shift, d1 = __uint_as_float(r[4 * i + 1]) + b.y - shift;
                const float d2v = __uint_as_float(r[4 * i + 2]) + b.z - shift, d3 = __uint_as_float(r[4 * i + 3]) + b.w - shift;
                s1 += (d0 + d1) + (d2v + d3);
                s2 = fmaf(d0, d0, s2);
                s2 = fmaf(d1, d1, s2);
                s2 = fmaf(d2v, d2v, s2);
                s2 = fmaf(d3, d3, s2);
            }
        }
        const float inv_c = 1.0f / (float)C;
        const float ms = s1 * inv_c;                               // mean - shift
        const float var = fmaxf(fmaf(-ms, ms, s2 * inv_c), 0.0f);
        const float rstd = rsqrtf(var + a->eps);
        const float nmr = -(shift + ms) * rstd;
        const long long r0 = row0_of(k);
        float *mrow = a->master + r0 * C + lane_off;
        bf16 *yrow = static_cast<bf16 *>(a->y) + r0 * C + lane_off;
        const long long left = a->M - r0;
        const int rows = (int)(left < 32 ? left : 32) - (lane >> 3);
        if (trace_lane) BTRACE(12, k);
        for (int b = 0; b < nb; ++b) {
            block(cslot, d2, b * 32, rstd, nmr, mrow, yrow, rows, (trace_lane && b == 1) ? k : -1);
            if (++cslot == LN_RING) cslot = 0;
            fetch_next();                                          // refill the ring slot just consumed
            if (trace_lane && b == 1) BTRACE(19, k);
        }
        if (trace_lane) BTRACE(13, k);
    }
    bool trace_lane = false;
};

template <bool MLP, bool RES>
__global__ void __launch_bounds__(NT, 1)
swin_block_tail_kernel(const __grid_constant__ CUtensorMap mx64, const __grid_constant__ CUtensorMap mx32,
                       const __grid_constant__ CUtensorMap mw1_64, const __grid_constant__ CUtensorMap mw1_32,
                       const __grid_constant__ CUtensorMap mw2_64, const __grid_constant__ CUtensorMap mw2_32,
                       const __grid_constant__ Params p) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;
    uint8_t *sW1 = smem + p.off_w1;
    uint8_t *sW2 = smem + p.off_w2;
    float *s_ln = reinterpret_cast<float *>(smem + p.off_ln);
    float *s_b1 = reinterpret_cast<float *>(smem + p.off_misc);
    float *s_b2 = reinterpret_cast<float *>(smem + p.off_misc + p.off_b2);
    float *s_g = reinterpret_cast<float *>(smem + p.off_misc + p.off_g);
    float *s_be = reinterpret_cast<float *>(smem + p.off_misc + p.off_be);
    Bars *bars = reinterpret_cast<Bars *>(smem + p.off_misc + p.off_bar);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const soccdpt_block_tail_t &a = p.a;
    const int C = a.C;
    const int kc1 = p.kc64 + p.ktail;
    const int my_tiles = (int)blockIdx.x < p.tiles ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mx64);
        prefetch_tensormap(&mw2_64);
        if (MLP) prefetch_tensormap(&mw1_64);
        if (p.ktail) {
            prefetch_tensormap(&mx32);
            prefetch_tensormap(MLP ? &mw1_32 : &mw2_32);
        }
    }
    if (warp == 1 && lane == 0) {
        mbar_init(&bars->a_full, 1);
        mbar_init(&bars->a_empty, 1);
        for (int s = 0; s < MAX_SLOTS; ++s) {
            mbar_init(&bars->w1_full[s], 1);
            mbar_init(&bars->w1_empty[s], 1);
            mbar_init(&bars->w2_full[s], 1);
            mbar_init(&bars->w2_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->d1_full[s], 1);
            mbar_init(&bars->h_full[s], 256);       // the 8 GELU warps
            mbar_init(&bars->d2_full[s], 1);
            mbar_init(&bars->d2_empty[s], 128);     // one LayerNorm warp group
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bars->tmem_slot, 512);
    if (warp >= 4) {
        // per-channel constants -> shared memory (constants: staged before griddepcontrol.wait)
        const int et = threadIdx.x - 128;
        if (MLP) for (int i = et; i < a.HID; i += NT - 128) s_b1[i] = a.b1[i];
        for (int i = et; i < C; i += NT - 128) {
            s_b2[i] = a.b2[i];
            s_g[i] = a.gamma[i];
            s_be[i] = a.beta[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_slot;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (my_tiles > 0) {      // whole warp in uniform control flow, elect-predicated issue (tc_ptx.cuh)
            int s1 = 0, s2 = 0;                 // ring slots and their phase bits (no runtime divisions: these threads are latency bound)
            uint32_t ph1 = 0, ph2 = 0;
            auto load_w1 = [&](int j, int kc, int slot) {
                const bool tail = kc >= p.kc64;
                // resident: the tiles of a hidden chunk are packed (the 8 KB tail tile does not take a 16 KB slot)
                uint8_t *dst = RES ? sW1 + (uint32_t)j * p.w1_chunk + kc * 16384 : sW1 + slot * W1_SLOT;
                mbar_expect_tx_elect(&bars->w1_full[slot], tail ? W1_SLOT / 2 : W1_SLOT);
                tma_load_2d_elect(dst, tail ? &mw1_32 : &mw1_64, &bars->w1_full[slot], kc * 64, j * HC);
            };
            auto load_w2 = [&](int k0, bool tail, int slot) {
                mbar_expect_tx_elect(&bars->w2_full[slot], (uint32_t)C * (tail ? 64u : 128u));
                tma_load_2d_elect(sW2 + slot * p.w2_slot, tail ? &mw2_32 : &mw2_64, &bars->w2_full[slot], k0, 0);
            };
            if (RES) {
                // weights are constants: fetched before griddepcontrol.wait, under the previous kernel's tail
                if (MLP) {
                    for (int j = 0; j < p.nch; ++j) {
                        for (int kc = 0; kc < kc1; ++kc) load_w1(j, kc, j * kc1 + kc);
                        for (int h = 0; h < 2; ++h) load_w2(j * HC + h * 64, false, j * 2 + h);
                    }
                } else {
                    for (int kc = 0; kc < kc1; ++kc) load_w2(kc * 64, kc >= p.kc64, kc);
                }
            }
            soccdpt::pdl_wait();                // x / master are the previous kernels' outputs
            if (!MLP && !RES && p.stream_a) {
                // K-chunk pipeline: slot = activation chunk [128 x 64] + weight chunk [C x 64] (two boxes when C > 256)
                const uint32_t stage_bytes = 16384u + p.w2_slot;
                const uint32_t tx = 16384u + (uint32_t)C * 128u;
                for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                    for (int kc = 0; kc < p.kc64; ++kc) {
                        mbar_wait(&bars->w2_empty[s2], ph2 ^ 1u);
                        uint8_t *dst = sW2 + s2 * stage_bytes;
                        mbar_expect_tx_elect(&bars->w2_full[s2], tx);
                        tma_load_2d_elect(dst, &mx64, &bars->w2_full[s2], kc * 64, tile * BM);
                        tma_load_2d_elect(dst + 16384, &mw2_64, &bars->w2_full[s2], kc * 64, 0);
                        if (C > 256) tma_load_2d_elect(dst + 16384 + 32768, &mw2_32, &bars->w2_full[s2], kc * 64, 256);
                        if (++s2 == p.ns2) { s2 = 0; ph2 ^= 1u; }
                    }
                }
            }
            int it = 0;
            for (int tile = blockIdx.x; tile < ((!MLP && !RES && p.stream_a) ? 0 : p.tiles); tile += gridDim.x, ++it) {
                const int row0 = tile * BM;
                mbar_wait(&bars->a_empty, (uint32_t)(it & 1) ^ 1u);
                mbar_expect_tx_elect(&bars->a_full, p.a_bytes);
                for (int kc = 0; kc < p.kc64; ++kc) tma_load_2d_elect(sA + kc * 16384, &mx64, &bars->a_full, kc * 64, row0);
                if (p.ktail) tma_load_2d_elect(sA + p.kc64 * 16384, &mx32, &bars->a_full, p.kc64 * 64, row0);
                if (RES) continue;
                if (MLP) {
                    // fc1 weight tiles only: the fc2 ring has its own producer (warp 3) -- with one thread feeding both, a full
                    // fc2 ring (its slots free up only after GELU) kept the fc1 tiles of the NEXT chunk from being requested
                    for (int j = 0; j < p.nch; ++j) {
                        for (int kc = 0; kc < kc1; ++kc) {
                            mbar_wait(&bars->w1_empty[s1], ph1 ^ 1u);
                            load_w1(j, kc, s1);
                            if (++s1 == p.ns1) { s1 = 0; ph1 ^= 1u; }
                        }
                    }
                } else {
                    for (int kc = 0; kc < kc1; ++kc) {
                        mbar_wait(&bars->w2_empty[s2], ph2 ^ 1u);
                        load_w2(kc * 64, kc >= p.kc64, s2);
                        if (++s2 == p.ns2) { s2 = 0; ph2 ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ============================== second TMA producer: the fc2 weight ring (streamed two-GEMM mode) ==============================
        if (MLP && !RES && my_tiles > 0) {
            int s2 = 0;
            uint32_t ph2 = 0;
            for (int it = 0; it < my_tiles; ++it) {
                for (int j = 0; j < p.nch; ++j) {
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(&bars->w2_empty[s2], ph2 ^ 1u);
                        mbar_expect_tx_elect(&bars->w2_full[s2], (uint32_t)C * 128u);
                        tma_load_2d_elect(sW2 + s2 * p.w2_slot, &mw2_64, &bars->w2_full[s2], j * HC + h * 64, 0);
                        if (++s2 == p.ns2) { s2 = 0; ph2 ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (my_tiles > 0) {      // whole warp in uniform control flow, elect-predicated issue (tc_ptx.cuh)
            const uint32_t idesc1 = umma_idesc(HC), idesc2 = umma_idesc(C);
            const uint32_t hi128 = (uint32_t)(umma_desc(0, 128) >> 32), hi64 = (uint32_t)(umma_desc(0, 64) >> 32);
            const uint32_t a_lo0 = (uint32_t)umma_desc(smem_u32(sA), 128);           // low words: (address & 0x3FFFF) >> 4 | LBO
            const uint32_t w1_lo0 = (uint32_t)umma_desc(smem_u32(sW1), 128);
            const uint32_t w2_lo0 = (uint32_t)umma_desc(smem_u32(sW2), 128);
            const uint32_t w2_step = p.w2_slot >> 4;
            int s1 = 0, s2 = 0;                 // ring slots / phase bits of the streamed weight rings
            uint32_t ph1 = 0, ph2 = 0;
            const bool d2_two = p.d2_bufs == 2;
            if (MLP) {
                const int total = my_tiles * p.nch;
                const int nch = p.nch, kc64 = p.kc64, ktail = p.ktail;
                const uint32_t w1_chunk_lo = p.w1_chunk >> 4;
                // fc1 of hidden chunk j of tile iteration it -> D1[g & 1] (g = global chunk index)
                auto fc1 = [&](int g, int it, int j) {
                    if (j == 0) {
                        if (lane == 0) BTRACE(6, it);
                        mbar_wait(&bars->a_full, (uint32_t)(it & 1));
                        tc_fence_after();
                        if (lane == 0) BTRACE(7, it);
                    }
                    const uint32_t d1 = tmem + (uint32_t)((g & 1) * HC);
                    uint32_t a_lo = a_lo0, acc = 0u;
                    uint32_t b_lo = w1_lo0 + (uint32_t)j * w1_chunk_lo;                  // resident: packed tiles of chunk j
                    for (int kc = 0; kc < kc64; ++kc) {
                        int slot = 0;
                        if (RES) {
                            if (it == 0) { mbar_wait(&bars->w1_full[j * (kc64 + ktail) + kc], 0u); tc_fence_after(); }
                        } else {
                            slot = s1;
                            mbar_wait(&bars->w1_full[slot], ph1);
                            tc_fence_after();
                            if (++s1 == p.ns1) { s1 = 0; ph1 ^= 1u; }
                            b_lo = w1_lo0 + (uint32_t)slot * (W1_SLOT >> 4);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_ss_lo_elect(d1, a_lo + 2 * k, b_lo + 2 * k, hi128, idesc1, acc);
                            acc = 1u;
                        }
                        if (!RES) umma_commit_elect(&bars->w1_empty[slot]);
                        a_lo += 16384 >> 4;
                        b_lo += 16384 >> 4;
                    }
                    if (ktail) {
                        int slot = 0;
                        if (RES) {
                            if (it == 0) { mbar_wait(&bars->w1_full[j * (kc64 + 1) + kc64], 0u); tc_fence_after(); }
                        } else {
                            slot = s1;
                            mbar_wait(&bars->w1_full[slot], ph1);
                            tc_fence_after();
                            if (++s1 == p.ns1) { s1 = 0; ph1 ^= 1u; }
                            b_lo = w1_lo0 + (uint32_t)slot * (W1_SLOT >> 4);
                        }
                        umma_ss_lo_elect(d1, a_lo, b_lo, hi64, idesc1, acc);
                        umma_ss_lo_elect(d1, a_lo + 2, b_lo + 2, hi64, idesc1, 1u);
                        if (!RES) umma_commit_elect(&bars->w1_empty[slot]);
                    }
                    umma_commit_elect(&bars->d1_full[g & 1]);
                    if (j == nch - 1) umma_commit_elect(&bars->a_empty);     // the activation tile is free once these MMAs retire
                };
                // fc2 of chunk j: D2 (+)= H (bf16 pairs in the consumed D1 columns) x W2[:, chunk]
                auto fc2 = [&](int g, int it, int j) {
                    const int db2 = d2_two ? (it & 1) : 0;
                    if (j == 0) {
                        if (lane == 0) BTRACE(8, it);
                        mbar_wait(&bars->d2_empty[db2], (uint32_t)((d2_two ? it >> 1 : it) & 1) ^ 1u);
                        tc_fence_after();
                        if (lane == 0) BTRACE(9, it);
                    }
                    if (lane == 0) BTRACE(4, g);
                    mbar_wait(&bars->h_full[g & 1], (uint32_t)((g >> 1) & 1));
                    tc_fence_after();
                    if (lane == 0) BTRACE(5, g);
                    const uint32_t d2 = tmem + 256u + (uint32_t)(db2 * C);
                    const uint32_t h = tmem + (uint32_t)((g & 1) * HC);
                    uint32_t acc = j != 0 ? 1u : 0u;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        int slot;
                        if (RES) {
                            slot = j * 2 + hh;
                            if (it == 0) { mbar_wait(&bars->w2_full[slot], 0u); tc_fence_after(); }
                        } else {
                            slot = s2;
                            mbar_wait(&bars->w2_full[slot], ph2);
                            tc_fence_after();
                            if (++s2 == p.ns2) { s2 = 0; ph2 ^= 1u; }
                        }
                        const uint32_t b_lo = w2_lo0 + (uint32_t)slot * w2_step;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {   // hidden [hh*64 + k*16, +16) of the chunk = H columns hh*64 + k*8 .. +8
                            umma_ts_lo_elect(d2, h + (uint32_t)(hh * 64 + k * 8), b_lo + 2 * k, hi128, idesc2, acc);
                            acc = 1u;
                        }
                        if (!RES) umma_commit_elect(&bars->w2_empty[slot]);
                    }
                    if (j == nch - 1) umma_commit_elect(&bars->d2_full[db2]);
                };
                fc1(0, 0, 0);
                int it = 0, j = 0;                      // (tile iteration, chunk) of global chunk g
                for (int g = 0; g < total; ++g) {
                    int itn = it, jn = j + 1;           // ... and of g + 1
                    if (jn == nch) { jn = 0; ++itn; }
                    const bool more = g + 1 < total;
                    // the next chunk's fc1 goes first (the GELU warps never wait for the tensor pipe) -- across a tile boundary
                    // only if the next tile's activations have already landed, else after this chunk's fc2
                    const bool fc1_first = more && (jn != 0 || mbar_test(&bars->a_full, (uint32_t)(itn & 1)));
                    if (fc1_first) fc1(g + 1, itn, jn);
                    fc2(g, it, j);
                    if (more && !fc1_first) fc1(g + 1, itn, jn);
                    it = itn; j = jn;
                }
            } else if (!RES && p.stream_a) {
                // K-chunk pipeline (see the producer); N = C as one MMA (C <= 256) or two (256 + (C - 256) columns)
                const uint32_t stage_lo = (16384u + p.w2_slot) >> 4;
                const int n1 = C > 256 ? 256 : C;
                const uint32_t idesc_a = umma_idesc(n1), idesc_b = umma_idesc(C > 256 ? C - 256 : 16);
                for (int it = 0; it < my_tiles; ++it) {
                    const int db2 = d2_two ? (it & 1) : 0;
                    mbar_wait(&bars->d2_empty[db2], (uint32_t)((d2_two ? it >> 1 : it) & 1) ^ 1u);
                    tc_fence_after();
                    const uint32_t d2 = tmem + (uint32_t)(db2 * 256);
                    uint32_t acc = 0u;
                    for (int kc = 0; kc < p.kc64; ++kc) {
                        mbar_wait(&bars->w2_full[s2], ph2);
                        tc_fence_after();
                        const uint32_t a_lo = w2_lo0 + (uint32_t)s2 * stage_lo;
                        const uint32_t b_lo = a_lo + (16384u >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_ss_lo_elect(d2, a_lo + 2 * k, b_lo + 2 * k, hi128, idesc_a, acc);
                            if (C > 256) umma_ss_lo_elect(d2 + 256u, a_lo + 2 * k, b_lo + (32768u >> 4) + 2 * k, hi128, idesc_b, acc);
                            acc = 1u;
                        }
                        umma_commit_elect(&bars->w2_empty[s2]);
                        if (++s2 == p.ns2) { s2 = 0; ph2 ^= 1u; }
                    }
                    umma_commit_elect(&bars->d2_full[db2]);
                }
            } else {
                const int kc64 = p.kc64, ktail = p.ktail;
                for (int it = 0; it < my_tiles; ++it) {
                    const int db2 = it & 1;
                    mbar_wait(&bars->d2_empty[db2], (uint32_t)((it >> 1) & 1) ^ 1u);
                    mbar_wait(&bars->a_full, (uint32_t)(it & 1));
                    tc_fence_after();
                    const uint32_t d2 = tmem + (uint32_t)(db2 * 256);
                    uint32_t a_lo = a_lo0, acc = 0u;
                    for (int kc = 0; kc < kc64 + ktail; ++kc) {
                        int slot;
                        if (RES) {
                            slot = kc;
                            if (it == 0) { mbar_wait(&bars->w2_full[slot], 0u); tc_fence_after(); }
                        } else {
                            slot = s2;
                            mbar_wait(&bars->w2_full[slot], ph2);
                            tc_fence_after();
                            if (++s2 == p.ns2) { s2 = 0; ph2 ^= 1u; }
                        }
                        const uint32_t b_lo = w2_lo0 + (uint32_t)slot * w2_step;
                        if (kc < kc64) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_ss_lo_elect(d2, a_lo + 2 * k, b_lo + 2 * k, hi128, idesc2, acc);
                                acc = 1u;
                            }
                        } else {
                            umma_ss_lo_elect(d2, a_lo, b_lo, hi64, idesc2, acc);
                            umma_ss_lo_elect(d2, a_lo + 2, b_lo + 2, hi64, idesc2, 1u);
                        }
                        if (!RES) umma_commit_elect(&bars->w2_empty[slot]);
                        a_lo += 16384 >> 4;
                    }
                    umma_commit_elect(&bars->a_empty);
                    umma_commit_elect(&bars->d2_full[db2]);
                }
            }
        }
    } else if (warp >= 4) {
        const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);      // this warp's 32 TMEM lanes
        const int grp = (warp - 4) >> 2;                                         // 0, 1: warps 4..11; 2: warps 12..15
        soccdpt::pdl_wait();
        if (MLP && grp < 2) {
            // ============================== GELU: thread = (token row, 64 of the chunk's 128 columns) ==============================
            int g = 0;
            for (int it = 0; it < my_tiles; ++it) {
                for (int j = 0; j < p.nch; ++j, ++g) {
                    if (threadIdx.x == 128) BTRACE(0, g);
                    mbar_wait(&bars->d1_full[g & 1], (uint32_t)((g >> 1) & 1));
                    tc_fence_after();
                    if (threadIdx.x == 128) BTRACE(1, g);
                    const uint32_t base = t_lane + (uint32_t)((g & 1) * HC + grp * 64);
                    const float4 *bb = reinterpret_cast<const float4 *>(s_b1 + j * HC + grp * 64);
                    uint32_t r0[32], r1[32];
                    tmem_ld32_nowait(base, r0);
                    tmem_ld32_nowait(base + 32, r1);
                    tmem_ld_wait();
                    if (threadIdx.x == 128) BTRACE(2, g);
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {       // 4 values per step: one broadcast LDS.128 of the bias
                        const float4 b4 = bb[i];
                        pk[2 * i] = gelu_erf_bf16x2(fadd2(mk2(__uint_as_float(r0[4 * i]), __uint_as_float(r0[4 * i + 1])), mk2(b4.x, b4.y)));
                        pk[2 * i + 1] = gelu_erf_bf16x2(fadd2(mk2(__uint_as_float(r0[4 * i + 2]), __uint_as_float(r0[4 * i + 3])), mk2(b4.z, b4.w)));
                    }
                    // hidden columns [grp*64, +32) as 16 bf16 pairs over D1 columns this thread has consumed
                    tmem_st16(base, pk);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b4 = bb[8 + i];
                        pk[2 * i] = gelu_erf_bf16x2(fadd2(mk2(__uint_as_float(r1[4 * i]), __uint_as_float(r1[4 * i + 1])), mk2(b4.x, b4.y)));
                        pk[2 * i + 1] = gelu_erf_bf16x2(fadd2(mk2(__uint_as_float(r1[4 * i + 2]), __uint_as_float(r1[4 * i + 3])), mk2(b4.z, b4.w)));
                    }
                    tmem_st16(base + 16, pk);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&bars->h_full[g & 1]);
                    if (threadIdx.x == 128) BTRACE(3, g);
                }
            }
        } else if (MLP ? grp == 2 : (grp < 2 && (p.d2_bufs == 2 || grp == 0))) {
            // ============================== LayerNorm + residual: thread = token row ==============================
            LnWarp w;
            w.a = &a;
            w.b2 = s_b2; w.g = s_g; w.be = s_be;
            const int ln_idx = MLP ? warp - 12 : warp - 4;
            w.scr = s_ln + ln_idx * (LN_WARP_BYTES / 4);
            w.ring = w.scr + 1024;
            w.lane = lane;
            w.sub = warp & 3;
            w.nb = C >> 5;
            // two-GEMM mode: every tile of this CTA; one-GEMM mode: group `grp` takes the CTA's tiles it = grp, grp + 2, ...
            // (one-GEMM mode with a single D2 buffer, C > 256: group 0 alone takes every tile)
            const bool every = MLP || p.d2_bufs == 1;
            w.tile0 = every ? (int)blockIdx.x : (int)blockIdx.x + grp * (int)gridDim.x;
            w.tile_step = every ? (int)gridDim.x : 2 * (int)gridDim.x;
            w.n_tiles = every ? my_tiles : (my_tiles - grp + 1) / 2;
            w.fk = 0; w.fb = 0; w.fslot = 0; w.cslot = 0;
            w.lane_off = (lane >> 3) * C + (lane & 7) * 4;
            w.rstep = 4 * C;
            w.set_fetch_tile();
            w.trace_lane = lane == 0 && ln_idx == 0;
#pragma unroll
            for (int q = 0; q < LN_RING; ++q) w.fetch_next();
            for (int k = 0; k < w.n_tiles; ++k) {
                const int it = every ? k : 2 * k + grp;
                if (it + 2 >= my_tiles) soccdpt::pdl_trigger();       // the CTA's tail: let the next kernel's CTAs in
                const bool d2_two = p.d2_bufs == 2;
                const int db2 = every ? (d2_two ? (it & 1) : 0) : grp;
                const uint32_t d2_par = every ? (uint32_t)((d2_two ? it >> 1 : it) & 1) : (uint32_t)(k & 1);
                if (w.trace_lane) BTRACE(10, k);
                mbar_wait(&bars->d2_full[db2], d2_par);
                tc_fence_after();
                if (w.trace_lane) BTRACE(11, k);
                w.tile(k, t_lane + (MLP ? 256u + (uint32_t)(db2 * C) : (uint32_t)(db2 * 256)));
                tc_fence_before();
                mbar_arrive(&bars->d2_empty[db2]);
            }
            cp_async_wait<0>();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace

#ifdef SOCCDPT_TAIL_TRACE
extern "C" int soccdpt_block_tail_trace_read(unsigned long long *dst) {     // BT_EVENTS x BT_IDS stamps of the last launch (debug builds)
    SOCCDPT_CUDA(cudaDeviceSynchronize());
    SOCCDPT_CUDA(cudaMemcpyFromSymbol(dst, g_bt_trace, sizeof(unsigned long long) * BT_EVENTS * BT_IDS));
    return 0;
}
#endif

extern "C" int soccdpt_swin_block_tail_fwd(const soccdpt_block_tail_t *a, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(a && a->x && a->w2 && a->b2 && a->gamma && a->beta && a->master && a->y, "block_tail: NULL pointer");
    const bool mlp = a->w1 != nullptr;
    const int C = a->C, K1 = a->K1, HID = mlp ? a->HID : 0;
    SOCCDPT_REQUIRE(a->M >= 1 && a->M < (1ll << 31) - BM, "block_tail: bad row count %lld", a->M);
    // one-GEMM mode with wide rows (C > 256: stage 2 of swin2_tiny / swin2_base) or a long K (fc2: K1 = 4C): activations and
    // weights stream through shared memory by K chunk instead of the activation tile being resident
    const bool stream_a = !mlp && (C > 256 || K1 > 512);
    SOCCDPT_REQUIRE(C >= 32 && C % 32 == 0 && C <= (stream_a ? 512 : 256),
                    "block_tail: C must be a multiple of 32, <= 256 (two-GEMM mode) / <= 512 (one-GEMM mode) (got %d)", C);
    SOCCDPT_REQUIRE(K1 >= 32 && K1 % 32 == 0 && K1 <= (stream_a ? 4096 : 512) && (!stream_a || K1 % 64 == 0),
                    "block_tail: unsupported K1 = %d", K1);
    if (mlp) {
        SOCCDPT_REQUIRE(a->b1 != nullptr, "block_tail: b1 is NULL");
        SOCCDPT_REQUIRE(HID >= 2 * HC && HID % HC == 0 && HID <= 4096, "block_tail: HID must be a multiple of 128 (got %d)", HID);
    }
    const void *ptrs[] = {a->x, a->w1, a->w2, a->master, a->y};
    for (const void *q : ptrs) SOCCDPT_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "block_tail: pointers must be 16-byte aligned");
    SOCCDPT_REQUIRE(tc::encode_fn() != nullptr, "cuTensorMapEncodeTiled is not available from the driver");

    Params p{};
    p.a = *a;
    p.tiles = (int)((a->M + BM - 1) / BM);
    p.kc64 = K1 / 64;
    p.ktail = (K1 % 64) ? 1 : 0;
    p.nch = HID / HC;
    const int kc1 = p.kc64 + p.ktail;
    p.a_bytes = (uint32_t)(p.kc64 * 16384 + p.ktail * 8192);
    p.w2_slot = (uint32_t)((C * 128 + 1023) / 1024 * 1024);
    p.stream_a = stream_a ? 1 : 0;
    // TMEM: D1 = columns 0..255 (two-GEMM mode), D2 from column 256; one-GEMM mode: D2 buffers at columns 0 / 256
    p.d2_bufs = mlp ? (2 * C <= 256 ? 2 : 1) : (C <= 256 ? 2 : 1);
    // misc block: b1 | b2 | gamma | beta | barriers
    uint32_t m = 0;
    m += (uint32_t)((HID * 4 + 15) / 16 * 16);
    p.off_b2 = m; m += (uint32_t)C * 4;
    p.off_g = m; m += (uint32_t)C * 4;
    p.off_be = m; m += (uint32_t)C * 4;
    m = (m + 15) / 16 * 16;
    p.off_bar = m; m += (uint32_t)sizeof(Bars);
    const uint32_t misc_bytes = (m + 1023) / 1024 * 1024;
    // LayerNorm warps: 4 (two-GEMM mode; one-GEMM mode with a single D2 buffer) or 8
    const uint32_t ln_bytes = (uint32_t)((mlp || p.d2_bufs == 1) ? 4 : 8) * LN_WARP_BYTES;
    const uint32_t a_slot = stream_a ? 0u : (p.a_bytes + 1023) / 1024 * 1024;
    const long long budget = SMEM_MAX - 1024 /*alignment slack*/ - (long long)misc_bytes - ln_bytes - a_slot;
    const int w1_tiles = mlp ? p.nch * kc1 : 0;
    const int w2_tiles = mlp ? p.nch * 2 : kc1;
    p.w1_chunk = (uint32_t)(p.kc64 * 16384 + p.ktail * 8192);
    uint32_t w1_bytes = 0, w2_bytes = 0;
    if (stream_a) {
        const long long stage = 16384 + (long long)p.w2_slot;
        p.resident = 0;
        p.ns1 = 0;
        p.ns2 = (int)(budget / stage);
        if (p.ns2 > MAX_SLOTS) p.ns2 = MAX_SLOTS;
        SOCCDPT_REQUIRE(p.ns2 >= 2, "block_tail: shapes do not fit shared memory (C=%d K1=%d)", C, K1);
        w2_bytes = (uint32_t)(p.ns2 * stage);
    } else if ((long long)p.nch * p.w1_chunk + (long long)w2_tiles * p.w2_slot <= budget && w1_tiles <= MAX_SLOTS && w2_tiles <= MAX_SLOTS) {
        p.resident = 1;
        p.ns1 = w1_tiles;
        p.ns2 = w2_tiles;
        w1_bytes = (uint32_t)p.nch * p.w1_chunk;
        w2_bytes = (uint32_t)p.ns2 * p.w2_slot;
    } else {
        p.resident = 0;
        // ring 2 first (one chunk = two tiles in two-GEMM mode; two chunks when they fit), the rest goes to ring 1
        p.ns2 = mlp ? 2 : 1;
        long long rest = budget - (long long)p.ns2 * p.w2_slot;
        if (mlp) {
            SOCCDPT_REQUIRE(rest >= 2ll * W1_SLOT, "block_tail: shapes do not fit shared memory (C=%d K1=%d)", C, K1);
            if (rest >= (long long)(2 * kc1) * W1_SLOT + 2ll * p.w2_slot) { p.ns2 = 4; rest -= 2ll * p.w2_slot; }
            else if (rest >= (long long)(kc1 + 1) * W1_SLOT + (long long)p.w2_slot) { p.ns2 = 3; rest -= p.w2_slot; }
            p.ns1 = (int)(rest / W1_SLOT);
            if (p.ns1 > MAX_SLOTS) p.ns1 = MAX_SLOTS;
        } else {
            SOCCDPT_REQUIRE(rest >= 0, "block_tail: shapes do not fit shared memory (C=%d K1=%d)", C, K1);
            while (p.ns2 < MAX_SLOTS && rest >= (long long)p.w2_slot) { ++p.ns2; rest -= p.w2_slot; }
            p.ns1 = 0;
        }
        w1_bytes = (uint32_t)p.ns1 * W1_SLOT;
        w2_bytes = (uint32_t)p.ns2 * p.w2_slot;
    }
    p.off_w1 = a_slot;
    p.off_w2 = p.off_w1 + w1_bytes;
    p.off_ln = p.off_w2 + w2_bytes;
    p.off_misc = p.off_ln + ln_bytes;
    const size_t smem_bytes = (size_t)p.off_misc + misc_bytes + 1024;
    SOCCDPT_REQUIRE(smem_bytes <= (size_t)SMEM_MAX, "block_tail: %zu bytes of shared memory needed (C=%d K1=%d HID=%d)", smem_bytes, C, K1, HID);

    CUtensorMap mx64, mx32, mw1_64, mw1_32, mw2_64, mw2_32;
    CUresult r = tc::encode_2d_bf16(&mx64, a->x, (uint64_t)a->M, (uint64_t)K1, 64, BM, false);
    SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "block_tail: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
    mx32 = mx64; mw1_64 = mx64; mw1_32 = mx64; mw2_32 = mx64;
    if (p.ktail) {
        r = tc::encode_2d_bf16(&mx32, a->x, (uint64_t)a->M, (uint64_t)K1, 32, BM, false);
        SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "block_tail: cuTensorMapEncodeTiled(x tail) failed with %d", (int)r);
    }
    if (mlp) {
        r = tc::encode_2d_bf16(&mw1_64, a->w1, (uint64_t)HID, (uint64_t)K1, 64, HC, true);
        SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "block_tail: cuTensorMapEncodeTiled(w1) failed with %d", (int)r);
        if (p.ktail) {
            r = tc::encode_2d_bf16(&mw1_32, a->w1, (uint64_t)HID, (uint64_t)K1, 32, HC, true);
            SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "block_tail: cuTensorMapEncodeTiled(w1 tail) failed with %d", (int)r);
        }
        r = tc::encode_2d_bf16(&mw2_64, a->w2, (uint64_t)C, (uint64_t)HID, 64, (uint32_t)C, true);
        SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "block_tail: cuTensorMapEncodeTiled(w2) failed with %d", (int)r);
    } else if (stream_a) {
        r = tc::encode_2d_bf16(&mw2_64, a->w2, (uint64_t)C, (uint64_t)K1, 64, (uint32_t)(C > 256 ? 256 : C), true);
        SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "block_tail: cuTensorMapEncodeTiled(w2) failed with %d", (int)r);
        if (C > 256) {      // second N part of a weight chunk: rows [256, C)
            r = tc::encode_2d_bf16(&mw2_32, a->w2, (uint64_t)C, (uint64_t)K1, 64, (uint32_t)(C - 256), true);
            SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "block_tail: cuTensorMapEncodeTiled(w2 part 2) failed with %d", (int)r);
        }
    } else {
        r = tc::encode_2d_bf16(&mw2_64, a->w2, (uint64_t)C, (uint64_t)K1, 64, (uint32_t)C, true);
        SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "block_tail: cuTensorMapEncodeTiled(w2) failed with %d", (int)r);
        if (p.ktail) {
            r = tc::encode_2d_bf16(&mw2_32, a->w2, (uint64_t)C, (uint64_t)K1, 32, (uint32_t)C, true);
            SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "block_tail: cuTensorMapEncodeTiled(w2 tail) failed with %d", (int)r);
        }
    }

    const int sms = soccdpt::sm_count();
    const int grid = p.tiles < sms ? p.tiles : sms;
    cudaStream_t st = soccdpt::as_stream(stream);
#define SOCC_TAIL(M_, R_)                                                                                                     \
    do {                                                                                                                      \
        static soccdpt::SmemAttr configured;                                                                                  \
        if (configured.need(SMEM_MAX))                                                                                        \
            SOCCDPT_CUDA(cudaFuncSetAttribute(swin_block_tail_kernel<M_, R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX)); \
        SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_CONV, swin_block_tail_kernel<M_, R_>, dim3(grid), dim3(NT), smem_bytes, st, mx64, \
                                         mx32, mw1_64, mw1_32, mw2_64, mw2_32, p));                                           \
    } while (0)
    if (mlp) {
        if (p.resident) SOCC_TAIL(true, true);
        else SOCC_TAIL(true, false);
    } else {
        if (p.resident) SOCC_TAIL(false, true);
        else SOCC_TAIL(false, false);
    }
#undef SOCC_TAIL
    return soccdpt::check_launch("swin_block_tail_kernel");
}
