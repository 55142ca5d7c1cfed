// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (sm_100a only).
//
//   D[pixel, cout] = sum_{tap, c} X[pixel + tap, c] * Wt[cout, tap, c]      bf16 x bf16 -> fp32 (TMEM)
//
// * A operand (activations, NHWC): loaded by TMA as a 4-D box {64 ch, BW, BH, BN} whose coordinates
//   are shifted by the filter tap -- the im2col never exists, and TMA's out-of-bounds zero fill IS the
//   convolution's zero padding (and the K / M tails).  128 pixels x 64 channels land in shared
//   memory in the 128B-swizzled K-major layout tcgen05.mma consumes directly.
// * B operand (weights [Cout][tap][Cin]): 3-D box {64, 1, BLOCK_N}, same swizzle.
// * tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N (16..256), K=16, issued by ONE thread;
//   accumulators live in TMEM (2 stages x 256 columns) so the epilogue of tile i overlaps the
//   MMAs of tile i+1.
// * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//   warps 4-11 = epilogue: two warpgroups, each draining half of the accumulator columns of all 128 rows
//   (tcgen05.ld -> bias / ReLU / GELU / residual adds / bf16 store, or the fused narrow projection
//   32->1 / 256->3 of the depth / seg heads).  The epilogue is template-specialised on <activation, mode>:
//   with per-element runtime branches it ran ~170 warp-instructions per 16 columns and was as long as the
//   whole K=2304 mainloop.
// * persistent: grid = min(tiles, SMs); 4-stage smem ring (A 16 KB + B <=32 KB per stage).
// * RESIDENT N BLOCK for 1x1 / linear layers whose N block of weights ([K/64][BLOCK_N][64] bf16) fits behind a >= 3-slot A ring
//   (<= 144 KB): the grid is rounded down to a multiple of the number of N blocks, so `tile += gridDim.x` keeps a CTA on ONE
//   N block; its weights are fetched once per CTA instead of once per tile (S2 qkv: 245 -> 98 KB arriving per tile) and the
//   whole ring (up to 24 x 16 KB) prefetches activations.
// * HALO variant for 3x3 convs on 128-pixel row segments: the kernel is bound by bytes ARRIVING per SM
//   (~64 B/clk), and re-loading the 16 KB activation tile for each of the 9 taps is most of them.  Instead one
//   TMA box {64 ch, 130 px, 3 rows} per channel block is loaded once and the nine taps are issued from it with
//   row-offset shared-memory descriptors (matrix-base-offset field) -> 3.4x fewer activation bytes.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace soccdpt {
int validate_conv(const soccdpt_conv_t *c);
}

// -DSOCCDPT_CONV_TRACE (debug builds only, tools/trace_conv.py): CTA 0 logs SM-clock stamps of its first tiles -- the MMA warp's and the
// first epilogue thread's hand-offs -- into a global buffer read back by soccdpt_conv_trace_read()
#ifdef SOCCDPT_CONV_TRACE
constexpr int CT_TILES = 24, CT_EVENTS = 32;
__device__ unsigned long long g_conv_trace[CT_TILES * CT_EVENTS];
#define CTRACE(tile_it, ev)                                                                                          \
    do {                                                                                                             \
        if (blockIdx.x == 0 && (tile_it) < CT_TILES) g_conv_trace[(tile_it) * CT_EVENTS + (ev)] = (unsigned long long)clock64(); \
    } while (0)
#else
#define CTRACE(tile_it, ev) do { } while (0)
#endif

namespace {

using bf16 = __nv_bfloat16;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;            // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int B_STAGE_BYTES = 256 * BLOCK_K * 2;       // 32 KB (BLOCK_N <= 256)
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 384;          // 4 control warps + 8 epilogue warps
constexpr int EPI_THREADS = 256;          // two epilogue warpgroups, each owns half of the tile's columns
constexpr int BAR_FREE = 5, BAR_STAGED = 7;   // named barriers (+ warpgroup): staging tile free / staged, 128 epilogue threads + the store warp
// dynamic smem: ring + epilogue constants (bias 256 f32 + proj 4x256 f32 + proj bias) + barriers
constexpr int EPI_CONST_BYTES = (256 + 4 * 256 + 4 + 128 * 4) * 4;   // bias, proj weights, proj bias, proj partials
// HALO variant (3x3, W % 128 == 0): one TMA box {64 ch, 130 px, 3 rows} per channel block serves all 9 taps
constexpr int HALO_W = BLOCK_M + 2;                      // 130 pixels per halo row
constexpr int HALO_ROWS = 3 * HALO_W;                    // 390 rows of 128 B
constexpr int HALO_BYTES = HALO_ROWS * BLOCK_K * 2;      // 49920 B moved by TMA
constexpr int HALO_STAGE_BYTES = 50 * 1024;              // 1024-aligned slot
constexpr int HALO_A_STAGES = 2;
// ROW-PAIR variant of HALO for N <= 128 (Params::m2): one box {64 ch, 130 px, 4 rows} feeds TWO M tiles (image rows h0 and h0 + 1),
// each weight tile is used by both -> half the weight bytes and 2/3 of the activation bytes per MMA; the two accumulators take
// columns [0,128) and [128,256) of the TMEM stage and one epilogue warpgroup each.
constexpr int HALO2_BYTES = 4 * HALO_W * BLOCK_K * 2;    // 66560 B = 65 x 1024: the slot size as well
constexpr int HALO_B_BYTES = 3 * B_STAGE_BYTES;          // 96 KB weight ring, split into b_stages slots of one tap tile
constexpr int MAX_B_STAGES = 24;
constexpr int RING_BYTES_PLAIN = STAGES * STAGE_BYTES;                                           // 196608
constexpr int RING_BYTES_HALO = HALO_A_STAGES * HALO_STAGE_BYTES + HALO_B_BYTES;                  // 200704
constexpr int RING_BYTES = RING_BYTES_HALO > RING_BYTES_PLAIN ? RING_BYTES_HALO : RING_BYTES_PLAIN;
// output staging for the TMA-store epilogue: one [128 rows][32 cols] bf16 tile (64-byte rows, SWIZZLE_64B) per warpgroup
constexpr int STAGING_BYTES = 128 * 64;
constexpr int SMEM_BYTES = RING_BYTES + 2 * STAGING_BYTES + EPI_CONST_BYTES + 512 /*barriers*/ + 1024 /*alignment slack*/;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params {
    // tile geometry
    int BW, BH, BN;            // box of output pixels handled by one M tile (BW*BH*BN <= 128)
    int tiles_w, tiles_h, tiles_n, n_blocks, total_tiles;
    int block_n;               // output channels per tile
    int k_blocks_per_tap;      // ceil(Cin / 64)
    int pad;                   // zero padding before the first row / column
    int stride, Ho, Wo;        // output grid = ceil(input / stride)
    int b_stages;              // HALO: depth of the weight ring (96 KB / bytes per tap tile, <= 24)
    int alt_split;             // epilogue: alternate the column split of odd chunk counts (SOCCDPT_CONV_ALT=0 switches it off)
    int m2;                    // HALO: two M tiles (rows h0, h0 + 1) per pass, see HALO2_BYTES
    int halo_slot, halo_bytes; // HALO: bytes between the two halo slots / moved per halo box
    int b_resident;            // HALO: the whole weight tensor (<= 96 KB) is loaded once per CTA and stays in smem
    int nres;                  // plain 1x1 path: the grid is a multiple of n_blocks, so a CTA keeps ONE N block for all its tiles and
                               // that block's weights ([k_blocks][block_n][64], <= 144 KB) stay resident; the ring holds A tiles only
    int a_stages;              // nres: depth of the A-only ring (16 KB slots in front of the resident weights)
    uint32_t a_bytes, b_bytes; // TMA transaction bytes per stage
    soccdpt_conv_t c;
};

// (N block, tile column, tile row, tile image) of tile = first + k * step, advanced WITHOUT a division per tile: the five
// runtime divisions / modulos of the decomposition were ~1000 cycles of dependent integer code at the top of every epilogue
// tile (tools/trace_conv.py) -- as long as a whole 32-column chunk of a narrow-K layer
struct TileIter {
    int nb, tw, th, tn, s_nb, s_tw, s_th, s_tn;
    __device__ __forceinline__ void init(const Params &p, int first, int step) {
        nb = first % p.n_blocks; int mt = first / p.n_blocks;
        tw = mt % p.tiles_w; mt /= p.tiles_w;
        th = mt % p.tiles_h; tn = mt / p.tiles_h;
        s_nb = step % p.n_blocks; mt = step / p.n_blocks;
        s_tw = mt % p.tiles_w; mt /= p.tiles_w;
        s_th = mt % p.tiles_h; s_tn = mt / p.tiles_h;
    }
    __device__ __forceinline__ void next(const Params &p) {
        nb += s_nb; int c = nb >= p.n_blocks ? 1 : 0; nb -= c ? p.n_blocks : 0;
        tw += s_tw + c; c = tw >= p.tiles_w ? 1 : 0; tw -= c ? p.tiles_w : 0;
        th += s_th + c; c = th >= p.tiles_h ? 1 : 0; th -= c ? p.tiles_h : 0;
        tn += s_tn + c;
    }
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// producer / consumer named barriers (PTX ISA, bar.arrive + bar.sync): arriving threads do not wait
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// The producer and issuer warps run their loops with all 32 lanes in uniform control flow; every asynchronous instruction is
// predicated on elect.sync inside its asm block (lane 0 of the full warp, every time).  Under `if (lane == 0)` the compiler must
// assume a divergent warp and wraps each tcgen05.mma / cp.async.bulk.tensor in an ELECT ... BRA.U.ANY loop over the active lanes:
// ~10 dependent SASS instructions, measured ~88 cycles per MMA on the issuing thread -- the "narrow-N floor" of round 1.  In
// convergent code the same source is a bare UTCHMMA / UTMALDG with uniform-register operands (tools/microbench/mma_latency.cu:
// SS N = 64 48 cycles, N = 128 63 cycles, TS N = 32 17 cycles per MMA).
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile(
        "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
        "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
        "@e cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n}\n"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
        "@e cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n}\n"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows are 128 B,
// 8-row groups are 1024 B apart (SBO), version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                            // version
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
// Same, for a tile that starts at an arbitrary 128-byte row of a swizzled buffer (tap offsets into the halo).
// MEASURED on B200: the tensor core applies the 128B swizzle on ABSOLUTE shared-memory address bits
// (bits [4,7) ^= bits [7,10)), exactly like TMA when it wrote the data, so a row-offset start address needs
// nothing else; setting the "matrix base offset" field (bits [49,52)) to (addr >> 7) & 7 gives WRONG results.
__device__ __forceinline__ uint64_t umma_desc_rows(uint32_t smem_addr) { return umma_desc(smem_addr); }
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Lean issue form for the single MMA thread: the two descriptors share their high words (constant per layout),
// only the 32-bit low words (start address >> 4) move.  ~25 instructions per MMA in the issuing thread cost
// ~150 cycles -- invisible behind a 128-cycle N=256 MMA, dominant for N=32 (16 cycles of tensor work).
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile(
        "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// exact-erf GELU (timm Mlp act) with erf from Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the
// bf16 output rounding): 1 MUFU.RCP + 1 MUFU.EX2 + ~12 FMAs instead of ~35 instructions for erff().
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
    const float erf_abs = fmaf(-poly * t, e, 1.0f);           // erf(|x|/sqrt2)
    return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// ------------------------------------------------------------------ the kernel
// v[8*N] += bf16 values at p (N x 16 bytes)
template <int N, int LEN>
__device__ __forceinline__ void add_bf16_impl(float (&v)[LEN], const uint4 *p) {
#pragma unroll
    for (int h = 0; h < N; ++h) {
        const uint4 u = p[h];
        const __nv_bfloat162 *b2 = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(b2[k]);
            v[h * 8 + 2 * k] += f.x;
            v[h * 8 + 2 * k + 1] += f.y;
        }
    }
}
template <int N, int LEN>
__device__ __forceinline__ void add_bf16(float (&v)[LEN], const uint4 *p) { add_bf16_impl<N, LEN>(v, p); }
// v[8*N] += w * bf16 values at p
template <int N, int LEN>
__device__ __forceinline__ void fma_bf16(float (&v)[LEN], const uint4 *p, float w) {
#pragma unroll
    for (int h = 0; h < N; ++h) {
        const uint4 u = __ldg(p + h);
        const __nv_bfloat162 *b2 = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(b2[k]);
            v[h * 8 + 2 * k] = fmaf(w, f.x, v[h * 8 + 2 * k]);
            v[h * 8 + 2 * k + 1] = fmaf(w, f.y, v[h * 8 + 2 * k + 1]);
        }
    }
}
template <int N, bool RELU, int LEN>
__device__ __forceinline__ void store_bf16(const float (&v)[LEN], uint4 *p) {
#pragma unroll
    for (int h = 0; h < N; ++h) {
        uint4 u;
        __nv_bfloat162 *b2 = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = RELU ? fmaxf(v[h * 8 + 2 * k], 0.0f) : v[h * 8 + 2 * k];
            const float b = RELU ? fmaxf(v[h * 8 + 2 * k + 1], 0.0f) : v[h * 8 + 2 * k + 1];
            b2[k] = __floats2bfloat162_rn(a, b);
        }
        p[h] = u;
    }
}

// Epilogue store of one 32-column chunk through shared memory + TMA: thread `row` writes its 64-byte row into the
// 64B-swizzled staging tile (conflict-free), the warpgroup synchronises, one thread issues the bulk tensor store
// (coalesced 64-byte row segments, clipped at the tensor bounds by the TMA unit).
template <bool RELU>
__device__ __forceinline__ void stage_row32(uint8_t *stage, int row, const float (&v)[32]) {
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
        uint4 u;
        __nv_bfloat162 *b2 = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = RELU ? fmaxf(v[c4 * 8 + 2 * k], 0.0f) : v[c4 * 8 + 2 * k];
            const float b = RELU ? fmaxf(v[c4 * 8 + 2 * k + 1], 0.0f) : v[c4 * 8 + 2 * k + 1];
            b2[k] = __floats2bfloat162_rn(a, b);
        }
        *reinterpret_cast<uint4 *>(stage + row * 64 + ((c4 ^ ((row >> 1) & 3)) << 4)) = u;
    }
}

// MODE 0: y = act(acc + bias)            (no residual, no ReLU copy, no projection)
// MODE 1: general: optional res1 / res2 / y / y_relu
// MODE 2: fused projection only (proj_out), no y
// HALO : 3x3 conv whose M tile is 128 consecutive pixels of one image row -> halo reuse (see constants above)
template <int ACT, int MODE, bool HALO>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_yr,
                    const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms, as an OFFSET on the __shared__ symbol: a uintptr_t round trip
    // makes the compiler lose the address space and emit generic LD/ST for every shared-memory access
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *s_stage = smem + RING_BYTES;                 // 2 x 8 KB, 1024-byte aligned
    float *s_bias = reinterpret_cast<float *>(smem + RING_BYTES + 2 * STAGING_BYTES);
    float *s_projw = s_bias + 256;
    float *s_projb = s_projw + 4 * 256;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_projb + 4 + 128 * 4);   // after the [128][4] projection partials
    uint64_t *empty = full + MAX_B_STAGES;
    uint64_t *acc_full = empty + MAX_B_STAGES;
    uint64_t *acc_empty = acc_full + ACC_STAGES;
    uint64_t *a_full = acc_empty + ACC_STAGES;           // HALO: ring of halo tiles (full/empty above = weight ring)
    uint64_t *a_empty = a_full + HALO_A_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_empty + HALO_A_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const soccdpt_conv_t &c = p.c;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < MAX_B_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_THREADS); }
        for (int s = 0; s < HALO_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barriers, TMEM, tensor-map fetch) overlapped the previous kernel's tail; its outputs from here on
    soccdpt::pdl_wait();

    const int taps = c.KH * c.KW;
    const int k_blocks = taps * p.k_blocks_per_tap;

    if (warp == 0) {
        // ===================== TMA producer (whole warp, elect-predicated issue) =====================
        {
            int stage = 0, a_stage = 0;
            uint32_t phase = 0, a_phase = 0;
            if (HALO && p.b_resident && blockIdx.x < p.total_tiles) {
                // narrow layers (e.g. 128->32: 72 KB of weights): fetch every (channel block, tap) tile once
                uint8_t *b_ring = smem + HALO_A_STAGES * p.halo_slot;
                mbar_expect_tx(&full[0], (uint32_t)(p.k_blocks_per_tap * 9) * p.b_bytes);
                for (int cb = 0; cb < p.k_blocks_per_tap; ++cb)
                    for (int tap = 0; tap < 9; ++tap)
                        tma_load_3d(b_ring + (cb * 9 + tap) * p.b_bytes, &map_b, &full[0], cb * BLOCK_K, tap, 0);
            }
            if (!HALO && p.nres && blockIdx.x < p.total_tiles) {
                // this CTA's N block never changes (grid % n_blocks == 0): fetch its weights once, behind the A ring
                uint8_t *b_res = smem + p.a_stages * A_STAGE_BYTES;
                mbar_expect_tx(&a_full[0], (uint32_t)k_blocks * p.b_bytes);
                for (int kb = 0; kb < k_blocks; ++kb)
                    tma_load_3d(b_res + kb * p.b_bytes, &map_b, &a_full[0], kb * BLOCK_K, 0, (blockIdx.x % p.n_blocks) * p.block_n);
            }
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nb = tile % p.n_blocks, mt = tile / p.n_blocks;
                const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
                const int w0 = tw * p.BW, h0 = th * (HALO && p.m2 ? 2 : p.BH), n0 = tn * p.BN;
                if (!HALO && p.nres) {
                    for (int kb = 0; kb < k_blocks; ++kb) {          // 1x1: k block == channel block, tap (0, 0)
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], p.a_bytes);
                        tma_load_4d(smem + stage * A_STAGE_BYTES, &map_a, &full[stage], kb * BLOCK_K, w0 - p.pad, h0 - p.pad, n0);
                        if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
                    }
                    continue;
                }
                if (HALO) {
                    // channel-block outer, tap inner: one halo tile (A ring) feeds nine weight tiles (B ring)
                    uint8_t *b_ring = smem + HALO_A_STAGES * p.halo_slot;
                    for (int cb = 0; cb < p.k_blocks_per_tap; ++cb) {
                        mbar_wait(&a_empty[a_stage], a_phase ^ 1);
                        mbar_expect_tx(&a_full[a_stage], (uint32_t)p.halo_bytes);
                        tma_load_4d(smem + a_stage * p.halo_slot, &map_a, &a_full[a_stage], cb * BLOCK_K, w0 - 1, h0 - 1, n0);
                        if (++a_stage == HALO_A_STAGES) { a_stage = 0; a_phase ^= 1; }
                        if (p.b_resident) continue;
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(&empty[stage], phase ^ 1);
                            mbar_expect_tx(&full[stage], p.b_bytes);
                            tma_load_3d(b_ring + stage * p.b_bytes, &map_b, &full[stage], cb * BLOCK_K, tap, nb * p.block_n);
                            if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
                        }
                    }
                    continue;
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int tap = kb / p.k_blocks_per_tap, cb = kb - tap * p.k_blocks_per_tap;
                    const int kh = tap / c.KW, kw = tap - kh * c.KW;
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], p.a_bytes + p.b_bytes);
                    uint8_t *sa = smem + stage * STAGE_BYTES;
                    tma_load_4d(sa, &map_a, &full[stage], cb * BLOCK_K, w0 * p.stride + kw - p.pad, h0 * p.stride + kh - p.pad, n0);
                    tma_load_3d(sa + A_STAGE_BYTES, &map_b, &full[stage], cb * BLOCK_K, tap, nb * p.block_n);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp in uniform control flow, elect-predicated issue) =====================
        {
            const uint32_t idesc = umma_idesc(p.block_n);
            int stage = 0, a_stage = 0;
            uint32_t phase = 0, a_phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            if (HALO && p.b_resident && blockIdx.x < p.total_tiles) {
                mbar_wait(&full[0], 0);                      // resident weights have landed
                tc_fence_after();
            }
            if (!HALO && p.nres && blockIdx.x < p.total_tiles) {
                mbar_wait(&a_full[0], 0);                    // resident weights of this CTA's N block have landed
                tc_fence_after();
            }
            int mit = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++mit) {
                if (lane == 0) CTRACE(mit, 0);
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                if (lane == 0) CTRACE(mit, 1);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                if (!HALO && p.nres) {
                    const uint32_t desc_hi = (uint32_t)(umma_desc(0) >> 32);
                    uint32_t b_lo = (uint32_t)umma_desc(smem_u32(smem + p.a_stages * A_STAGE_BYTES));
                    const uint32_t b_step = p.b_bytes >> 4;
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        const int rem = c.Cin - kb * BLOCK_K;
                        const int ksteps = rem >= BLOCK_K ? BLOCK_K / UMMA_K : (rem + UMMA_K - 1) / UMMA_K;
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_lo = (uint32_t)umma_desc(smem_u32(smem + stage * A_STAGE_BYTES));
                        if (ksteps == 4) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        } else {
                            for (int k = 0; k < ksteps; ++k) umma_f16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(&empty[stage]);
                        if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
                        b_lo += b_step;
                        if (lane == 0 && kb == 0) CTRACE(mit, 2);
                    }
                    umma_commit(&acc_full[acc]);
                    if (lane == 0) CTRACE(mit, 3);
                    if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
                    continue;
                }
                if (HALO) {
                    const uint32_t b_ring = smem_u32(smem + HALO_A_STAGES * p.halo_slot);
                    const uint32_t desc_hi = (uint32_t)(umma_desc(0) >> 32);
                    uint32_t first = 0u;                     // becomes 1 after the first MMA of the tile
                    for (int cb = 0; cb < p.k_blocks_per_tap; ++cb) {
                        const int rem = c.Cin - cb * BLOCK_K;
                        const int ksteps = rem >= BLOCK_K ? BLOCK_K / UMMA_K : (rem + UMMA_K - 1) / UMMA_K;
                        mbar_wait(&a_full[a_stage], a_phase);
                        tc_fence_after();
                        // descriptor low word of the halo tile; a tap is a constant row offset: output pixel bw of
                        // the tile reads halo row kh, pixel bw + kw
                        const uint32_t a_lo = (uint32_t)umma_desc(smem_u32(smem + a_stage * p.halo_slot));
                        if (p.m2) {
                            // row pair: every weight tile multiplies the halo rows of BOTH output rows (second tile: one halo row lower)
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint32_t at = a_lo + (uint32_t)(((tap / 3) * HALO_W + (tap % 3)) * 8);
                                mbar_wait(&full[stage], phase);
                                tc_fence_after();
                                const uint32_t b_lo = (uint32_t)umma_desc(b_ring + stage * p.b_bytes);
                                if (ksteps == 4) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) umma_f16_lo(d_tmem, at + 2 * k, b_lo + 2 * k, desc_hi, idesc, k == 0 ? first : 1u);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) umma_f16_lo(d_tmem + 128u, at + HALO_W * 8 + 2 * k, b_lo + 2 * k, desc_hi, idesc, k == 0 ? first : 1u);
                                } else {
                                    for (int k = 0; k < ksteps; ++k) umma_f16_lo(d_tmem, at + 2 * k, b_lo + 2 * k, desc_hi, idesc, k == 0 ? first : 1u);
                                    for (int k = 0; k < ksteps; ++k) umma_f16_lo(d_tmem + 128u, at + HALO_W * 8 + 2 * k, b_lo + 2 * k, desc_hi, idesc, k == 0 ? first : 1u);
                                }
                                first = 1u;
                                umma_commit(&empty[stage]);
                                if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
                            }
                        } else if (p.b_resident) {
                            uint32_t b_lo = (uint32_t)umma_desc(b_ring + (uint32_t)(cb * 9) * p.b_bytes);
                            const uint32_t b_step = p.b_bytes >> 4;
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint32_t at = a_lo + (uint32_t)(((tap / 3) * HALO_W + (tap % 3)) * 8);
                                if (ksteps == 4) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        umma_f16_lo(d_tmem, at + 2 * k, b_lo + 2 * k, desc_hi, idesc, first);
                                        first = 1u;
                                    }
                                } else {
                                    for (int k = 0; k < ksteps; ++k) {
                                        umma_f16_lo(d_tmem, at + 2 * k, b_lo + 2 * k, desc_hi, idesc, first);
                                        first = 1u;
                                    }
                                }
                                b_lo += b_step;
                            }
                        } else {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint32_t at = a_lo + (uint32_t)(((tap / 3) * HALO_W + (tap % 3)) * 8);
                                mbar_wait(&full[stage], phase);
                                tc_fence_after();
                                const uint32_t b_lo = (uint32_t)umma_desc(b_ring + stage * p.b_bytes);
                                if (ksteps == 4) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        umma_f16_lo(d_tmem, at + 2 * k, b_lo + 2 * k, desc_hi, idesc, first);
                                        first = 1u;
                                    }
                                } else {
                                    for (int k = 0; k < ksteps; ++k) {
                                        umma_f16_lo(d_tmem, at + 2 * k, b_lo + 2 * k, desc_hi, idesc, first);
                                        first = 1u;
                                    }
                                }
                                umma_commit(&empty[stage]);
                                if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
                            }
                        }
                        umma_commit(&a_empty[a_stage]);      // halo slot reusable once all nine taps retired
                        if (++a_stage == HALO_A_STAGES) { a_stage = 0; a_phase ^= 1; }
                    }
                    umma_commit(&acc_full[acc]);
                    if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
                    continue;
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int cb = kb % p.k_blocks_per_tap;
                    const int rem = c.Cin - cb * BLOCK_K;
                    const int ksteps = rem >= BLOCK_K ? BLOCK_K / UMMA_K : (rem + UMMA_K - 1) / UMMA_K;
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t a_lo = (uint32_t)umma_desc(sa), b_lo = a_lo + (A_STAGE_BYTES >> 4);
                    const uint32_t desc_hi = (uint32_t)(umma_desc(0) >> 32);
                    // advance both descriptors by k * 32 B inside the 128 B swizzle row
                    if (ksteps == 4) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                    } else {
                        for (int k = 0; k < ksteps; ++k) umma_f16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);              // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[acc]);                 // accumulator complete -> epilogue
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 2 warpgroups x 128 threads; TMEM lane = tile row =====================
        const int et = threadIdx.x - 128;          // 0..255
        const int row = et & 127;                  // tile row == TMEM lane
        const int wg = et >> 7;                    // which half of the columns this warpgroup drains
        const int quarter = warp & 3;              // TMEM lanes [32*quarter, +32) are visible to this warp
        float *s_part = s_projb + 4;               // [128][4] projection partials of warpgroup 1
        int acc = 0;
        uint32_t acc_phase = 0;
        int cached_nb = -1;
        const bf16 *res1 = static_cast<const bf16 *>(c.res1);
        const bf16 *res2 = static_cast<const bf16 *>(c.res2);
        bf16 *y = static_cast<bf16 *>(c.y);
        bf16 *y_relu = static_cast<bf16 *>(c.y_relu);
        const int m_valid = p.BW * p.BH * p.BN;
        // column split in whole 32-column chunks (the TMA-store granule): warpgroup 0 takes ceil(chunks/2) of them,
        // warpgroup 1 the rest plus a possible 16-column tail (N = 144, 16)
        const int c_split = (((p.block_n >> 5) + 1) >> 1) << 5;
        const bool m2 = HALO && p.m2;              // row pair: warpgroup wg drains ALL columns of M tile wg (TMEM columns wg * 128 ...)
        // an ODD number of chunks (N = 96: the S0 qkv / S1 fc layers) would give warpgroup 0 two chunks and warpgroup 1 one on
        // every tile; the split alternates with the tile instead (floor / ceil), so both drain 3 chunks per 2 tiles -- these
        // narrow-K layers are epilogue-bound
        const bool alt_split = p.alt_split && !m2 && MODE != 2 && ((p.block_n >> 5) & 1) && (p.block_n & 31) == 0 && p.block_n > 32;
        int it = 0;
        int eit_count = 0;      // tiles this CTA's epilogue has started (trace builds)
        const int bw = row % p.BW, bh = (row / p.BW) % p.BH, bn = row / (p.BW * p.BH);     // row -> pixel of the tile box
        const float up_sh = c.up_src ? (float)(c.up_h - 1) / (float)(2 * c.up_h - 1) : 0.0f;
        const float up_sw = c.up_src ? (float)(c.up_w - 1) / (float)(2 * c.up_w - 1) : 0.0f;
        bool first_store = true;                   // the staging tile starts out free
        TileIter ti;
        ti.init(p, (int)blockIdx.x, (int)gridDim.x);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ti.next(p)) {
            const int nb = ti.nb, tw = ti.tw, th = ti.th, tn = ti.tn;
            const int cout0 = nb * p.block_n;
            const int split = (alt_split && (it++ & 1)) ? c_split - 32 : c_split;
            const int col_begin = m2 || wg == 0 ? 0 : split, col_end = m2 ? p.block_n : (wg == 0 ? split : p.block_n);
            if (tile + (int)gridDim.x >= p.total_tiles) soccdpt::pdl_trigger();   // this CTA's last tile: let the next kernel in
            if (nb != cached_nb) {
                // per-channel constants of this N block -> smem (visible to the 256 epilogue threads only)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int i = et; i < p.block_n; i += EPI_THREADS) s_bias[i] = c.bias ? c.bias[cout0 + i] : 0.0f;
                if (MODE == 2) {
                    for (int i = et; i < c.proj_n * c.Cout; i += EPI_THREADS) s_projw[i] = c.proj_w[i];
                    if (et < c.proj_n) s_projb[et] = c.proj_b[et];
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                cached_nb = nb;
            }
            const int h_tile = m2 ? th * 2 + wg : th * p.BH;
            const int wx = tw * p.BW + bw, hy = h_tile + bh, ni = tn * p.BN + bn;
            const bool valid = (row < m_valid) && (wx < p.Wo) && (hy < p.Ho) && (ni < c.N);
            const long long pix = ((long long)ni * p.Ho + hy) * p.Wo + wx;
            const long long obase = pix * c.Cout + cout0;
            // up-sampled residual (soccdpt_conv_t.up_src): the four low-resolution corners of this pixel and their weights
            const bf16 *up00 = nullptr, *up01 = nullptr, *up10 = nullptr, *up11 = nullptr;
            float uw00 = 0.f, uw01 = 0.f, uw10 = 0.f, uw11 = 0.f;
            if (MODE == 1 && c.up_src && valid) {
                const float fy = up_sh * (float)hy, fx = up_sw * (float)wx;
                const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + (y0 < c.up_h - 1 ? 1 : 0), x1 = x0 + (x0 < c.up_w - 1 ? 1 : 0);
                const float ly = fy - (float)y0, lx = fx - (float)x0;
                const bf16 *ub = static_cast<const bf16 *>(c.up_src) + (long long)ni * c.up_h * c.up_w * c.Cout + cout0;
                up00 = ub + ((long long)y0 * c.up_w + x0) * c.Cout;
                up01 = ub + ((long long)y0 * c.up_w + x1) * c.Cout;
                up10 = ub + ((long long)y1 * c.up_w + x0) * c.Cout;
                up11 = ub + ((long long)y1 * c.up_w + x1) * c.Cout;
                uw00 = (1.0f - ly) * (1.0f - lx); uw01 = (1.0f - ly) * lx; uw10 = ly * (1.0f - lx); uw11 = ly * lx;
            }

            if (et == 0) CTRACE(eit_count, 8);
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            if (et == 0) CTRACE(eit_count, 9);
            int cev = 10;
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + (m2 ? wg * 128 : 0));
            float proj[4] = {0.f, 0.f, 0.f, 0.f};
            int col = col_begin;
            // ---- 32-column steps
            for (; col + 32 <= col_end; col += 32) {
                uint32_t raw[32];
                if (et == 0 && cev < 26) CTRACE(eit_count, cev++);      // chunk start
                tmem_ld32(t_row + (uint32_t)col, raw);
                tmem_ld_wait();
                if (et == 0 && cev < 26) CTRACE(eit_count, cev++);      // TMEM load done
                float v[32];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4 *>(s_bias + col + j4 * 4);
                    v[j4 * 4 + 0] = __uint_as_float(raw[j4 * 4 + 0]) + bb.x;
                    v[j4 * 4 + 1] = __uint_as_float(raw[j4 * 4 + 1]) + bb.y;
                    v[j4 * 4 + 2] = __uint_as_float(raw[j4 * 4 + 2]) + bb.z;
                    v[j4 * 4 + 3] = __uint_as_float(raw[j4 * 4 + 3]) + bb.w;
                }
                if (MODE == 0 && ACT == SOCCDPT_ACT_NONE && c.qk_heads > 0) {
                    // cosine attention (timm WindowAttention): the 32 columns of this step are one head of q or k (N blocks
                    // and column steps are head aligned): L2-normalise from the fp32 accumulator, q also takes the logit scale
                    const int gc = cout0 + col;
                    if (gc < 64 * c.qk_heads) {
                        // x / max(|x|, 1e-12) = x * rsqrt(max(|x|^2, 1e-24)); four partial sums instead of a 32-deep FMA chain
                        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int j = 0; j < 32; ++j) s4[j & 3] = fmaf(v[j], v[j], s4[j & 3]);
                        float inv = rsqrtf(fmaxf((s4[0] + s4[1]) + (s4[2] + s4[3]), 1e-24f));
                        if (gc < 32 * c.qk_heads) inv *= c.qk_scale[gc >> 5];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= inv;
                    }
                }
                if (ACT == SOCCDPT_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                } else if (ACT == SOCCDPT_ACT_GELU) {
                    // packed fp32x2 one-MUFU erf GELU (tc_ptx.cuh): 8 issue slots per element instead of ~20 -- the GELU epilogue
                    // is issue bound (profiles/r1_progress.md step 16)
#pragma unroll
                    for (int j = 0; j < 32; j += 2) tc::un2(tc::gelu_erf_f32x2(tc::mk2(v[j], v[j + 1])), v[j], v[j + 1]);
                }
                if (MODE == 2) {
                    for (int q = 0; q < c.proj_n; ++q) {
                        const float *pw = s_projw + q * c.Cout + col;
                        float sacc = proj[q];
#pragma unroll
                        for (int j = 0; j < 32; ++j) sacc = fmaf(pw[j], v[j], sacc);
                        proj[q] = sacc;
                    }
                } else {
                    if (MODE == 1 && valid) {
                        const long long o = obase + col;
                        if (res1) add_bf16<4>(v, reinterpret_cast<const uint4 *>(res1 + o));
                        if (res2) add_bf16<4>(v, reinterpret_cast<const uint4 *>(res2 + o));
                        if (up00) {
                            fma_bf16<4>(v, reinterpret_cast<const uint4 *>(up00 + col), uw00);
                            fma_bf16<4>(v, reinterpret_cast<const uint4 *>(up01 + col), uw01);
                            fma_bf16<4>(v, reinterpret_cast<const uint4 *>(up10 + col), uw10);
                            fma_bf16<4>(v, reinterpret_cast<const uint4 *>(up11 + col), uw11);
                        }
                    }
                    // staged stores: the bulk tensor store, its commit and the wait for it to have READ the staging tile cost the
                    // issuing thread ~220 cycles each with the other 127 threads parked at the next barrier (tools/trace_conv.py);
                    // warp 2 / 3 (store warp of warpgroup 0 / 1) does all three, this warpgroup only waits for "tile free"
                    // (normally long passed) and announces "tile staged"
                    uint8_t *stage = s_stage + wg * STAGING_BYTES;
                    if (MODE == 0 || y) {
                        if (et == 0 && cev < 26) CTRACE(eit_count, cev++);      // math done
                        if (!first_store) bar_sync(BAR_FREE + wg, 160);
                        first_store = false;
                        if (et == 0 && cev < 26) CTRACE(eit_count, cev++);      // staging tile free
                        stage_row32<false>(stage, row, v);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        bar_arrive(BAR_STAGED + wg, 160);
                        if (et == 0 && cev < 26) CTRACE(eit_count, cev++);      // staged + proxy fence + arrive
                    }
                    if (MODE == 1 && y_relu) {
                        if (!first_store) bar_sync(BAR_FREE + wg, 160);
                        first_store = false;
                        stage_row32<true>(stage, row, v);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        bar_arrive(BAR_STAGED + wg, 160);
                    }
                }
            }
            // ---- 16-column tail
            for (; col < col_end; col += 16) {
                uint32_t raw[16];
                tmem_ld16(t_row + (uint32_t)col, raw);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float t = __uint_as_float(raw[j]) + s_bias[col + j];
                    if (ACT == SOCCDPT_ACT_RELU) t = fmaxf(t, 0.0f);
                    else if (ACT == SOCCDPT_ACT_GELU) t = gelu_erf(t);
                    v[j] = t;
                }
                if (MODE == 2) {
                    for (int q = 0; q < c.proj_n; ++q) {
                        const float *pw = s_projw + q * c.Cout + col;
                        float sacc = proj[q];
#pragma unroll
                        for (int j = 0; j < 16; ++j) sacc = fmaf(pw[j], v[j], sacc);
                        proj[q] = sacc;
                    }
                } else if (valid) {
                    const long long o = obase + col;
                    if (MODE == 1) {
                        if (res1) add_bf16<2>(v, reinterpret_cast<const uint4 *>(res1 + o));
                        if (res2) add_bf16<2>(v, reinterpret_cast<const uint4 *>(res2 + o));
                        if (y) store_bf16<2, false>(v, reinterpret_cast<uint4 *>(y + o));
                        if (y_relu) store_bf16<2, true>(v, reinterpret_cast<uint4 *>(y_relu + o));
                    } else {
                        store_bf16<2, false>(v, reinterpret_cast<uint4 *>(y + o));
                    }
                }
            }
            // accumulator fully read by this thread: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (et == 0) CTRACE(eit_count, 30);
            ++eit_count;
            if (MODE == 2) {
                // combine the two column halves: warpgroup 1 -> smem -> warpgroup 0 adds, finishes, stores
                if (wg == 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) s_part[row * 4 + q] = proj[q];
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (wg == 0 && valid) {
                    for (int q = 0; q < c.proj_n; ++q) {
                        float sacc = proj[q] + s_part[row * 4 + q] + s_projb[q];
                        if (c.proj_relu) sacc = fmaxf(sacc, 0.0f);
                        c.proj_out[pix * c.proj_n + q] = sacc;
                    }
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    } else if (MODE != 2) {
        // ===================== store warps: warp 2 / 3 issue the bulk tensor stores of epilogue warpgroup 0 / 1 =====================
        // (same tile / chunk sequence as the warpgroup; MODE 2 writes no staged output)
        const int swg = warp - 2;
        const bool m2 = HALO && p.m2;
        const int c_split = (((p.block_n >> 5) + 1) >> 1) << 5;
        const bool alt_split = p.alt_split && !m2 && ((p.block_n >> 5) & 1) && (p.block_n & 31) == 0 && p.block_n > 32;
        const bool has_y = MODE == 0 || c.y != nullptr, has_yr = MODE == 1 && c.y_relu != nullptr;
        uint8_t *stage = s_stage + swg * STAGING_BYTES;
        int it = 0;
        TileIter ti;
        ti.init(p, (int)blockIdx.x, (int)gridDim.x);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ti.next(p)) {
            const int split = (alt_split && (it++ & 1)) ? c_split - 32 : c_split;
            const int col_begin = m2 || swg == 0 ? 0 : split, col_end = m2 ? p.block_n : (swg == 0 ? split : p.block_n);
            const int cout0 = ti.nb * p.block_n, h_tile = m2 ? ti.th * 2 + swg : ti.th * p.BH;
            const bool last_tile = tile + (int)gridDim.x >= p.total_tiles;
            for (int col = col_begin; col + 32 <= col_end; col += 32) {
                const bool last_chunk = last_tile && col + 64 > col_end;
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    if (which == 0 ? !has_y : !has_yr) continue;
                    bar_sync(BAR_STAGED + swg, 160);                   // the warpgroup has staged (and proxy-fenced) the chunk
                    if (lane == 0) {
                        tma_store_4d(which == 0 ? &map_y : &map_yr, stage, cout0 + col, ti.tw * p.BW, h_tile, ti.tn * p.BN);
                        tma_store_commit();
                        tma_store_wait_read();                         // the tile may be overwritten
                    }
                    __syncwarp();
                    const bool last = last_chunk && (which == 1 || !has_yr);
                    if (!last) bar_arrive(BAR_FREE + swg, 160);      // nobody waits for the tile after the last store
                }
            }
        }
        if (lane == 0) tma_store_wait_all();       // staged tiles must outlive the bulk stores reading them
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int largest_divisor_leq(int n, int cap) {
    for (int d = cap < n ? cap : n; d >= 1; --d)
        if (n % d == 0) return d;
    return 1;
}

int pick_block_n(int cout, int step = 16) {   // step 32: N blocks made of whole 32-channel heads (cosine-attention epilogue)
    for (int bn = 256; bn >= step; bn -= step)
        if (cout % bn == 0) return bn;
    return 0;
}

}  // namespace

#ifdef SOCCDPT_CONV_TRACE
extern "C" int soccdpt_conv_trace_read(unsigned long long *dst) {     // CT_TILES x CT_EVENTS stamps of the last launches (debug builds)
    SOCCDPT_CUDA(cudaDeviceSynchronize());
    SOCCDPT_CUDA(cudaMemcpyFromSymbol(dst, g_conv_trace, sizeof(unsigned long long) * CT_TILES * CT_EVENTS));
    return 0;
}
#endif

extern "C" int soccdpt_conv_fwd(const soccdpt_conv_t *c, soccdpt_stream_t stream) {
    int rc = soccdpt::validate_conv(c);
    if (rc) return rc;
    EncodeTiledFn encode = encode_fn();
    SOCCDPT_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    SOCCDPT_REQUIRE((reinterpret_cast<uintptr_t>(c->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(c->wgt) & 15) == 0,
                    "conv: x / wgt must be 16-byte aligned");
    SOCCDPT_REQUIRE(c->Cout % 16 == 0, "conv: Cout must be a multiple of 16 (got %d)", c->Cout);

    Params p{};
    p.c = *c;
    p.block_n = pick_block_n(c->Cout, c->qk_heads > 0 ? 32 : 16);
    SOCCDPT_REQUIRE(p.block_n >= 16, "conv: no valid N tile for Cout=%d", c->Cout);
    if (c->proj_n > 0) SOCCDPT_REQUIRE(p.block_n == c->Cout, "conv: fused projection needs the whole Cout in one tile");
    p.stride = c->stride > 1 ? c->stride : 1;
    p.pad = c->KH / 2 - c->pad_trim;
    p.Ho = (c->H + p.stride - 1) / p.stride;
    p.Wo = (c->W + p.stride - 1) / p.stride;
    // M tile = box of OUTPUT pixels {BW, BH, BN}
    if (p.Wo >= BLOCK_M) {
        p.BW = (p.Wo % BLOCK_M == 0) ? BLOCK_M : largest_divisor_leq(p.Wo, BLOCK_M);
        if (p.BW < 64) p.BW = BLOCK_M;   // poor divisor: use full tiles, mask the tail
        p.BH = 1;
        p.BN = 1;
    } else {
        p.BW = p.Wo;
        const int rows = BLOCK_M / p.Wo;
        if (p.Ho >= rows) {
            p.BH = largest_divisor_leq(p.Ho, rows);
            p.BN = 1;
        } else {
            p.BH = p.Ho;
            p.BN = BLOCK_M / (p.Wo * p.Ho);
            if (p.BN > c->N) p.BN = c->N;
            if (p.BN < 1) p.BN = 1;
        }
    }
    // halo reuse: 3x3, full 128-pixel row segments (the 128^2 / 256^2 levels: both heads of the tiny model)
    static const bool halo_enabled = !(getenv("SOCCDPT_CONV_HALO") && getenv("SOCCDPT_CONV_HALO")[0] == '0');
    const bool halo = halo_enabled && c->KH == 3 && p.stride == 1 && p.pad == 1 && c->W % BLOCK_M == 0 && p.BW == BLOCK_M &&
                      p.BH == 1 && p.BN == 1;
    p.b_resident = (halo && p.block_n == c->Cout &&
                    (long long)((c->Cin + BLOCK_K - 1) / BLOCK_K) * 9 * p.block_n * BLOCK_K * 2 <= HALO_B_BYTES) ? 1 : 0;
    const int mode = c->proj_n > 0 ? 2 : ((c->res1 || c->res2 || c->up_src || c->y_relu || !c->y) ? 1 : 0);
    if (c->up_src) SOCCDPT_REQUIRE((p.block_n & 31) == 0, "conv: the up-sampled residual needs N blocks of whole 32-column chunks (Cout=%d)", c->Cout);
    static const bool alt_enabled = !(getenv("SOCCDPT_CONV_ALT") && getenv("SOCCDPT_CONV_ALT")[0] == '0');
    p.alt_split = alt_enabled ? 1 : 0;
    static const bool m2_enabled = !(getenv("SOCCDPT_CONV_M2") && getenv("SOCCDPT_CONV_M2")[0] == '0');
    p.m2 = (m2_enabled && halo && !p.b_resident && mode != 2 && p.block_n <= 128 && p.Ho % 2 == 0 && c->qk_heads == 0) ? 1 : 0;
    p.halo_slot = p.m2 ? HALO2_BYTES : HALO_STAGE_BYTES;
    p.halo_bytes = p.m2 ? HALO2_BYTES : HALO_BYTES;
    p.tiles_w = (p.Wo + p.BW - 1) / p.BW;
    p.tiles_h = p.m2 ? p.Ho / 2 : (p.Ho + p.BH - 1) / p.BH;
    p.tiles_n = (c->N + p.BN - 1) / p.BN;
    p.n_blocks = c->Cout / p.block_n;
    // sub-wave layers (the 8x8 level of the decoder: 32 M tiles on 148 SMs, each a serial chain of K / 16 MMAs): narrower N blocks
    // put more SMs on the same work -- the launch is latency-bound, not throughput-bound (SOCCDPT_CONV_NSPLIT=0 switches it off)
    static const bool nsplit_enabled = !(getenv("SOCCDPT_CONV_NSPLIT") && getenv("SOCCDPT_CONV_NSPLIT")[0] == '0');
    while (nsplit_enabled && !halo && c->proj_n == 0 && c->qk_heads == 0 && p.block_n >= 128 && p.block_n % 64 == 0 &&
           2ll * p.tiles_w * p.tiles_h * p.tiles_n * p.n_blocks <= soccdpt::sm_count()) {
        p.block_n /= 2;
        p.n_blocks *= 2;
    }
    const long long total = (long long)p.tiles_w * p.tiles_h * p.tiles_n * p.n_blocks;
    SOCCDPT_REQUIRE(total < (1ll << 31), "conv: too many tiles");
    p.total_tiles = (int)total;
    p.k_blocks_per_tap = (c->Cin + BLOCK_K - 1) / BLOCK_K;
    p.a_bytes = (uint32_t)(p.BW * p.BH * p.BN) * BLOCK_K * 2;
    p.b_bytes = (uint32_t)p.block_n * BLOCK_K * 2;
    p.b_stages = (p.m2 ? RING_BYTES - HALO_A_STAGES * HALO2_BYTES : HALO_B_BYTES) / (int)p.b_bytes;
    if (p.b_stages > MAX_B_STAGES) p.b_stages = MAX_B_STAGES;
    // resident N block (see Params::nres): 1x1, stride 1, the block's weights fit behind an A ring of >= 3 slots, and enough
    // M tiles per CTA for the one-off weight fetch to pay
    // (per-layer A/B in profiles/r1_progress.md step 18: qkv -7..-10 %, S0 fc2 -13 %, out_conv / tap GEMM -4..-5 %, nothing slower)
    static const bool nres_enabled = !(getenv("SOCCDPT_CONV_NRES") && getenv("SOCCDPT_CONV_NRES")[0] == '0');
    const long long bres_bytes = (long long)p.k_blocks_per_tap * p.b_bytes;
    p.nres = 0;
    p.a_stages = STAGES;
    if (nres_enabled && !halo && c->KH == 1 && c->KW == 1 && p.stride == 1 && p.n_blocks <= soccdpt::sm_count() &&
        bres_bytes + 3 * A_STAGE_BYTES <= RING_BYTES_PLAIN && p.total_tiles >= 2 * soccdpt::sm_count()) {
        p.nres = 1;
        long long slots = (RING_BYTES_PLAIN - bres_bytes) / A_STAGE_BYTES;
        p.a_stages = (int)(slots < MAX_B_STAGES ? slots : MAX_B_STAGES);
    }

    CUtensorMap map_a, map_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)c->Cin, (cuuint64_t)c->W, (cuuint64_t)c->H, (cuuint64_t)c->N};
        cuuint64_t strides[3] = {(cuuint64_t)c->Cin * 2, (cuuint64_t)c->W * c->Cin * 2, (cuuint64_t)c->H * c->W * c->Cin * 2};
        // stride 2: the box spans stride*BW input pixels and the traversal stride keeps every second one
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)(p.BW * p.stride), (cuuint32_t)(p.BH * p.stride), (cuuint32_t)p.BN};
        if (halo) { box[1] = HALO_W; box[2] = p.m2 ? 4 : 3; }
        cuuint32_t estr[4] = {1, (cuuint32_t)p.stride, (cuuint32_t)p.stride, 1};
        SOCCDPT_REQUIRE(box[1] <= 256 && box[2] <= 256, "conv: tile too wide for a strided TMA box");
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(c->x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed with %d (N=%d H=%d W=%d Cin=%d box=%d,%d,%d)", (int)r,
                        c->N, c->H, c->W, c->Cin, p.BW, p.BH, p.BN);
    }
    {
        const int taps = c->KH * c->KW;
        cuuint64_t dims[3] = {(cuuint64_t)c->Cin, (cuuint64_t)taps, (cuuint64_t)c->Cout};
        cuuint64_t strides[2] = {(cuuint64_t)c->Cin * 2, (cuuint64_t)taps * c->Cin * 2};
        cuuint32_t box[3] = {BLOCK_K, 1, (cuuint32_t)p.block_n};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(c->wgt), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed with %d (Cout=%d taps=%d Cin=%d)", (int)r, c->Cout, taps, c->Cin);
    }

    // output maps for the TMA-store epilogue: same pixel box as the A tile, 32 channels, 64-byte swizzle
    CUtensorMap map_y = map_a, map_yr = map_a;
    for (int which = 0; which < 2; ++which) {
        void *dst = which == 0 ? c->y : c->y_relu;
        if (!dst) continue;
        SOCCDPT_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "conv: outputs must be 16-byte aligned");
        cuuint64_t dims[4] = {(cuuint64_t)c->Cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)c->N};
        cuuint64_t strides[3] = {(cuuint64_t)c->Cout * 2, (cuuint64_t)p.Wo * c->Cout * 2, (cuuint64_t)p.Ho * p.Wo * c->Cout * 2};
        cuuint32_t box[4] = {32, (cuuint32_t)p.BW, (cuuint32_t)p.BH, (cuuint32_t)p.BN};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(which == 0 ? &map_y : &map_yr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dst, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SOCCDPT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(out) failed with %d", (int)r);
    }
    if (mode == 2) SOCCDPT_REQUIRE(c->y == nullptr && c->y_relu == nullptr && !c->res1 && !c->res2,
                                   "conv: the fused projection epilogue produces proj_out only");
    int grid = p.total_tiles < soccdpt::sm_count() ? p.total_tiles : soccdpt::sm_count();
    if (p.nres) grid = grid / p.n_blocks * p.n_blocks;      // tile += grid keeps tile % n_blocks, i.e. the CTA's N block
    cudaStream_t st = soccdpt::as_stream(stream);
#define SOCC_LAUNCH(A, M, HL)                                                                                           \
    do {                                                                                                                \
        static soccdpt::SmemAttr configured;                                                                            \
        if (configured.need(SMEM_BYTES)) {                                                                              \
            SOCCDPT_CUDA(cudaFuncSetAttribute(conv_tcgen05_kernel<A, M, HL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              SMEM_BYTES));                                                             \
        }                                                                                                               \
        SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_CONV, conv_tcgen05_kernel<A, M, HL>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, map_a, map_b, \
                                         map_y, map_yr, p));                                                            \
    } while (0)
#define SOCC_MODES(A)                                                    \
    do {                                                                 \
        if (halo) {                                                      \
            if (mode == 0) SOCC_LAUNCH(A, 0, true);                      \
            else if (mode == 1) SOCC_LAUNCH(A, 1, true);                 \
            else SOCC_LAUNCH(A, 2, true);                                \
        } else {                                                         \
            if (mode == 0) SOCC_LAUNCH(A, 0, false);                     \
            else if (mode == 1) SOCC_LAUNCH(A, 1, false);                \
            else SOCC_LAUNCH(A, 2, false);                               \
        }                                                                \
    } while (0)
    if (c->act == SOCCDPT_ACT_NONE) SOCC_MODES(0);
    else if (c->act == SOCCDPT_ACT_RELU) SOCC_MODES(1);
    else SOCC_MODES(2);
#undef SOCC_MODES
#undef SOCC_LAUNCH
    return soccdpt::check_launch("conv_tcgen05_kernel");
}
