// Tail of the DPT depth head (reference SOccDPT/model/dpt.py:209-219):
//     Interpolate(x2, bilinear, align_corners=True) -> Conv2d(128, 32, 3, pad 1) -> ReLU -> Conv2d(32, 1, 1) -> ReLU
//
// conv3x3 after a bilinear upsample is linear in its input, so it is evaluated at LOW resolution first:
//     T[n, y, x, tap*32 + c] = sum_k W2[c, k, tap] * d0[n, y, x, k]          one GEMM, N = 9*32 = 288 (tcgen05 kernel)
//     out[n, Y, X]           = relu( pb + sum_c pw[c] * relu( b2[c] + sum_tap bilerp(T[..., tap*32 + c])(Y+dy, X+dx) ) )
// with taps that fall outside the upsampled image contributing zero (the conv's zero padding).  This kernel is
// the second line, evaluated separably (vertical pass into a shared-memory row buffer, then horizontal pass),
// fp32 accumulation.  It replaces a 1.07 GB (B=64) upsampled intermediate and a narrow N=32 implicit GEMM that
// is bound by the tensor core's A-operand read; FLOPs drop 4x.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using bf16 = __nv_bfloat16;
constexpr int CO = 32;          // head_features_2
constexpr int TAPS = 9;
constexpr int TC = TAPS * CO;   // 288 channels of T

constexpr int VC = 3 * CO;      // 96 values per source column: (dx, c)
constexpr int VP = VC + 4;      // smem pitch in floats (400 B): conflict-free float4 reads across neighbouring columns

#ifndef SOCCDPT_DT_KP
#define SOCCDPT_DT_KP 32
#endif
#ifndef SOCCDPT_DT_NRING
#define SOCCDPT_DT_NRING 4
#endif
#ifndef SOCCDPT_DT_CTAS
#define SOCCDPT_DT_CTAS 3
#endif
constexpr int SW = 16;          // low-resolution columns per CTA (32 output columns)
constexpr int NC = SW + 2;      // + one halo column on each side
constexpr int KP = SOCCDPT_DT_KP;          // low-resolution rows (= output row PAIRS) per CTA
constexpr int NRING = SOCCDPT_DT_NRING;    // T rows resident in shared memory: three in use, the rest in flight.  Four rows (70 KB with V) let
                                           // THREE CTAs share an SM: 221 us against 247 us for six rows / two CTAs (one row in flight per CTA is
                                           // enough when two other CTAs cover its wait)
constexpr int ROWB = NC * TC * 2;                    // bytes of one staged T row segment (10368)
constexpr int VROW = NC * VP;                        // floats of one V row
constexpr int SMEM_BYTES = NRING * ROWB + 2 * 2 * VROW * 4 + NRING * 8;

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
// bf16 pair -> two fp32 lanes of one packed register
__device__ __forceinline__ tc::f32x2 bf2(uint32_t w) { return tc::mk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

// acc[0..3] (8 channels, packed pairs) += wa * a + wb * b
__device__ __forceinline__ void lerp8(tc::f32x2 (&acc)[4], const uint4 &a, const uint4 &b, tc::f32x2 wa, tc::f32x2 wb) {
    acc[0] = tc::ffma2(wa, bf2(a.x), tc::ffma2(wb, bf2(b.x), acc[0]));
    acc[1] = tc::ffma2(wa, bf2(a.y), tc::ffma2(wb, bf2(b.y), acc[1]));
    acc[2] = tc::ffma2(wa, bf2(a.z), tc::ffma2(wb, bf2(b.z), acc[2]));
    acc[3] = tc::ffma2(wa, bf2(a.w), tc::ffma2(wb, bf2(b.w), acc[3]));
}
__device__ __forceinline__ void lerp4(float (&acc)[4], const float4 &a, const float4 &b, float wa, float wb) {
    acc[0] = fmaf(wa, a.x, fmaf(wb, b.x, acc[0]));
    acc[1] = fmaf(wa, a.y, fmaf(wb, b.y, acc[1]));
    acc[2] = fmaf(wa, a.z, fmaf(wb, b.z, acc[2]));
    acc[3] = fmaf(wa, a.w, fmaf(wb, b.w, acc[3]));
}

// Both the bilinear interpolation and the zero padding of the 3x3 conv are separable, so an output row is produced in two
// passes instead of 9 taps x 4 corners per pixel:
//   pass A  V[x][dx*32+c] = sum_{dy: 0 <= Y+dy-1 < H} lerp_y(T[y0|y1][x][(dy*3+dx)*32 + c])      (columns x 96 fp32 in smem)
//   pass B  out[X]        = relu(pb + sum_c pw[c] relu(b2[c] + sum_{dx: 0 <= X+dx-1 < W} lerp_x(V[x0|x1][dx*32+c])))
// With align_corners=True and an exact x2 factor the source index of tap position t is floor(t (h-1)/(2h-1)): rows (j-1, j)
// for t = 2j and (j, j+1) for t = 2j+1.  The output row pair (2k, 2k+1) therefore touches the tap rows 2k-1 .. 2k+2 and only
// the source rows k-1, k, k+1 -- and the same holds for columns -- so the kernel works on 2x2 output blocks:
//   * a CTA owns SW source columns (+ halo) and MARCHES down KP source rows.  Every T row segment is fetched ONCE, by one
//     cp.async.bulk into a ring of NRING rows (three rows in use, one in flight), completion on one mbarrier per slot.
//     The first version (one CTA per output row, every row re-reading its six source row slices from L2: 2.4 GB of L2 reads
//     per 64-frame step for 0.6 GB of T) was L2-bound at 0.33 ms;
//   * pass A: a thread owns 8 (dx, c) values of one column: 7 shared-memory loads (instead of 12) feed both rows of the pair,
//     packed fma.f32x2;   pass B: a thread owns 4 channels of one 2-pixel block of one row: 7 loads instead of 12, then a
//     three-step shuffle reduction over the 8 channel groups.  V is double-buffered: one __syncthreads per row pair.
__global__ void __launch_bounds__(256, SOCCDPT_DT_CTAS)
depth_tail_kernel(const bf16 *__restrict__ T, const float *__restrict__ b2, const float *__restrict__ pw,
                  const float *__restrict__ pb, float *__restrict__ out, int N, int h, int w, int segs, int strips) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *ring = smem;
    float *V = reinterpret_cast<float *>(smem + NRING * ROWB);               // [2 buffers][2 rows][NC][VP]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + NRING * ROWB + 2 * 2 * VROW * 4);
    const int tid = threadIdx.x;
    const int seg = blockIdx.x % segs, strip = (blockIdx.x / segs) % strips, n = blockIdx.x / (segs * strips);
    const int H = 2 * h, W = 2 * w;
    const float sh = (float)(h - 1) / (float)(H - 1), sw = (float)(w - 1) / (float)(W - 1);   // align_corners=True
    const int xs = seg * SW, xlo = max(xs - 1, 0), xhi = min(xs + SW, w - 1);
    const int ncols = xhi - xlo + 1, mcols = min(SW, w - xs);
    const int k0 = strip * KP, k1 = min(k0 + KP, h), r0 = max(k0 - 1, 0), r1 = min(k1, h - 1);
    const uint32_t rowbytes = (uint32_t)ncols * (TC * 2);
    const bf16 *src0 = T + ((size_t)n * h * w + xlo) * TC;                  // row r of the segment: src0 + r * w * TC

    if (tid == 0) {
        for (int i = 0; i < NRING; ++i) tc::mbar_init(bars + i, 1);
        tc::fence_barrier_init();
    }
    // ---- per-thread constants of pass B: (row of the pair, 2-pixel block, 4-channel group)
    const int cg = tid & 7, m = (tid >> 3) & (SW - 1), jy = tid >> 7;
    const int M = xs + m;
    const bool activeB = m < mcols;
    float lx0[4], lx1[4];                          // weights of the tap columns 2M-1 .. 2M+2 on their (left, right) source column
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int Xt = 2 * M - 1 + t;
        const float base = (float)(t < 2 ? M - 1 : M);
        const float l = sw * (float)Xt - base;
        const bool ok = Xt >= 0 && Xt < W;         // zero padding of the 3x3 conv
        lx0[t] = ok ? 1.0f - l : 0.0f;
        lx1[t] = ok ? l : 0.0f;
    }
    const int colm = max(M - 1, 0) - xlo, col0 = min(M, w - 1) - xlo, colp = min(M + 1, w - 1) - xlo;
    const int vb_m = colm * VP + cg * 4, vb_0 = col0 * VP + cg * 4, vb_p = colp * VP + cg * 4;
    const float4 bias4 = *reinterpret_cast<const float4 *>(b2 + cg * 4), pw4 = *reinterpret_cast<const float4 *>(pw + cg * 4);
    const float pbias = pb[0];
    // ---- per-thread constants of pass A: (column, 8 of its 96 (dx, c) values)
    const int ax = tid / 12, aq = tid - ax * 12;
    const bool activeA = tid < ncols * 12;
    const int a_src = ax * (TC * 2) + aq * 16, a_dst = ax * VP + aq * 8;
    __syncthreads();
    soccdpt::pdl_wait();        // T is the previous kernel's output

    int issued = r0 - 1;
    auto issue_to = [&](int last) {
        while (issued < last) {
            ++issued;
            const int slot = (issued - r0) % NRING;
            tc::mbar_expect_tx(bars + slot, rowbytes);
            bulk_g2s(ring + slot * ROWB, src0 + (size_t)issued * w * TC, rowbytes, bars + slot);
        }
    };
    if (tid == 0) issue_to(min(r0 + NRING - 1, r1));
    int ready = r0 - 1;

    for (int k = k0; k < k1; ++k) {
        const int rA = max(k - 1, 0), rC = min(k + 1, h - 1);
        while (ready < rC) {
            ++ready;
            tc::mbar_wait(bars + (ready - r0) % NRING, (uint32_t)((ready - r0) / NRING) & 1u);
        }
        float *Vb = V + ((k - k0) & 1) * (2 * VROW);
        if (activeA) {
            // weights of the tap rows 2k-1 .. 2k+2 on their (upper, lower) source row
            tc::f32x2 wy0[4], wy1[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int Yt = 2 * k - 1 + t;
                const float l = sh * (float)Yt - (float)(t < 2 ? k - 1 : k);
                const bool ok = Yt >= 0 && Yt < H;
                const float u = ok ? 1.0f - l : 0.0f, v = ok ? l : 0.0f;
                wy0[t] = tc::mk2(u, u);
                wy1[t] = tc::mk2(v, v);
            }
            const uint8_t *pA = ring + ((rA - r0) % NRING) * ROWB + a_src;
            const uint8_t *pB = ring + ((k - r0) % NRING) * ROWB + a_src;
            const uint8_t *pC = ring + ((rC - r0) % NRING) * ROWB + a_src;
            tc::f32x2 acc0[4], acc1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc0[i] = acc1[i] = 0ull;
            {   // dy = 0: tap rows 2k-1 (row 2k) and 2k (row 2k+1), both on source rows (k-1, k)
                const uint4 a = *reinterpret_cast<const uint4 *>(pA), b = *reinterpret_cast<const uint4 *>(pB);
                lerp8(acc0, a, b, wy0[0], wy1[0]);
                lerp8(acc1, a, b, wy0[1], wy1[1]);
            }
            {   // dy = 1: tap rows 2k on (k-1, k) and 2k+1 on (k, k+1)
                const uint4 a = *reinterpret_cast<const uint4 *>(pA + VC * 2), b = *reinterpret_cast<const uint4 *>(pB + VC * 2),
                            c = *reinterpret_cast<const uint4 *>(pC + VC * 2);
                lerp8(acc0, a, b, wy0[1], wy1[1]);
                lerp8(acc1, b, c, wy0[2], wy1[2]);
            }
            {   // dy = 2: tap rows 2k+1 and 2k+2, both on (k, k+1)
                const uint4 b = *reinterpret_cast<const uint4 *>(pB + 2 * VC * 2), c = *reinterpret_cast<const uint4 *>(pC + 2 * VC * 2);
                lerp8(acc0, b, c, wy0[2], wy1[2]);
                lerp8(acc1, b, c, wy0[3], wy1[3]);
            }
            ulonglong2 *d0 = reinterpret_cast<ulonglong2 *>(Vb + a_dst), *d1 = reinterpret_cast<ulonglong2 *>(Vb + VROW + a_dst);
            d0[0] = make_ulonglong2(acc0[0], acc0[1]);
            d0[1] = make_ulonglong2(acc0[2], acc0[3]);
            d1[0] = make_ulonglong2(acc1[0], acc1[1]);
            d1[1] = make_ulonglong2(acc1[2], acc1[3]);
        }
        __syncthreads();
        // every thread is past pass A of pair k: rows < k are free, so the slots of rows <= k - 1 + NRING may be refilled
        if (tid == 0) issue_to(min(k - 1 + NRING, r1));

        // ---- pass B: horizontal interpolation + sum over the tap columns, bias, ReLU, 32 -> 1, ReLU
        const float *Vr = Vb + jy * VROW;
        float p0[4] = {bias4.x, bias4.y, bias4.z, bias4.w}, p1[4] = {bias4.x, bias4.y, bias4.z, bias4.w};
        {   // dx = 0: tap columns 2M-1 (pixel 2M) and 2M (pixel 2M+1), both on source columns (M-1, M)
            const float4 a = *reinterpret_cast<const float4 *>(Vr + vb_m), b = *reinterpret_cast<const float4 *>(Vr + vb_0);
            lerp4(p0, a, b, lx0[0], lx1[0]);
            lerp4(p1, a, b, lx0[1], lx1[1]);
        }
        {   // dx = 1: tap columns 2M on (M-1, M) and 2M+1 on (M, M+1)
            const float4 a = *reinterpret_cast<const float4 *>(Vr + vb_m + CO), b = *reinterpret_cast<const float4 *>(Vr + vb_0 + CO),
                         c = *reinterpret_cast<const float4 *>(Vr + vb_p + CO);
            lerp4(p0, a, b, lx0[1], lx1[1]);
            lerp4(p1, b, c, lx0[2], lx1[2]);
        }
        {   // dx = 2: tap columns 2M+1 and 2M+2, both on (M, M+1)
            const float4 b = *reinterpret_cast<const float4 *>(Vr + vb_0 + 2 * CO), c = *reinterpret_cast<const float4 *>(Vr + vb_p + 2 * CO);
            lerp4(p0, b, c, lx0[2], lx1[2]);
            lerp4(p1, b, c, lx0[3], lx1[3]);
        }
        float s0 = fmaf(pw4.x, fmaxf(p0[0], 0.f), fmaf(pw4.y, fmaxf(p0[1], 0.f), fmaf(pw4.z, fmaxf(p0[2], 0.f), pw4.w * fmaxf(p0[3], 0.f))));
        float s1 = fmaf(pw4.x, fmaxf(p1[0], 0.f), fmaf(pw4.y, fmaxf(p1[1], 0.f), fmaf(pw4.z, fmaxf(p1[2], 0.f), pw4.w * fmaxf(p1[3], 0.f))));
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (cg == 0 && activeB)
            __stcs(reinterpret_cast<float2 *>(out + ((size_t)n * H + 2 * k + jy) * W + 2 * M),
                   make_float2(fmaxf(s0 + pbias, 0.0f), fmaxf(s1 + pbias, 0.0f)));
    }
}

}  // namespace

extern "C" int soccdpt_depth_tail_fwd(const void *T, const float *b2, const float *pw, const float *pb, float *depth,
                                      int N, int h, int w, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(T && b2 && pw && pb && depth, "depth_tail: NULL pointer");
    SOCCDPT_REQUIRE(N >= 1 && h >= 2 && w >= 2, "depth_tail: bad shape %dx%dx%d", N, h, w);
    const int segs = (w + SW - 1) / SW, strips = (h + KP - 1) / KP;
    SOCCDPT_REQUIRE((long long)N * segs * strips < (1ll << 31), "depth_tail: grid too large");
    static soccdpt::SmemAttr configured;
    if (configured.need(SMEM_BYTES)) {
        SOCCDPT_CUDA(cudaFuncSetAttribute(depth_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    SOCCDPT_CUDA(soccdpt::launch_pdl(soccdpt::PDL_ELEMENTWISE, depth_tail_kernel, dim3((unsigned)(N * segs * strips)), dim3(256), (size_t)SMEM_BYTES,
                                     soccdpt::as_stream(stream), static_cast<const bf16 *>(T), b2, pw, pb, depth, N, h, w, segs, strips));
    return soccdpt::check_launch("depth_tail_kernel");
}
