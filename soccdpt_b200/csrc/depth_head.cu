// Tail of the DPT depth head (reference SOccDPT/model/dpt.py:209-219):
//     Interpolate(x2, bilinear, align_corners=True) -> Conv2d(128, 32, 3, pad 1) -> ReLU -> Conv2d(32, 1, 1) -> ReLU
//
// conv3x3 after a bilinear upsample is linear in its input, so it is evaluated at LOW resolution first:
//     T[n, y, x, tap*32 + c] = sum_k W2[c, k, tap] * d0[n, y, x, k]          one GEMM, N = 9*32 = 288 (tcgen05 kernel)
//     out[n, Y, X]           = relu( pb + sum_c pw[c] * relu( b2[c] + sum_tap bilerp(T[..., tap*32 + c])(Y+dy, X+dx) ) )
// with taps that fall outside the upsampled image contributing zero (the conv's zero padding).  This kernel is
// the second line: per output pixel 9 taps x 4 corners x 32 channels gathered from T (bf16, L1/L2 resident),
// fp32 accumulation.  It replaces a 1.07 GB (B=64) upsampled intermediate and a narrow N=32 implicit GEMM that
// is bound by the tensor core's A-operand read; FLOPs drop 4x.
#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;
constexpr int CO = 32;          // head_features_2
constexpr int TAPS = 9;
constexpr int TC = TAPS * CO;   // 288 channels of T

__device__ __forceinline__ void fma8(float (&acc)[CO], int c0, const uint4 &u, float w) {
    const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(p[i]);
        acc[c0 + 2 * i] = fmaf(w, t.x, acc[c0 + 2 * i]);
        acc[c0 + 2 * i + 1] = fmaf(w, t.y, acc[c0 + 2 * i + 1]);
    }
}

// one thread per output pixel; blocks are 16x16 output tiles so that the ~11x11 source pixels they touch
// (70 KB of T) stay in L1
__global__ void __launch_bounds__(256)
depth_tail_kernel(const bf16 *__restrict__ T, const float *__restrict__ b2, const float *__restrict__ pw,
                  const float *__restrict__ pb, float *__restrict__ out, int N, int h, int w) {
    __shared__ float s_b2[CO], s_pw[CO];
    if (threadIdx.x < CO) {
        s_b2[threadIdx.x] = b2[threadIdx.x];
        s_pw[threadIdx.x] = pw[threadIdx.x];
    }
    __syncthreads();
    const int H = 2 * h, W = 2 * w;
    const int tiles_x = (W + 15) / 16, tiles_y = (H + 15) / 16;
    const int tile = blockIdx.x % (tiles_x * tiles_y), n = blockIdx.x / (tiles_x * tiles_y);
    const int X = (tile % tiles_x) * 16 + (threadIdx.x & 15), Y = (tile / tiles_x) * 16 + (threadIdx.x >> 4);
    if (X >= W || Y >= H) return;
    const float sh = (float)(h - 1) / (float)(H - 1), sw = (float)(w - 1) / (float)(W - 1);   // align_corners=True
    const bf16 *Tn = T + (size_t)n * h * w * TC;

    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = s_b2[c];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int Yt = Y + dy - 1;
        if (Yt < 0 || Yt >= H) continue;                    // zero padding of the 3x3 conv
        const float fy = sh * (float)Yt;
        const int y0 = (int)fy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
        const float ly = fy - (float)y0;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int Xt = X + dx - 1;
            if (Xt < 0 || Xt >= W) continue;
            const float fx = sw * (float)Xt;
            const int x0 = (int)fx, x1 = x0 + (x0 < w - 1 ? 1 : 0);
            const float lx = fx - (float)x0;
            const int tap = dy * 3 + dx;
            const uint4 *p00 = reinterpret_cast<const uint4 *>(Tn + ((size_t)y0 * w + x0) * TC + tap * CO);
            const uint4 *p01 = reinterpret_cast<const uint4 *>(Tn + ((size_t)y0 * w + x1) * TC + tap * CO);
            const uint4 *p10 = reinterpret_cast<const uint4 *>(Tn + ((size_t)y1 * w + x0) * TC + tap * CO);
            const uint4 *p11 = reinterpret_cast<const uint4 *>(Tn + ((size_t)y1 * w + x1) * TC + tap * CO);
            const float w00 = (1.0f - ly) * (1.0f - lx), w01 = (1.0f - ly) * lx, w10 = ly * (1.0f - lx), w11 = ly * lx;
#pragma unroll
            for (int q = 0; q < CO / 8; ++q) {
                fma8(acc, q * 8, __ldg(p00 + q), w00);
                fma8(acc, q * 8, __ldg(p01 + q), w01);
                fma8(acc, q * 8, __ldg(p10 + q), w10);
                fma8(acc, q * 8, __ldg(p11 + q), w11);
            }
        }
    }
    float s = pb[0];
#pragma unroll
    for (int c = 0; c < CO; ++c) s = fmaf(s_pw[c], fmaxf(acc[c], 0.0f), s);
    out[((size_t)n * H + Y) * W + X] = fmaxf(s, 0.0f);
}

}  // namespace

extern "C" int soccdpt_depth_tail_fwd(const void *T, const float *b2, const float *pw, const float *pb, float *depth,
                                      int N, int h, int w, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(T && b2 && pw && pb && depth, "depth_tail: NULL pointer");
    SOCCDPT_REQUIRE(N >= 1 && h >= 2 && w >= 2, "depth_tail: bad shape %dx%dx%d", N, h, w);
    const int tiles = ((2 * w + 15) / 16) * ((2 * h + 15) / 16);
    depth_tail_kernel<<<(unsigned)(tiles * N), 256, 0, soccdpt::as_stream(stream)>>>(static_cast<const bf16 *>(T), b2, pw, pb,
                                                                                      depth, N, h, w);
    return soccdpt::check_launch("depth_tail_kernel");
}
