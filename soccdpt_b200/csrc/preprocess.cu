// Input pipeline of the hot path on the GPU: uint8 HWC camera frames -> the network's fp32 NCHW input.
//
// Replaces the reference's per-frame CPU transform (SOccDPT/model/loader.py:256-270 -> transforms.py:53-251):
//   Resize(cv2.resize INTER_CUBIC to (net_w, net_h))  ->  NormalizeImage(mean 0.5, std 0.5)  ->  PrepareForNet (HWC -> CHW, fp32)
// The frames the loaders hand over are uint8 and are NOT divided by 255 (bengaluru_driving_dataset.py:118-128), so the resize
// runs in OpenCV's 8-bit path and the network input is 2*v - 1 for v in 0..255 (exact in fp32).
//
// cv2.resize is a third-party dependency of the reference (requirements.txt:9, opencv-python-headless, 4.13 in this image);
// its 8-bit bicubic path (modules/imgproc/src/resize.cpp) is restated here bit for bit:
//   taps    fx = (float)((dx + 0.5) * (1 / (dst / src)) - 0.5), sx = floor(fx), t = fx - sx, source columns sx-1 .. sx+2 clamped;
//           cubic weights with A = -0.75 in fp32 (interpolateCubic, no FMA), the 4th = 1 - w0 - w1 - w2;
//           fixed point: short(cvRound(w * 2048))
//   rows    horizontal pass in int32: sum of 4 taps (HResizeCubic<uchar, int, short>)
//   columns vertical pass of the 4 int32 rows with the 4 fixed-point row weights b:
//           - the first (dst_w * 3) / 8 * 8 elements of an output row go through the SIMD kernel VResizeCubicVec_32s8u
//             (128-bit universal intrinsics of the SSE3 baseline build): fp32, b' = b * 2^-22,
//             r = S0*b0' + (S1*b1' + (S2*b2' + S3*b3')) with separately rounded multiplies and adds, round to nearest even,
//             saturate to 0..255;
//           - the remaining elements of the row through the scalar tail: (sum(S_k * b_k) + 2^21) >> 22 in int32, saturated.
// Parity: bit-equal to cv2.resize with cv2.ipp.setUseIPP(False) (tests/test_gpu_preprocess.py, oracle/preprocess_oracle.py);
// wheels that route 8-bit resizes through Intel IPP (closed source) differ from OpenCV's own code by +-1 in ~3 % of the elements.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct Taps {          // one output coordinate
    int idx[4];        // clamped source indices
    short w[4];        // fixed-point weights (x 2048)
};

__device__ __forceinline__ Taps cv_cubic_taps(int d, int dst, int src) {
    const double scale = 1.0 / ((double)dst / (double)src);                  // cv::resize: scale_x = 1. / inv_scale_x
    const float f = (float)(((double)d + 0.5) * scale - 0.5);
    const int s = (int)floorf(f);
    const float x = __fsub_rn(f, (float)s);
    const float A = -0.75f;
    const float x1 = __fadd_rn(x, 1.0f), xm = __fsub_rn(1.0f, x);
    float c[4];
    // ((A*(x + 1) - 5*A)*(x + 1) + 8*A)*(x + 1) - 4*A
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), 5.0f * A), x1), 8.0f * A), x1), 4.0f * A);
    // ((A + 2)*x - (A + 3))*x*x + 1
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.0f, x), A + 3.0f), x), x), 1.0f);
    // ((A + 2)*(1 - x) - (A + 3))*(1 - x)*(1 - x) + 1
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.0f, xm), A + 3.0f), xm), xm), 1.0f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, c[0]), c[1]), c[2]);
    Taps t;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        t.idx[k] = min(max(s - 1 + k, 0), src - 1);
        const int r = __float2int_rn(__fmul_rn(c[k], 2048.0f));             // cvRound
        t.w[k] = (short)min(max(r, -32768), 32767);                          // saturate_cast<short>
    }
    return t;
}

// tables: [dst_w] column taps, then [dst_h] row taps
__global__ void preprocess_tables_kernel(Taps *tab, int H, int W, int dh, int dw) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < dw) tab[t] = cv_cubic_taps(t, dw, W);
    else if (t < dw + dh) tab[t] = cv_cubic_taps(t - dw, dh, H);
}

// one thread per output pixel (all 3 channels): 4 rows x 4 columns x 3 channels of uint8 taps
template <int C>
__global__ void __launch_bounds__(kThreads)
preprocess_kernel(const uint8_t *__restrict__ frames, const Taps *__restrict__ tab, float *__restrict__ out, int B, int H, int W,
                  int dh, int dw) {
    const long long total = (long long)B * dh * dw;
    const int simd_elems = (dw * C) / 8 * 8;          // elements of an output row handled by cv2's SIMD kernel
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(p % dw);
        const long long q = p / dw;
        const int y = (int)(q % dh), b = (int)(q / dh);
        const Taps tx = tab[x], ty = tab[dw + y];
        const uint8_t *img = frames + (size_t)b * H * W * C;
        int S[4][C];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint8_t *row = img + (size_t)ty.idx[r] * W * C;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc += (int)__ldg(row + (size_t)tx.idx[k] * C + c) * (int)tx.w[k];
                S[r][c] = acc;
            }
        }
        const float sc = 1.0f / (2048.0f * 2048.0f);
        const float b0 = (float)ty.w[0] * sc, b1 = (float)ty.w[1] * sc, b2 = (float)ty.w[2] * sc, b3 = (float)ty.w[3] * sc;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            int v;
            if (x * C + c < simd_elems) {
                float a = __fmul_rn((float)S[3][c], b3);
                a = __fadd_rn(__fmul_rn((float)S[2][c], b2), a);
                a = __fadd_rn(__fmul_rn((float)S[1][c], b1), a);
                a = __fadd_rn(__fmul_rn((float)S[0][c], b0), a);
                v = __float2int_rn(a);
            } else {
                const int s = S[0][c] * (int)ty.w[0] + S[1][c] * (int)ty.w[1] + S[2][c] * (int)ty.w[2] + S[3][c] * (int)ty.w[3];
                v = (s + (1 << 21)) >> 22;
            }
            v = min(max(v, 0), 255);
            // NormalizeImage: (v - 0.5) / 0.5 evaluated in float64 by numpy, then cast to fp32 = 2*v - 1 exactly
            out[(((size_t)b * C + c) * dh + y) * dw + x] = (float)(2 * v - 1);
        }
    }
}

}  // namespace

extern "C" size_t soccdpt_preprocess_workspace_bytes(int dst_h, int dst_w) {
    return (size_t)(dst_h + dst_w) * sizeof(Taps);
}

extern "C" int soccdpt_preprocess_fwd(const uint8_t *frames, int batch, int H, int W, int channels, float *out, int dst_h, int dst_w,
                                      void *workspace, size_t workspace_bytes, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(frames && out && workspace, "preprocess: NULL pointer");
    SOCCDPT_REQUIRE(channels == 3, "preprocess: frames must be HWC with 3 channels (got %d)", channels);
    SOCCDPT_REQUIRE(batch >= 1 && H >= 1 && W >= 1 && dst_h >= 1 && dst_w >= 1, "preprocess: bad shape");
    SOCCDPT_REQUIRE(workspace_bytes >= soccdpt_preprocess_workspace_bytes(dst_h, dst_w), "preprocess: workspace too small");
    SOCCDPT_REQUIRE((size_t)H * W * 3 < (1ull << 31), "preprocess: frame too large");
    cudaStream_t st = soccdpt::as_stream(stream);
    Taps *tab = static_cast<Taps *>(workspace);
    preprocess_tables_kernel<<<(dst_h + dst_w + kThreads - 1) / kThreads, kThreads, 0, st>>>(tab, H, W, dst_h, dst_w);
    int rc = soccdpt::check_launch("preprocess_tables_kernel");
    if (rc != SOCCDPT_OK) return rc;
    const long long total = (long long)batch * dst_h * dst_w;
    const int blocks = (int)min((total + kThreads - 1) / kThreads, (long long)soccdpt::sm_count() * 32);
    preprocess_kernel<3><<<blocks, kThreads, 0, st>>>(frames, tab, out, batch, H, W, dst_h, dst_w);
    return soccdpt::check_launch("preprocess_kernel");
}
