/*
 * soccdpt_b200 -- C ABI of the B200 (sm_100a) implementation of SOccDPT's inference hot path.
 *
 * The reference (AdityaNG/SOccDPT) is pure Python/PyTorch: it has no FFI of its own, so the
 * entry points below are the operators its Python hot path would bind if it had one.  Each
 * declaration cites the reference code it replaces (paths under /root/reference/).  The Python
 * host side (soccdpt_b200/model/*.py) binds them with ctypes (soccdpt_b200/_cabi.py); see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the library never
 *     allocates user-visible memory (callers pass outputs and workspaces);
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t), never synchronises, keeps
 *     no mutable global state besides the last-error string -> safe under CUDA-graph capture;
 *   - return value: 0 = ok, negative = error (see SOCCDPT_E_*), text via soccdpt_last_error();
 *   - activations are NHWC ("token-major") bf16; network inputs / final outputs are fp32.
 */
#ifndef SOCCDPT_B200_H
#define SOCCDPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOCCDPT_ABI_VERSION 1

#define SOCCDPT_OK 0
#define SOCCDPT_E_INVALID (-1)   /* bad argument / unsupported shape */
#define SOCCDPT_E_CUDA (-2)      /* a CUDA runtime / driver call failed */
#define SOCCDPT_E_ARCH (-3)      /* device is not sm_100 */

typedef void *soccdpt_stream_t;  /* cudaStream_t */

int soccdpt_abi_version(void);
const char *soccdpt_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long soccdpt_launch_count(void);
int soccdpt_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* programmatic dependent launch between the kernels of a launch plan (csrc/common.cuh), per kernel family:
 * mask bits 1 = implicit-GEMM conv / linear, 2 = attention, 4 = normalisation / element-wise, 8 = post-processing;
 * 0 = plain stream order everywhere, negative = back to the default (SOCCDPT_PDL in the environment, else
 * SOCCDPT_PDL_DEFAULT).  Returns the previous mask.  No reference counterpart: the reference's eager PyTorch ops are
 * stream-ordered (every kernel boundary a full drain). */
#define SOCCDPT_PDL_DEFAULT 15
int soccdpt_set_pdl(int mask);

/* ------------------------------------------------------------------ sparse occupancy outputs
 * The reference's occupancy_grid_to_points (SOccDPT/utils/__init__.py:532-568, numpy on the CPU) on the device:
 * rows (x, y, z, class) as f64 with x = f32(i / G0 * occ_shape[0]) etc. for every cell >= 0.5, ordered by class,
 * then by (i, j, k).  Input is the bit-packed mask (layout: SOCCDPT_OCC_PACKED above), straight from the voxeliser's
 * workspace or packed from a dense grid (G0,G1,G2,C) by soccdpt_grid_pack_fwd.
 * soccdpt_occupancy_points_fwd writes the number of rows to *count (device memory); with points == NULL it only counts
 * (size the output, then call again); rows beyond `capacity` are dropped. */
size_t soccdpt_occupancy_mask_bytes(const int grid[3]);
size_t soccdpt_occupancy_points_workspace_bytes(const int grid[3], int num_classes);
int soccdpt_grid_pack_fwd(const float *grid_dense, const int grid[3], int num_classes, uint32_t *mask,
                          soccdpt_stream_t stream);
int soccdpt_occupancy_points_fwd(const uint32_t *mask, const int grid[3], const float occ_shape[3], int num_classes,
                                 double *points, long long capacity, long long *count, void *workspace,
                                 size_t workspace_bytes, soccdpt_stream_t stream);

/* Counting voxeliser ("count(threshold) / argmax" modes): the reference's ground-truth generator
 * OccupancyProcessor.transform_points_to_occupancy_grid_vect, SOccDPT/datasets/bdd_helper.py:289-362 (numpy on the CPU).
 *   points     f64 or f32 [n,3] camera points (points_f64 selects; numpy's promotion rules decide the arithmetic:
 *              f64 points -> all in fp64, f32 points -> fp32 division, fp64 product), NaN / inf rows are skipped
 *   semantics  i64 or i32 [n] class ids; negative ids index from the end like numpy; ids outside [-C, C) are counted in
 *              *bad_class (numpy raises IndexError there) and skipped
 *   counts     i32 [G0,G1,G2,C], ACCUMULATED into (the caller zeroes it; several calls = several point sets)
 * soccdpt_voxel_count_finish_fwd thresholds the counts: grid_gt u8 [G0,G1,G2,C] = count > threshold (the reference's
 * "occupancy_grid", :357), mask_ge = bit-packed cells with count >= threshold (layout of SOCCDPT_OCC_PACKED; feed it to
 * soccdpt_occupancy_points_fwd for the reference's "occupancy_points", :340-355), labels u8 [G0,G1,G2] = 0 for empty
 * cells, else 1 + argmax_c count (first maximum).  Any of the three outputs may be NULL. */
int soccdpt_voxel_count_fwd(const void *points, int points_f64, const void *semantics, int semantics_i64, long long n,
                            const int grid[3], const float occ_shape[3], int num_classes, int32_t *counts,
                            unsigned long long *bad_class, soccdpt_stream_t stream);
int soccdpt_voxel_count_finish_fwd(const int32_t *counts, const int grid[3], int num_classes, float threshold,
                                   uint8_t *grid_gt, uint32_t *mask_ge, uint8_t *labels, soccdpt_stream_t stream);

/* ------------------------------------------------------------------ input pipeline
 * The reference's per-frame CPU transform, SOccDPT/model/loader.py:256-270 -> transforms.py:53-251:
 * cv2.resize(INTER_CUBIC) of a uint8 HWC frame to (dst_w, dst_h), NormalizeImage(0.5, 0.5), PrepareForNet.
 *   frames u8  [batch, H, W, 3] (what cv2.imread / the dataset loaders produce; NOT divided by 255)
 *   out    f32 [batch, 3, dst_h, dst_w] = 2 * resized - 1, the tensor the network consumes
 * Bit-equal to OpenCV's own 8-bit bicubic code (cv2.ipp.setUseIPP(False)); see csrc/preprocess.cu. */
size_t soccdpt_preprocess_workspace_bytes(int dst_h, int dst_w);
int soccdpt_preprocess_fwd(const uint8_t *frames, int batch, int H, int W, int channels, float *out, int dst_h,
                           int dst_w, void *workspace, size_t workspace_bytes, soccdpt_stream_t stream);

/* ------------------------------------------------------------------ post-processing (A8/A9)
 * Geometry constants of the reference base class, SOccDPT/model/SOccDPT.py:134-228.  All
 * values are prepared by the host exactly as the reference prepares them (fp32 casts,
 * rotation matrices of SOccDPT.py:82-111, occupancy_shape of SOccDPT.py:175-181). */
typedef struct {
    float fx, fy, cx, cy;        /* Camera.* of the calib YAML, cast to fp32 */
    int height, width;           /* camera resolution the maps are resized to */
    int num_classes;             /* <= 4 (the reference only works for 3) */
    int grid[3];                 /* occupancy grid size in voxels */
    float occ_shape[3];          /* grid / scale, metres, fp32 */
    float pc_scale[3];           /* applied to POINTS 0,1,2 of each frame (SOccDPT.py:351-353) */
    float pc_shift[3];
    float rot[27];               /* Ra, Rb, Rc row-major; points are multiplied p @ Ra @ Rb @ Rc */
    int scalar_div_by_reciprocal; /* device convention of "tensor / host scalar" in X = (V - cx) * depth / fx (SOccDPT.py:311-313):
                                     0 = true division (ATen's CPU kernel; the convention of the reference fixtures),
                                     1 = multiply by the fp32 reciprocal of the scalar (ATen's CUDA kernel): the reference's own
                                     eager path differs between devices by 1-2 ulps in X / Y (profiles/r2_device_convention.jsonl).
                                     With 1 the caller also passes rotation matrices built from the CUDA cos / sin. */
    float rcp_fx, rcp_fy;        /* used when scalar_div_by_reciprocal: fp32(1.0 / fx), fp32(1.0 / fy) with the reciprocal taken in
                                     DOUBLE on the yaml's double (what ATen's CUDA div kernel does with a Python-float scalar;
                                     it is NOT 1.0f / fp32(fx): for fx = 1250.6 the two differ by one ulp) */
} soccdpt_geometry_t;

#define SOCCDPT_OCC_REFERENCE_UNION 0 /* reference semantics: OR over the batch, written to every b */
#define SOCCDPT_OCC_PER_FRAME 1       /* extension: each frame gets only its own voxels */
#define SOCCDPT_OCC_PACKED 2          /* flag, OR-ed in: keep the bit-packed voxel mask at the start of the workspace as
                                         an output (grid may then be NULL): uint32 words, voxel v = (i*G1 + j)*G2 + k lives in
                                         word v >> 3, class c is bit (v & 7) * 4 + c; one mask per call, or per frame in
                                         PER_FRAME mode (soccdpt_occupancy_mask_bytes() apart) */

/* bytes of scratch the two calls below need: bit-packed voxel mask (+ per-row / per-column resize tables) */
size_t soccdpt_voxel_workspace_bytes(const soccdpt_geometry_t *g, int batch, int mode);

/* Replaces SOccDPT.get_semantic_occupancy from the clamp onwards + points_to_occupancy_grid
 * (SOccDPT/model/SOccDPT.py:288-372, :374-463) for maps ALREADY at camera resolution.
 *   inv_depth_up (B,H,W) f32 in/out (clamped in place like the reference, :289)
 *   seg_up       (B,C,H,W) f32 in
 *   points       (B,H,W,3) f32 out (un-rotated, incl. the three altered points)
 *   grid         (B,G0,G1,G2,C) f32 out, or NULL for compute_occ=False
 * Bit-exact against the reference's CPU path. */
int soccdpt_voxelize_fwd(float *inv_depth_up, const float *seg_up, int batch,
                         const soccdpt_geometry_t *g, float *points, float *grid, int mode,
                         void *workspace, size_t workspace_bytes, soccdpt_stream_t stream);

/* Same, fused with the two resizes of SOccDPT.py:270-282 (bicubic align_corners=False for the
 * inverse depth, legacy nearest for the classes) from network resolution (h,w):
 *   inv_depth (B,h,w) f32, seg (B,C,h,w) f32  ->  inv_depth_up, seg_up (B,C,H,W), points, grid */
int soccdpt_postprocess_fwd(const float *inv_depth, const float *seg, int batch, int h, int w,
                            const soccdpt_geometry_t *g, float *inv_depth_up, float *seg_up,
                            float *points, float *grid, int mode, void *workspace,
                            size_t workspace_bytes, soccdpt_stream_t stream);

/* ------------------------------------------------------------------ dense contractions
 * One implicit-GEMM operator covers every convolution / linear layer of the path:
 *   scratch.layerN_rn           SOccDPT/model/blocks.py:155-191
 *   ResidualConvUnit_custom     SOccDPT/model/blocks.py:391-414   (bias, ReLU, residual adds)
 *   FeatureFusionBlock out_conv SOccDPT/model/blocks.py:472-497   (1x1)
 *   depth head                  SOccDPT/model/dpt.py:199-219      (incl. the fused 32->1 projection)
 *   seg head                    SOccDPT/model/SOccDPT.py:660-674  (BN folded, fused 256->3 projection)
 *   timm qkv/proj/fc1/fc2/reduction linears (K=1x1 over a (1,1,M,K) "image")
 * y[n,h,w,:] = act( sum_{kh,kw,c} x[n,h*s+kh-pad,w*s+kw-pad,c]   (pad = KH/2 - pad_trim) * wgt[:,kh,kw,c] + bias ) + res1 + res2
 * stride s in {1,2}, zero padding (pad before, whatever is needed after), fp32 accumulation, bf16 storage. */
#define SOCCDPT_ACT_NONE 0
#define SOCCDPT_ACT_RELU 1
#define SOCCDPT_ACT_GELU 2 /* exact erf GELU (timm Mlp) */

typedef struct {
    const void *x;        /* bf16 [N,H,W,Cin] */
    const void *wgt;      /* bf16 [Cout][KH*KW][Cin] */
    const float *bias;    /* f32 [Cout] or NULL */
    const void *res1;     /* bf16 [N,H,W,Cout] or NULL, added after the activation */
    const void *res2;     /* bf16 [N,H,W,Cout] or NULL */
    void *y;              /* bf16 [N,H,W,Cout] or NULL */
    void *y_relu;         /* bf16 relu(y), second output for consumers that need both, or NULL */
    int N, H, W, Cin, Cout, KH, KW;
    int act;              /* SOCCDPT_ACT_* applied to (acc + bias) */
    /* optional fused projection epilogue: p = proj_w @ y_row + proj_b (needs Cout <= 256) */
    const float *proj_w;  /* f32 [proj_n][Cout] or NULL */
    const float *proj_b;  /* f32 [proj_n] */
    float *proj_out;      /* f32 [N*H*W][proj_n] */
    int proj_n;           /* 1..4 */
    int proj_relu;        /* ReLU on the projection */
    int stride;           /* 1 or 2 (0 is read as 1); outputs are [N, ceil(H/stride), ceil(W/stride), Cout] */
    int pad_trim;         /* rows/columns of zero padding REMOVED before the first row/column: the padding in front is
                             KH/2 - pad_trim (0 = symmetric torch padding).  TF-"SAME" 3x3 stride-2 convs on even
                             inputs (timm StdConv2dSame) pad 0 in front / 1 behind: pad_trim = 1. */
    /* optional cosine-attention epilogue of a SwinV2 qkv linear (timm WindowAttention.forward: F.normalize(q) @
     * F.normalize(k)^T * logit_scale): with qk_heads > 0 the output columns [0, 64*qk_heads) = (q | k), 32 per head, are
     * L2-normalised per head from the fp32 accumulator (x / max(|x|, 1e-12)) and the q columns [0, 32*qk_heads) are also
     * multiplied by qk_scale[head]; the v columns behind them are stored as they are.  Plain outputs only (no residual,
     * projection or activation); consumer: soccdpt_window_attention_normed_fwd. */
    const float *qk_scale; /* f32 [qk_heads] or NULL */
    int qk_heads;
    /* optional third residual, UP-SAMPLED on the fly (FeatureFusionBlock_custom.forward, SOccDPT/model/blocks.py:476-487:
     * output = xs[0] + resConfUnit1(xs[1]) where xs[0] is the previous block's
     * interpolate(scale_factor=2, mode="bilinear", align_corners=True) output): up_src is the LOW-resolution map
     * bf16 [N, up_h, up_w, Cout] with H == 2*up_h and W == 2*up_w; the epilogue adds its bilinear x2 (align_corners=True)
     * interpolation at the output pixel, evaluated in fp32 -- the up-sampled tensor is never written.  Needs stride 1,
     * Cout % 32 == 0, no fused projection. */
    const void *up_src;
    int up_h, up_w;
} soccdpt_conv_t;

/* tcgen05 / TMEM / TMA kernel (the product path) */
int soccdpt_conv_fwd(const soccdpt_conv_t *c, soccdpt_stream_t stream);
/* plain CUDA-core kernel with identical semantics: on-device cross-check used by the tests */
int soccdpt_conv_ref_fwd(const soccdpt_conv_t *c, soccdpt_stream_t stream);

/* ------------------------------------------------------------------ encoder pieces (timm SwinV2 0.6.12)
 * timm PatchEmbed: conv4x4 s4 + bias + LayerNorm (eps 1e-5); x f32 NCHW -> tokens bf16 [B,H/4*W/4,E]
 * (+ an optional fp32 copy that seeds the fp32 residual stream, or NULL) */
int soccdpt_patch_embed_fwd(const float *x, const float *w, const float *b, const float *ln_w,
                            const float *ln_b, void *tokens, float *tokens_f32, int batch, int H, int W,
                            int E, soccdpt_stream_t stream);

/* timm WindowAttention (v2, cosine) incl. window partition / cyclic shift / reverse:
 *   qkv   bf16 [B, Hs*Ws, 3*C] (q|k|v, each heads x 32), biases already added by the qkv GEMM
 *   bias  f32 [heads][(2ws-1)^2] = 16*sigmoid(cpb_mlp(coords_table)), timm's relative-position table BEFORE the
 *         [rel_idx] expansion (input independent, baked at load); entry (qy-ky+ws-1)*(2ws-1) + (qx-kx+ws-1)
 *   scale f32 [heads]        = exp(min(logit_scale, ln 100))
 *   out   bf16 [B, Hs*Ws, C]
 * window ws x ws (N = ws*ws tokens), shift in {0, ws/2}; the -100 shift mask is generated from
 * region ids exactly as timm's attn_mask buffer. head_dim must be 32. */
int soccdpt_window_attention_fwd(const void *qkv, const float *bias, const float *scale, void *out,
                                 int batch, int Hs, int Ws, int C, int heads, int ws, int shift,
                                 soccdpt_stream_t stream);

/* The same operator for 16x16 windows on operands the qkv linear has already normalised (soccdpt_conv_t.qk_heads):
 *   qkvn  bf16 [B, Hs*Ws, 3*C] = normalize(q) * scale * log2(e) | normalize(k) | v
 *   bias, scale, out as above; shift in {0, 8}
 * TMA-fed (2 x 2 boxes of 8 x 8 tokens per window: the cyclic shift is a box coordinate), persistent, pipelined across
 * (window, head) items; one-pass softmax against the analytic logit bound, so every scale[h] must satisfy
 * 2.01 * scale + 16 < 80 (the caller checks; soccdpt_window_attention_fwd has the exact row-max path for larger scales). */
int soccdpt_window_attention_normed_fwd(const void *qkvn, const float *bias, const float *scale, void *out,
                                        int batch, int Hs, int Ws, int C, int heads, int shift,
                                        soccdpt_stream_t stream);

/* y = (res ? res : 0) + LayerNorm(t) over the last dim (eps), bf16 in/out, fp32 math.
 * Swin res-post-norm: x + norm1(attn(x)), x + norm2(mlp(x)); PatchMerging norm (res = NULL). */
int soccdpt_layernorm_fwd(const void *t, const void *res, const float *gamma, const float *beta,
                          void *y, long long rows, int C, float eps, soccdpt_stream_t stream);
/* Same with the residual stream kept in fp32: master = (accumulate ? master : 0) + LayerNorm(t), updated in
 * place; y = bf16(master) is the copy the next GEMM reads. */
int soccdpt_layernorm_master_fwd(const void *t, float *master, int accumulate, const float *gamma,
                                 const float *beta, void *y, long long rows, int C, float eps,
                                 soccdpt_stream_t stream);

/* Fused tail of one branch of a timm SwinTransformerV2Block (res-post-norm; SURVEY.md App. A.1, driven by the reference
 * through SOccDPT/model/backbones/swin_common.py:16-27 / utils.py:64-81):
 *   two-GEMM mode (w1 != NULL), the MLP branch  x = x + norm2(mlp(x)):
 *       master += LayerNorm( GELU(x @ w1^T + b1) @ w2^T + b2 ) * gamma + beta ;  y = bf16(master)
 *   one-GEMM mode (w1 == NULL), the attention branch after the window attention  x = x + norm1(proj(attn)):
 *       master += LayerNorm( x @ w2^T + b2 ) * gamma + beta ;  y = bf16(master)
 * The hidden activation and the branch output never leave the SM (TMEM); fp32 accumulation, fp32 LayerNorm, fp32
 * residual stream.  y may alias x.  C in [32, 256] (multiple of 32), HID a multiple of 128. */
typedef struct {
    const void *x;        /* bf16 [M, K1]   rows entering the first GEMM */
    const void *w1;       /* bf16 [HID, K1] fc1 weight, or NULL for one-GEMM mode */
    const float *b1;      /* f32 [HID] */
    const void *w2;       /* bf16 [C, HID] fc2 weight (two-GEMM mode) or [C, K1] proj weight (one-GEMM mode) */
    const float *b2;      /* f32 [C] */
    const float *gamma;   /* f32 [C] LayerNorm weight */
    const float *beta;    /* f32 [C] LayerNorm bias */
    float *master;        /* f32 [M, C] residual stream, updated in place */
    void *y;              /* bf16 [M, C] copy of the updated stream (the next GEMM's operand) */
    long long M;
    int K1, HID, C;
    float eps;
} soccdpt_block_tail_t;
int soccdpt_swin_block_tail_fwd(const soccdpt_block_tail_t *a, soccdpt_stream_t stream);

/* timm PatchMerging gather: [B,H,W,C] -> [B,H/2,W/2,4C] in (x0,x1,x2,x3) = (0,0),(1,0),(0,1),(1,1) order */
int soccdpt_patch_merge_gather_fwd(const void *x, void *y, int batch, int H, int W, int C,
                                   soccdpt_stream_t stream);

/* ------------------------------------------------------------------ decoder glue
 * bilinear, align_corners=True, NHWC bf16 [N,h,w,C] -> [N,H,W,C] (blocks.py:488-493, dpt.py:209) */
int soccdpt_upsample_bilinear_fwd(const void *x, void *y, int N, int h, int w, int H, int W, int C,
                                  soccdpt_stream_t stream);

/* seg head tail (SOccDPT.py:671-673): logits f32 [N,h,w,P] -> bilinear x2 (align_corners=True)
 * -> sigmoid (act=0) or 0.5*tanh+0.5 (act=1) -> f32 NCHW [N,P,2h,2w] */
int soccdpt_seg_finish_fwd(const float *logits, float *seg, int N, int h, int w, int P, int act,
                           soccdpt_stream_t stream);

/* depth head tail (dpt.py:209-219): Interpolate x2 (bilinear, align_corners=True) -> Conv2d(128,32,3) -> ReLU ->
 * Conv2d(32,1,1) -> ReLU, evaluated from T = the nine 128->32 tap matrices applied at LOW resolution
 * (T bf16 [N,h,w,9*32], channel = tap*32 + c, produced by soccdpt_conv_fwd with the re-packed weights):
 * depth f32 [N,2h,2w] = relu(pb + pw . relu(b2 + sum_tap bilerp(T_tap)(Y+dy, X+dx))), zero outside the image */
int soccdpt_depth_tail_fwd(const void *T, const float *b2, const float *pw, const float *pb, float *depth,
                           int N, int h, int w, soccdpt_stream_t stream);

/* Diagnostic: sweeps all 2^32 fp32 bit patterns on the device and counts where the fast exact-arithmetic sequences of the
 * voxeliser differ from the IEEE operation they replace (the reference's torch ops, SOccDPT.py:288-316, 393-437):
 * mismatches[0]: MUFU.RCP + Newton step vs correctly rounded 1/x; mismatches[1..5]: fp64-reciprocal division by fx, fy,
 * occ_shape[0..2] vs correctly rounded x / c.  All six must be 0.  Synchronises the stream; ~0.1 s. */
int soccdpt_selftest_exact_math(const soccdpt_geometry_t *g, unsigned long long mismatches[6], soccdpt_stream_t stream);

/* dtype plumbing for the boundary: f32 <-> bf16 round-to-nearest-even, n elements */
int soccdpt_f32_to_bf16(const float *x, void *y, long long n, soccdpt_stream_t stream);
int soccdpt_bf16_to_f32(const void *x, float *y, long long n, soccdpt_stream_t stream);

/* ---- ViT-hybrid encoder only (dpt_hybrid_384 = timm vit_base_resnet50_384; reference
 * SOccDPT/model/backbones/vit.py:245-258 builds it, :44-85 forward_flex runs it, :147-242 post-processes the taps) */

/* ResNetV2 stem: StdConv2dSame(3, 64, 7, stride 2) on the f32 NCHW frame; w f32 [64][3][7][7] ALREADY weight-standardised.
 * y bf16 NHWC [batch, ceil(H/2), ceil(W/2), 64] */
int soccdpt_stem_conv7_fwd(const float *x, const float *w, void *y, int batch, int H, int W, soccdpt_stream_t stream);

/* GroupNormAct(32 groups) on NHWC bf16 [batch, HW, C]: y = [relu]( gn(x) * gamma + beta [+ shortcut] ).
 * stats_scratch: batch*64 doubles of device memory (zeroed by the call).  C = 64, 128 or a multiple of 256. */
int soccdpt_groupnorm_fwd(const void *x, const float *gamma, const float *beta, const void *shortcut, void *y, int batch,
                          int HW, int C, float eps, int relu, void *stats_scratch, soccdpt_stream_t stream);

/* MaxPool2dSame(3, stride 2), NHWC bf16 [batch,H,W,C] -> [batch, ceil(H/2), ceil(W/2), C] (padding value -inf) */
int soccdpt_maxpool3s2_fwd(const void *x, void *y, int batch, int H, int W, int C, soccdpt_stream_t stream);

/* ---- fp32-storage PARITY mode of the hybrid encoder's ResNetV2 trunk (csrc/trunk_fp32.cu; NetworkEngine(trunk_fp32=True) /
 * SOCCDPT_HYBRID_TRUNK=fp32).  Replaces, for that mode only, the bf16 tensor-core path of timm's StdConv2dSame / GroupNormAct /
 * MaxPool2dSame as reached from SOccDPT/model/backbones/vit.py:147-258; CUDA-core fp32 FMA, fp32 activations and weights.
 * conv: NHWC f32 x [batch,H,W,Cin], w [Cout][k][k][Cin] (already weight-standardised), TF "SAME" padding, stride 1 or 2,
 *       y [batch, ceil(H/stride), ceil(W/stride), Cout], no bias. */
int soccdpt_conv_f32_fwd(const float *x, const float *w, float *y, int batch, int H, int W, int Cin, int Cout, int k, int stride,
                         soccdpt_stream_t stream);
/* GroupNorm(32 groups, eps) [+ shortcut] [+ ReLU] on NHWC f32; y may alias x */
int soccdpt_groupnorm_f32_fwd(const float *x, const float *gamma, const float *beta, const float *shortcut, float *y, int batch,
                              int HW, int C, float eps, int relu, soccdpt_stream_t stream);
/* MaxPool2dSame(3, stride 2) on NHWC f32 (padding value -inf) */
int soccdpt_maxpool3s2_f32_fwd(const float *x, float *y, int batch, int H, int W, int C, soccdpt_stream_t stream);
/* f32 [batch,C,HW] -> f32 [batch,HW,C] (the network input for the fp32 trunk) */
int soccdpt_nchw_to_nhwc_f32(const float *x, float *y, int batch, int C, int HW, soccdpt_stream_t stream);

/* tokens bf16 [batch, 1+L, D] (+ optional f32 copy) = cat(cls f32 [D], patches bf16 [batch, L, D]) + pos f32 [1+L, D]
 * (reference vit.py:66-80; the position embedding is used at its native 24x24 grid: 384x384 frames only) */
int soccdpt_vit_tokens_fwd(const void *patches, const float *cls, const float *pos, void *tokens, float *tokens_f32, int batch,
                           int L, int D, soccdpt_stream_t stream);

/* Pre-norm residual step of a timm ViT block (x = x + branch(LN(x)), eps 1e-6): master f32 [rows, C] += t (bf16, may be
 * NULL); y = LayerNorm(master) bf16 (NULL: skipped); stream_bf16 = bf16(master) (NULL: skipped; the hooked block outputs
 * of reference vit.py:179-219).  C <= 1024. */
int soccdpt_prenorm_fwd(const void *t, float *master, const float *gamma, const float *beta, void *y, void *stream_bf16,
                        long long rows, int C, float eps, soccdpt_stream_t stream);

/* ProjectReadout input (reference backbones/utils.py:27-40): feats bf16 [batch, L, 2D] = cat(tokens[:,1:], tokens[:,0] expanded) */
int soccdpt_readout_concat_fwd(const void *tokens, void *feats, int batch, int L, int D, soccdpt_stream_t stream);

/* Global multi-head attention of the ViT blocks: qkv bf16 [batch, N, 3*heads*64] (q | k | v), out bf16 [batch, N, heads*64],
 * softmax(q k^T / sqrt(64)) v with fp32 softmax.  head_dim must be 64. */
int soccdpt_global_attention_fwd(const void *qkv, void *out, int batch, int N, int heads, int head_dim,
                                 soccdpt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SOCCDPT_B200_H */
