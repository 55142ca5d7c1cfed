"""ORACLE (test infrastructure only -- never imported by the product path): numpy restatement of the reference's input
transform for uint8 frames, SOccDPT/model/loader.py:256-270 -> transforms.py:179-190 (Resize: cv2.resize INTER_CUBIC),
transforms.py:223-226 (NormalizeImage), transforms.py:235-237 (PrepareForNet).

The arithmetic lives in a third-party dependency that is not under /root/reference: OpenCV (requirements.txt:9,
`opencv-python-headless`, unpinned; 4.13.0 in this image).  Its published 8-bit bicubic algorithm
(modules/imgproc/src/resize.cpp: interpolateCubic, the fixed-point coefficient tables of cv::resize, HResizeCubic<uchar,int,short>,
VResizeCubicVec_32s8u + the VResizeCubic scalar tail with FixedPtCast<int, uchar, 22>) is restated below.

Pinned: bit-equal to cv2.resize of this image with Intel IPP switched off (cv2.ipp.setUseIPP(False)) on every size tried
(tests/test_oracle_preprocess.py), and to the unmodified reference transform run the same way.  With IPP on (the default of
the x86 wheel on Intel CPUs) cv2 itself returns +-1 in ~3 % of the elements -- closed-source code, not restated.
"""
import numpy as np

SIMD_LANES = 8     # v_int16 lanes of the 128-bit universal intrinsics the vertical pass is compiled with


def cubic_taps(dst, src):
    """Per output coordinate: clamped source indices (dst, 4) and fixed-point weights (dst, 4), int32."""
    scale = 1.0 / (dst / src)                                   # cv::resize: scale_x = 1. / inv_scale_x (double)
    d = np.arange(dst)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    x = (f - s.astype(np.float32)).astype(np.float32)
    A, one = np.float32(-0.75), np.float32(1)
    x1, xm = x + one, one - x
    c0 = ((A * x1 - np.float32(5) * A) * x1 + np.float32(8) * A) * x1 - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * xm - (A + np.float32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    c = np.stack([c0, c1, c2, c3], 1).astype(np.float32)
    w = np.clip(np.rint(c * np.float32(2048)), -32768, 32767).astype(np.int32)      # saturate_cast<short>(cvRound)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, src - 1)
    return idx, w


def cv2_resize_cubic_u8(img, dw, dh):
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC) for a uint8 HWC image (OpenCV's own code path)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, C = img.shape
    xi, xw = cubic_taps(dw, W)
    yi, yw = cubic_taps(dh, H)
    im = img.astype(np.int32)
    rows = np.zeros((H, dw, C), np.int32)
    for k in range(4):                                           # HResizeCubic: int32 accumulation
        rows += im[:, xi[:, k], :] * xw[None, :, k, None]
    # scalar tail: FixedPtCast<int, uchar, 22> on an int32 sum (wrap-around like the C code)
    acc = np.zeros((dh, dw, C), np.int32)
    with np.errstate(over="ignore"):
        for k in range(4):
            acc += rows[yi[:, k]] * yw[:, k, None, None]
        fixed = np.clip((acc + np.int32(1 << 21)) >> 22, 0, 255).astype(np.uint8)
    # SIMD body: fp32, b' = b * 2^-22, S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), round half to even, saturate
    b = yw.astype(np.float32) * np.float32(1.0 / (2048.0 * 2048.0))
    fa = rows[yi[:, 3]].astype(np.float32) * b[:, 3, None, None]
    for k in (2, 1, 0):
        fa = rows[yi[:, k]].astype(np.float32) * b[:, k, None, None] + fa
    simd = np.clip(np.rint(fa), 0, 255).astype(np.uint8)
    out = simd.reshape(dh, dw * C).copy()
    n = (dw * C) // SIMD_LANES * SIMD_LANES
    out[:, n:] = fixed.reshape(dh, dw * C)[:, n:]
    return out.reshape(dh, dw, C)


def constrain_to_multiple_of(x, multiple_of=32):
    return int(np.round(x / multiple_of) * multiple_of)          # transforms.py:105-106 ('minimal': no min / max)


def get_size(width, height, net_w, net_h, keep_aspect_ratio, multiple_of=32):
    """transforms.py:120-177 for resize_method == 'minimal'."""
    scale_height, scale_width = net_h / height, net_w / width
    if keep_aspect_ratio:
        if abs(1 - scale_width) < abs(1 - scale_height):
            scale_height = scale_width
        else:
            scale_width = scale_height
    return (constrain_to_multiple_of(scale_width * width, multiple_of),
            constrain_to_multiple_of(scale_height * height, multiple_of))


def reference_transform(img, net_w, net_h, keep_aspect_ratio):
    """The composed transform on one uint8 HWC frame -> fp32 CHW (what load_transforms(...)[0]({'image': img})['image'] is)."""
    w, h = get_size(img.shape[1], img.shape[0], net_w, net_h, keep_aspect_ratio)
    r = cv2_resize_cubic_u8(img, w, h)
    x = (r - np.array([0.5, 0.5, 0.5])) / np.array([0.5, 0.5, 0.5])      # float64, like numpy does it in the reference
    return np.ascontiguousarray(np.transpose(x, (2, 0, 1))).astype(np.float32)
