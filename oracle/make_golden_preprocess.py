"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/preprocess.npz from the UNMODIFIED reference transform
(/root/reference/SOccDPT/model/loader.py load_transforms -> transforms.py) run on seeded uint8 frames with OpenCV's own
resize code (cv2.ipp.setUseIPP(False); see oracle/preprocess_oracle.py for why).  Only sha256 digests + a few raw rows are
stored; the tests regenerate the frames from the seeds.

    python oracle/make_golden_preprocess.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_env  # noqa: E402

# name: (model_type, H, W, seed)
CASES = {
    "tiny_1080p": ("dpt_swin2_tiny_256", 1080, 1920, 0),
    "tiny_720p": ("dpt_swin2_tiny_256", 720, 1280, 1),
    "hybrid_1080p": ("dpt_hybrid_384", 1080, 1920, 2),      # keep_aspect_ratio -> 384 x 672
    "tiny_ragged": ("dpt_swin2_tiny_256", 333, 517, 3),
}


def frame(H, W, seed):
    """Seeded uint8 frame with smooth structure + noise + saturated patches (exercises the 0 / 255 clamps)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = 127 + 100 * np.sin(xx / 37.0 + seed)[..., None] * np.cos(yy / 23.0)[..., None] * np.array([1.0, 0.7, -0.8])
    img = np.clip(base + rng.normal(0, 40, (H, W, 3)), 0, 255).astype(np.uint8)
    img[: H // 8, : W // 8] = 255
    img[-H // 8:, -W // 8:] = 0
    img[H // 2: H // 2 + 3, :] = rng.integers(0, 2, (3, W, 3)) * 255          # 0 / 255 stripes: overshoot both ways
    return img


def main():
    import cv2
    cv2.ipp.setUseIPP(False)
    ref_loader, _ = ref_env.import_reference()
    load_transforms = ref_loader.load_transforms
    out = {}
    for name, (mt, H, W, seed) in CASES.items():
        t, _, _ = load_transforms(model_type=mt)
        x = t({"image": frame(H, W, seed)})["image"]
        assert x.dtype == np.float32
        out[name + "_shape"] = np.array(x.shape)
        out[name + "_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest(), np.uint8)
        out[name + "_row"] = x[:, x.shape[1] // 2, :].copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), cv2_version=np.array(cv2.__version__), **out)
    print("wrote tests/golden/preprocess.npz", {k: v.tolist() for k, v in out.items() if k.endswith("_shape")})


if __name__ == "__main__":
    main()
