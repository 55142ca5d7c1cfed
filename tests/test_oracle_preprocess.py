"""Pins oracle/preprocess_oracle.py (CPU, no GPU): against the committed fixture recorded from the UNMODIFIED reference
transform, against cv2 of this image (OpenCV's own code: bit-equal; the IPP path: +-1), and -- when /root/reference exists --
against the reference transform run live."""
import hashlib
import os

import numpy as np
import pytest

import preprocess_oracle as P
import ref_env
from make_golden_preprocess import CASES, frame

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")
_CFG = {"dpt_swin2_tiny_256": (256, 256, False), "dpt_hybrid_384": (384, 384, True)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_fixture(name):
    g = np.load(GOLD)
    mt, H, W, seed = CASES[name]
    x = P.reference_transform(frame(H, W, seed), *_CFG[mt])
    assert x.dtype == np.float32 and list(x.shape) == g[name + "_shape"].tolist()
    assert np.array_equal(x[:, x.shape[1] // 2, :], g[name + "_row"])
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest() == g[name + "_sha256"].tobytes()


@pytest.mark.parametrize("H,W,dh,dw", [(1080, 1920, 256, 256), (480, 640, 384, 384), (100, 37, 64, 32), (333, 517, 97, 131),
                                       (64, 64, 256, 256), (9, 7, 32, 32)])
def test_oracle_vs_cv2_live(H, W, dh, dw):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(H * 7 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    mine = P.cv2_resize_cubic_u8(img, dw, dh)
    was = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        own = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC)
        cv2.ipp.setUseIPP(True)
        ipp = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC)
    finally:
        cv2.ipp.setUseIPP(was)
    assert np.array_equal(mine, own)                                         # OpenCV's own code: bit-equal
    assert np.abs(mine.astype(int) - ipp.astype(int)).max() <= 1             # Intel IPP (if the wheel uses it): +-1


@pytest.mark.skipif(not ref_env.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("mt,H,W", [("dpt_swin2_tiny_256", 540, 960), ("dpt_swin2_base_384", 1080, 1920), ("dpt_hybrid_384", 720, 1280)])
def test_oracle_and_mirror_vs_reference_live(mt, H, W):
    cv2 = pytest.importorskip("cv2")
    ref_loader, _ = ref_env.import_reference()
    from soccdpt_b200 import load_transforms
    from soccdpt_b200.preprocess import transform_config
    img = frame(H, W, 11)
    was = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        t_ref, w_ref, h_ref = ref_loader.load_transforms(model_type=mt)
        ref = t_ref({"image": img})["image"]
        t_mine, w, h = load_transforms(mt)
        mirror = t_mine({"image": img})["image"]
    finally:
        cv2.ipp.setUseIPP(was)
    assert (w, h) == (w_ref, h_ref)
    assert mirror.dtype == ref.dtype and np.array_equal(mirror, ref)        # host-side mirror of load_transforms
    assert np.array_equal(P.reference_transform(img, *transform_config(mt)), ref)


def test_get_size_mirror_matches_oracle():
    from soccdpt_b200.preprocess import get_size
    for (W, H) in [(1920, 1080), (1280, 720), (640, 480), (517, 333), (100, 3000), (48, 48)]:
        for (nw, nh) in [(256, 256), (384, 384), (320, 320)]:
            for keep in (False, True):
                assert get_size(W, H, nw, nh, keep) == P.get_size(W, H, nw, nh, keep)
