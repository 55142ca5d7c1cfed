"""GPU parity of the input-pipeline kernel (csrc/preprocess.cu) through the C-ABI: bit-equal to the oracle
(oracle/preprocess_oracle.py) and to the fixture recorded from the unmodified reference transform."""
import hashlib
import os

import numpy as np
import pytest
import torch

import preprocess_oracle as P
from make_golden_preprocess import CASES, frame
from soccdpt_b200 import _cabi
from soccdpt_b200.preprocess import GpuTransform, load_gpu_transforms

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")


@pytest.mark.parametrize("name", sorted(CASES))
def test_kernel_matches_reference_fixture(name):
    g = np.load(GOLD)
    mt, H, W, seed = CASES[name]
    t, _, _ = load_gpu_transforms(mt)
    x = t(torch.from_numpy(frame(H, W, seed)).cuda())[0].cpu().numpy()
    assert list(x.shape) == g[name + "_shape"].tolist()
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest() == g[name + "_sha256"].tobytes()


@pytest.mark.parametrize("B,H,W,dh,dw", [(3, 1080, 1920, 256, 256), (2, 480, 640, 384, 384), (1, 100, 37, 64, 32),
                                          (2, 333, 517, 97, 131), (1, 64, 64, 256, 256), (4, 9, 7, 32, 32)])
def test_kernel_matches_oracle_bit_exact(B, H, W, dh, dw):
    """direct C-ABI call (any output size, also ones the transform would never pick): tails of the SIMD body, clamped taps,
    up- and down-scaling."""
    rng = np.random.default_rng(B * 1000 + H)
    imgs = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    lib = _cabi.load()
    x = torch.from_numpy(imgs).cuda()
    out = torch.empty((B, 3, dh, dw), dtype=torch.float32, device="cuda")
    ws = torch.empty(int(lib.soccdpt_preprocess_workspace_bytes(dh, dw)), dtype=torch.uint8, device="cuda")
    _cabi.check(lib.soccdpt_preprocess_fwd(x.data_ptr(), B, H, W, 3, out.data_ptr(), dh, dw, ws.data_ptr(), ws.numel(),
                                           _cabi.current_stream()), "preprocess")
    got = out.cpu().numpy()
    for b in range(B):
        ref = 2.0 * P.cv2_resize_cubic_u8(imgs[b], dw, dh).astype(np.float32).transpose(2, 0, 1) - 1.0
        assert np.array_equal(got[b], ref)


def test_transform_protocol_and_errors():
    t = GpuTransform(256, 256, False)
    img = torch.from_numpy(frame(90, 160, 5)).cuda()
    a = t({"image": img})["image"]
    b = t(img.unsqueeze(0))[0]
    assert a.shape == (3, 256, 256) and torch.equal(a, b)
    with pytest.raises(RuntimeError):
        t(img.cpu())
    lib = _cabi.load()
    assert lib.soccdpt_preprocess_fwd(None, 1, 8, 8, 3, None, 8, 8, None, 0, None) != 0
    assert lib.soccdpt_preprocess_fwd(img.data_ptr(), 1, 90, 160, 4, a.data_ptr(), 256, 256, a.data_ptr(), 1 << 20, None) != 0
