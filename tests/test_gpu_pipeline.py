"""FrameStream (the public host-frame path): results equal the direct calls, for network-resolution fp32 batches and for
uint8 camera frames that go through the input-pipeline kernel on the device."""
import numpy as np
import pytest
import torch

import golden_util as GU
from make_golden_preprocess import frame
from soccdpt_b200 import SOccDPT_versions, load_model
from soccdpt_b200.pipeline import FrameStream
from soccdpt_b200.preprocess import load_gpu_transforms
from soccdpt_b200.synthetic import synthetic_frames, write_calib_yaml

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net(tmp_path_factory):
    yml = write_calib_yaml(str(tmp_path_factory.mktemp("calib") / "c.yaml"))
    net = load_model(arch=SOccDPT_versions[3],
                     model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                                       camera_intrinsics_yaml=yml, model_type="dpt_swin2_tiny_256"),
                     device=torch.device("cuda"), model_path=None, model_type="dpt_swin2_tiny_256")
    net.load_state_dict(GU.tiny_state_dict(0), strict=True)
    return net.eval()


def _direct(net, x):
    with torch.no_grad():
        out = net(x)
        d, s = net.network(x)
        return d.clone().cpu(), s.clone().cpu(), out[3][0].clone().cpu()


@pytest.mark.parametrize("overlap_post", [False, True])      # True: post-processing of batch i on its own stream
def test_stream_of_fp32_batches_equals_direct_calls(net, overlap_post):
    B = 2
    batches = [synthetic_frames(B, 256, 20 + i).pin_memory() for i in range(5)]      # > 2: the buffers are recycled
    fs = FrameStream(net, B, overlap_post=overlap_post)
    got = [(r.index, r.inv_depth.clone(), r.segmentation.clone(), r.occupancy.clone()) for r in fs.run(batches)]
    assert [g[0] for g in got] == [0, 1, 2, 3, 4]
    for (_, d, s, g), xb in zip(got, batches):
        d0, s0, g0 = _direct(net, xb.cuda())
        assert torch.equal(d, d0) and torch.equal(s, s0) and torch.equal(g, g0)


def test_stream_of_uint8_camera_frames_uses_the_gpu_transform(net):
    B, H, W = 2, 270, 480
    batches = [torch.from_numpy(np.stack([frame(H, W, 40 + 2 * i + j) for j in range(B)])).pin_memory() for i in range(3)]
    fs = FrameStream(net, B, camera_frames=(H, W))
    assert fs.h2d_bytes == B * H * W * 3
    t, _, _ = load_gpu_transforms("dpt_swin2_tiny_256")
    for r, fb in zip(fs.run(batches), batches):
        d0, s0, g0 = _direct(net, t(fb.cuda()))
        assert torch.equal(r.inv_depth, d0) and torch.equal(r.segmentation, s0) and torch.equal(r.occupancy, g0)


def test_packed_result_and_uint8_network_resolution_frames(net):
    """The e2e form bench.py times: uint8 frames at network resolution in, bf16 maps + bit-packed occupancy mask out."""
    from soccdpt_b200.occupancy import pack_grid
    B = 2
    g = torch.Generator().manual_seed(7)
    batches = [torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(3)]
    fs = FrameStream(net, B, device="cuda", frames="u8", frame_shape=(256, 256), result="packed")   # "cuda": no index
    assert fs.h2d_bytes == B * 256 * 256 * 3
    assert fs.d2h_bytes == B * 256 * 256 * 2 * (1 + 3) + 256 * 256 * 32 // 8 * 4
    t, _, _ = load_gpu_transforms("dpt_swin2_tiny_256")
    for r, fb in zip(fs.run(batches), batches):
        x = t(fb.cuda())
        assert torch.equal(x, 2.0 * fb.cuda().permute(0, 3, 1, 2).float() - 1.0)       # identity resize: normalisation only
        d0, s0, g0 = _direct(net, x)
        assert r.inv_depth.dtype == torch.bfloat16 and torch.equal(r.inv_depth, d0.bfloat16())
        assert torch.equal(r.segmentation, s0.bfloat16())
        assert torch.equal(r.occupancy, pack_grid(g0.cuda()).cpu())
    assert net.occupancy_output == "dense"               # the model's own setting is restored


def test_per_frame_mode_returns_every_grid_and_settings_are_validated(net, tmp_path):
    B = 2
    xb = synthetic_frames(B, 256, 33).pin_memory()
    net.occupancy_mode = "per_frame"
    try:
        fs = FrameStream(net, B)
        (r,) = list(fs.run([xb]))
        with torch.no_grad():
            ref = net(xb.cuda())[3]
        assert r.occupancy.shape == ref.shape and torch.equal(r.occupancy, ref.cpu())
    finally:
        net.occupancy_mode = "reference_union"
    net.compute_occ = False
    try:
        with pytest.raises(ValueError):
            FrameStream(net, B)
    finally:
        net.compute_occ = True


def test_plan_cache_is_bounded_and_device_spelling_is_normalised(net):
    eng = net.engine()
    with torch.no_grad():
        for b in (1, 2, 3, 1, 4, 5, 2):
            net.network(synthetic_frames(b, 256, b).cuda())
    assert len(eng._plans) <= eng.max_plans
    p1 = eng.plan_for(2, torch.device("cuda"))
    p2 = eng.plan_for(2, torch.device("cuda", torch.cuda.current_device()))
    assert p1 is p2
