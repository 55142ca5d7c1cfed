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


def test_stream_of_fp32_batches_equals_direct_calls(net):
    B = 2
    batches = [synthetic_frames(B, 256, 20 + i).pin_memory() for i in range(4)]      # > 2: the buffers are recycled
    fs = FrameStream(net, B)
    got = [(r.index, r.inv_depth.clone(), r.segmentation.clone(), r.occupancy.clone()) for r in fs.run(batches)]
    assert [g[0] for g in got] == [0, 1, 2, 3]
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
