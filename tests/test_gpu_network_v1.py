"""SOccDPT_V1 on the GPU (reference SOccDPT.py:470-523): depth DPT + segmentation DPT (BatchNorm
folded into its residual conv units) on two streams, then the shared voxeliser -- against the fixture of the unmodified
reference and the oracle.  Tolerances as for SOccDPT_V3 (tests/test_gpu_network.py): inverse depth
|err| <= 2e-2 max|depth| + 2e-2 |depth|; segmentation (post-sigmoid) mean |err| <= 8e-3, max |err| <= 8e-2."""
import os

import numpy as np
import pytest
import torch

import golden_util as GU
import soccdpt_oracle as O
from soccdpt_b200 import DepthNet, SegNet, SOccDPT_versions, load_model
from soccdpt_b200.synthetic import synthetic_frames, write_calib_yaml

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    yml = write_calib_yaml(str(tmp_path_factory.mktemp("calib") / "c.yaml"))
    sd = GU.v1_tiny_state_dict(0)
    mt = "dpt_swin2_tiny_256"
    net = load_model(arch=SOccDPT_versions[1],
                     model_kwargs=dict(load_depth=False, load_seg=False, num_classes=3, compute_occ=True,
                                       camera_intrinsics_yaml=yml, model_type=mt),
                     device=torch.device("cuda"), model_path=None, model_type=mt)
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net, sd


@pytest.mark.parametrize("impl", ["ref", "tcgen05"])
def test_v1_network_matches_reference_fixture(setup, impl):
    net, sd = setup
    z = np.load(os.path.join(GU.GOLD, "net_v1_tiny_b2.npz"))
    x = synthetic_frames(2, 256, 0).cuda()
    with torch.no_grad():
        depth, seg = net.network(x, conv_impl=impl)
    torch.cuda.synchronize()
    d, s = depth.float().cpu(), seg.float().cpu()
    d_ref, s_ref = torch.from_numpy(z["depth"]), torch.from_numpy(z["seg"])
    derr = (d - d_ref).abs()
    serr, smean = (s - s_ref).abs().max().item(), (s - s_ref).abs().mean().item()
    print(f"V1 {impl}: depth max-abs err {derr.max().item():.3e} (max|depth| {d_ref.abs().max().item():.3e}), "
          f"seg max-abs err {serr:.3e} mean {smean:.3e}")
    assert bool((derr <= 2e-2 * d_ref.abs().max() + 2e-2 * d_ref.abs()).all())
    assert serr <= 8e-2 and smean <= 8e-3, (serr, smean)


def test_v1_forward_tuple_and_occupancy(setup):
    net, sd = setup
    orc = O.OracleV1(sd)
    net.engine("tcgen05"), net.seg_engine("tcgen05")
    for B in (1, 2):
        x = synthetic_frames(B, 256, 0)
        ref = orc(x)
        with torch.no_grad():
            out = net(x.cuda())
        torch.cuda.synchronize()
        for r, o in zip(ref, out):
            assert tuple(r.shape) == tuple(o.shape)
        # the grid is the bit-exact voxeliser applied to the maps the call returned ...
        inv_up, seg_up = out[0].reshape(B, 1080, 1920), out[1].reshape(B, 3, 1080, 1920)
        _, pts_o, grid_o = O.voxelize(inv_up.cpu().numpy().copy(), seg_up.cpu().numpy(), orc.geom)
        assert np.array_equal(out[2].cpu().numpy().view(np.uint32), pts_o.view(np.uint32))
        assert np.array_equal(out[3].cpu().numpy(), grid_o)
        # ... and close to the fp32 oracle's grid
        a, b_ = out[3][0].cpu().bool(), torch.as_tensor(ref[3][0]).bool()
        inter, union = (a & b_).sum().item(), (a | b_).sum().item()
        print(f"V1 B={B}: occupied cells ours {a.sum().item()} oracle {b_.sum().item()} IoU {inter / max(1, union):.4f}")
        assert union > 0 and inter / union > 0.9
    assert DepthNet(net)(x.cuda()).shape == out[0].shape and SegNet(net)(x.cuda()).shape == out[1].shape


def test_v1_two_streams_equal_one_stream(setup):
    """the segmentation plan on its side stream gives the bits a single-stream run gives."""
    net, sd = setup
    x = synthetic_frames(2, 256, 1).cuda()
    with torch.no_grad():
        d2, s2 = net.network(x)
        d2, s2 = d2.clone(), s2.clone()
        torch.cuda.synchronize()
        d1, _ = net.engine().run(x)
        _, s1 = net.seg_engine().run(x)
    torch.cuda.synchronize()
    assert torch.equal(d1, d2) and torch.equal(s1, s2)


def test_v1_frame_stream_equals_direct_calls(setup):
    """the host-frame path (three streams, double buffering) on top of SOccDPT_V1's own two-stream forward."""
    from soccdpt_b200.pipeline import FrameStream
    net, sd = setup
    B = 2
    batches = [synthetic_frames(B, 256, 30 + i).pin_memory() for i in range(4)]
    fs = FrameStream(net, B)
    got = [(r.index, r.inv_depth.clone(), r.segmentation.clone(), r.occupancy.clone()) for r in fs.run(batches)]
    assert [g[0] for g in got] == [0, 1, 2, 3]
    for (_, d, s, g), xb in zip(got, batches):
        with torch.no_grad():
            out = net(xb.cuda())
            d0, s0 = (t.clone().cpu() for t in net.network(xb.cuda()))
        assert torch.equal(d, d0) and torch.equal(s, s0) and torch.equal(g, out[3][0].cpu())
