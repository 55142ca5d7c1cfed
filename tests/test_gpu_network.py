"""End-to-end parity of the CUDA network path against the oracle / reference fixture on identical seeded
weights and inputs (BASELINE configs 1-2 shape: SOccDPT V3 dpt_swin2_tiny_256, image -> depth+seg+occupancy).

Tolerances (bf16 storage + bf16 tensor-core operands, fp32 accumulation and fp32 residual stream, vs the
fp32 reference; random N(0, 1/sqrt(fan_in)) weights give O(1) logits, i.e. no trained-network damping):
  inverse depth : |err| <= 2e-2 * max|depth| + 2e-2 * |depth| per element      
  segmentation  : after the sigmoid, mean |err| <= 8e-3 and max |err| <= 8e-2 over all B*3*256*256 values
                  (achieved on B200: mean ~4e-3, max ~4-5e-2 -- printed by the test; the max is a ~5 sigma
                  tail of bf16 rounding noise through ~45 layers, the CUDA-core fp32-accumulate reference
                  kernel shows the same figure as the tcgen05 kernel)
  encoder taps / path_1 : max |err| <= 3e-2 * max|x|
"""
import os

import numpy as np
import pytest
import torch

import golden_util as GU
import soccdpt_oracle as O
from soccdpt_b200 import SOccDPT_versions, load_model
from soccdpt_b200.synthetic import synthetic_frames, write_calib_yaml

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    yml = write_calib_yaml(str(tmp_path_factory.mktemp("calib") / "c.yaml"))
    sd = GU.tiny_state_dict(0)
    net = load_model(arch=SOccDPT_versions[3],
                     model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                                       camera_intrinsics_yaml=yml, model_type="dpt_swin2_tiny_256"),
                     device=torch.device("cuda"), model_path=None, model_type="dpt_swin2_tiny_256")
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net, sd


def _check(depth, seg, d_ref, s_ref):
    d, s = depth.float().cpu(), seg.float().cpu()
    derr = (d - d_ref).abs()
    dtol = 2e-2 * d_ref.abs().max() + 2e-2 * d_ref.abs()
    serr = (s - s_ref).abs().max().item()
    smean = (s - s_ref).abs().mean().item()
    print(f"depth max-abs err {derr.max().item():.3e} mean {derr.mean().item():.3e} (max|depth| {d_ref.abs().max().item():.3e}), "
          f"seg max-abs err {serr:.3e} mean {smean:.3e}")
    assert bool((derr <= dtol).all()), (derr.max().item(), d_ref.abs().max().item())
    assert serr <= 8e-2 and smean <= 8e-3, (serr, smean)


@pytest.mark.parametrize("impl", ["ref", "tcgen05"])
def test_network_matches_reference_fixture(setup, impl):
    net, sd = setup
    z = np.load(os.path.join(GU.GOLD, "net_tiny_b2.npz"))
    x = synthetic_frames(2, 256, 0).cuda()
    net.engine(impl)
    with torch.no_grad():
        depth, seg = net.network(x)
    torch.cuda.synchronize()
    _check(depth, seg, torch.from_numpy(z["depth"]), torch.from_numpy(z["seg"]))


def test_encoder_taps_and_path1_match_oracle(setup):
    net, sd = setup
    orc = O.OracleV3(sd)
    x = synthetic_frames(1, 256, 7)
    d_ref, s_ref, path_1, taps = orc.network(x)
    eng = net.engine("tcgen05")
    with torch.no_grad():
        depth, seg = net.network(x.cuda())
    torch.cuda.synchronize()
    plan = eng.plan_for(1, torch.device("cuda", torch.cuda.current_device()))
    for i, (t, Hs, Ws, C) in enumerate(plan["taps"]):
        got = t.float().cpu().view(1, Hs, Ws, C).permute(0, 3, 1, 2)
        err = (got - taps[i]).abs().max().item()
        print(f"tap{i + 1} max-abs err {err:.3e} (max {taps[i].abs().max().item():.3e})")
        assert err <= 3e-2 * taps[i].abs().max().item() + 3e-2
    p1 = plan["path_1"].float().cpu().permute(0, 3, 1, 2)
    assert (p1 - path_1).abs().max().item() <= 3e-2 * path_1.abs().max().item()
    _check(depth, seg, d_ref, s_ref)


def test_forward_tuple_and_occupancy_vs_oracle(setup):
    """net(x): shapes / squeeze quirks of the reference's 4-tuple, B=1 and B=2; the occupancy grid equals the
    bit-exact voxeliser applied to the returned maps, and is close to the fp32 oracle's grid."""
    net, sd = setup
    orc = O.OracleV3(sd)
    net.engine("tcgen05")
    for B in (1, 2):
        x = synthetic_frames(B, 256, 0)
        ref = orc(x)
        with torch.no_grad():
            out = net(x.cuda())
        torch.cuda.synchronize()
        for r, o in zip(ref, out):
            assert tuple(r.shape) == tuple(o.shape)
        assert set(torch.unique(out[3]).tolist()) <= {0.0, 1.0}
        for b in range(1, B):
            assert torch.equal(out[3][0], out[3][b])
        pts2, grid2 = net.voxelize(out[0].reshape(B, 1080, 1920).clone(), out[1].reshape(B, 3, 1080, 1920))
        assert torch.equal(grid2, out[3])
        a, b_ = out[3][0].cpu().bool(), ref[3][0].bool()
        inter, union = (a & b_).sum().item(), (a | b_).sum().item()
        print(f"B={B}: occupied cells ours {a.sum().item()} oracle {b_.sum().item()} IoU {inter / max(1, union):.4f}")
        assert union > 0 and inter / union > 0.9


def test_batch_invariance(setup):
    net, sd = setup
    net.engine("tcgen05")
    x = synthetic_frames(3, 256, 9).cuda()
    with torch.no_grad():
        d3, s3 = (t.clone() for t in net.network(x))
        d1, s1 = (t.clone() for t in net.network(x[1:2]))
    assert torch.allclose(d3[1:2], d1, rtol=1e-3, atol=1e-4) and torch.allclose(s3[1:2], s1, rtol=1e-3, atol=1e-4)


def test_bench_size_paths_match_small_batches(setup):
    """B = 8 reaches the tile counts of the bench (>= 2 x 148 tiles per GEMM: the resident-N-block path of the linear layers,
    full persistent grids) -- its per-frame results must equal the B = 2 runs of the same frames."""
    net, sd = setup
    net.engine("tcgen05")
    x = synthetic_frames(8, 256, 21).cuda()
    with torch.no_grad():
        d8, s8 = (t.clone() for t in net.network(x))
        for i in (0, 6):
            d2, s2 = net.network(x[i:i + 2])
            assert torch.allclose(d8[i:i + 2], d2, rtol=1e-3, atol=1e-4) and torch.allclose(s8[i:i + 2], s2, rtol=1e-3, atol=1e-4)


def test_cuda_graph_replay_is_bit_identical_to_eager(setup):
    """engine.enable_graphs(): the captured launch list replays to the same bytes as the eager launches (B = 1 and 3)."""
    net, _ = setup
    eng = net.engine("tcgen05")
    for B in (1, 3):
        x = synthetic_frames(B, 256, 5).cuda()
        with torch.no_grad():
            eng.enable_graphs(False)
            d0, s0 = [t.clone() for t in net.network(x)]
            eng.enable_graphs(True)
            for _ in range(3):                      # warm call, capture call, replay
                d1, s1 = net.network(x)
            assert eng.plan_for(B, x.device).get("graph") is not None
            assert torch.equal(d0, d1) and torch.equal(s0, s1)
            x2 = synthetic_frames(B, 256, 6).cuda()
            d2, s2 = [t.clone() for t in net.network(x2)]          # replay with new input
            eng.enable_graphs(False)
            d3, s3 = net.network(x2)
            assert torch.equal(d2, d3) and torch.equal(s2, s3)


def test_programmatic_dependent_launch_equals_plain_stream_order(setup):
    """PDL (csrc/common.cuh: kernels become resident and run their prologue under the previous kernel, then block in
    griddepcontrol.wait) must give the bits plain stream order gives -- alternating inputs, so that a kernel that read a
    buffer before its producer had finished (stale data of the previous call) shows up."""
    from soccdpt_b200 import _cabi
    net, sd = setup
    net.engine("tcgen05")
    lib = _cabi.load()
    xs = [synthetic_frames(B, 256, s).cuda() for s, B in ((0, 2), (1, 2), (2, 1), (3, 2))]

    def bits(t):
        return t.contiguous().view(torch.int32).clone()

    lib.soccdpt_set_pdl(0)
    try:
        with torch.no_grad():
            ref = [[bits(o) for o in net(x)] for x in xs]
            torch.cuda.synchronize()
            assert lib.soccdpt_set_pdl(15) == 0          # every kernel family
            for rep in range(4):
                for x, r in zip(xs, ref):
                    out = net(x)
                    for a, b in zip(out, r):
                        assert torch.equal(bits(a), b)
    finally:
        lib.soccdpt_set_pdl(-1)


def test_bench_batch_of_64_sampled_frames_match_the_oracle(setup):
    """the bench batch (B = 64) was only compared with the kernel's own small-batch results.  Here frames
    0 / 21 / 42 / 63 of a B = 64 forward are compared with the fp32 oracle DIRECTLY, at the tolerances of the B = 2 test, plus
      * segmentation LOGITS (before the x2 upsample and the sigmoid): |err| <= 0.25 + 3e-2 |logit|, mean |err| <= 8e-2 (the sigmoid compresses
        errors by >= 4x, so the post-activation bound alone would hide a logit-level problem);
      * occupancy: the grid computed from our maps and the grid computed from the oracle's maps agree on >= 90 % of the union
        of their set cells (a 0.5 m voxel flips when the depth error moves a point across a cell wall)."""
    net, sd = setup
    orc = O.OracleV3(sd)
    eng = net.engine("tcgen05")
    B = 64
    x = synthetic_frames(B, 256, 123)
    with torch.no_grad():
        depth, seg = (t.clone() for t in net.network(x.cuda()))
        plan = eng.plan_for(B, torch.device("cuda", torch.cuda.current_device()))
        logits = plan["seg_logits"].clone()
    torch.cuda.synchronize()
    pick = [0, 21, 42, 63]
    d_ref, s_ref, path_1, _ = orc.network(x[pick])
    _check(depth[pick], seg[pick], d_ref, s_ref)
    lg_ref = orc.seg_logits(path_1).permute(0, 2, 3, 1)                       # (4, h/2, w/2, C) like the plan's buffer
    lerr = (logits[pick].cpu() - lg_ref).abs()
    ltol = 0.25 + 3e-2 * lg_ref.abs()
    print(f"seg logits: max |err| {lerr.max().item():.3e} mean {lerr.mean().item():.3e} (max |logit| {lg_ref.abs().max().item():.2f}, "
          f"std {lg_ref.std().item():.2f})")
    assert bool((lerr <= ltol).all()) and lerr.mean().item() <= 8e-2, (lerr.max().item(), lerr.mean().item())
    # occupancy from our maps vs from the oracle's maps, frame by frame (per-frame grids: B = 1 calls)
    for k, b in enumerate(pick):
        ours = net.get_semantic_occupancy(depth[b:b + 1], seg[b:b + 1])[3][0].cpu().bool()
        ref = O.get_semantic_occupancy(d_ref[k:k + 1].clone(), s_ref[k:k + 1].clone(), orc.geom)[3][0].bool()
        inter, union = (ours & ref).sum().item(), (ours | ref).sum().item()
        print(f"frame {b}: occupied cells ours {ours.sum().item()} oracle {ref.sum().item()} IoU {inter / max(1, union):.4f}")
        assert union > 0 and inter / union >= 0.90
