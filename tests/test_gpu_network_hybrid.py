"""SOccDPT V3 dpt_hybrid_384 (ResNetV2-50 + ViT-B/16 hybrid encoder, hooks [0, 1, 8, 11], readout
"project"; decoder levels 96/48/24/12, outputs 384x384) on identical seeded weights.

Two weight sets, because plain random init makes THIS network numerically chaotic under bf16 storage: the 16 GroupNorm
bottlenecks amplify every rounding flip, so that the CPU emulation of bf16 storage (oracle/storage_emulation.py, fp32
arithmetic) moves by 0.5% / 2% / 10% / 10% at the four taps when its input is perturbed by 1e-7, and sits 8% of max|d|
away from the fp32 oracle.  No kernel can be closer to the emulation than the emulation is to itself.

  (A) seed-0 random init (the weights of the fixture recorded from the repaired reference, tests/golden/net_hybrid_b1.npz):
      * vs the storage emulation: every statistic within 3x the emulation's own 1e-7-perturbation noise floor, computed
        in the test (self-calibrating: the CUDA path is as close to the emulation as the emulation is to itself);
      * vs the fp32 oracle / the reference's fixture (bf16-vs-fp32 gate): depth |err| <= 0.12 * max|d|, segmentation
        mean <= 4e-2, max <= 0.3 -- the emulation's own distance to fp32 (tests/test_storage_emulation.py prints it).
  (B) the same draw with the residual branches damped (norm3 gamma/beta x 0.1, like trained weights or timm's
      zero_init_last): well conditioned, so the Swin-base tolerances apply against the fp32 oracle:
      depth |err| <= 4e-2 * max|d| + 2e-2 * |d|, segmentation mean <= 8e-3, max <= 8e-2, taps mean |err| <= 3e-2 * mean|tap|."""
import os

import numpy as np
import pytest
import torch

import golden_util as GU
import soccdpt_oracle as O
import storage_emulation as E
from soccdpt_b200 import SOccDPT_versions, load_model
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml

pytestmark = pytest.mark.gpu
MT = "dpt_hybrid_384"


@pytest.fixture(scope="module")
def net(tmp_path_factory):
    yml = write_calib_yaml(str(tmp_path_factory.mktemp("calib") / "c.yaml"))
    n = load_model(arch=SOccDPT_versions[3],
                   model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                                     camera_intrinsics_yaml=yml, model_type=MT),
                   device=torch.device("cpu"), model_path=None, model_type=MT)
    sd = seeded_state_dict(n.state_dict(), 0)
    n.load_state_dict(sd, strict=True)
    n.sd = sd
    return n.to("cuda").eval()


def _run(net, x):
    with torch.no_grad():
        depth, seg = (t.clone().cpu() for t in net.network(x.cuda()))
    torch.cuda.synchronize()
    plan = net.engine().plan_for(x.shape[0], torch.device("cuda", torch.cuda.current_device()))
    taps = [t.float().reshape(x.shape[0], H, W, C).permute(0, 3, 1, 2).cpu() for (t, H, W, C) in plan["taps"]]
    return depth, seg, taps


def test_hybrid_random_init_within_emulation_noise_floor(net):
    x = synthetic_frames(1, 384, 0)
    depth, seg, taps = _run(net, x)
    with torch.no_grad():
        out = net(x.cuda())
    torch.cuda.synchronize()
    # (1) algorithmic parity: the storage-rounding emulation and its own noise floor
    d_emu, s_emu, _, t_emu = E.hybrid_network(net.sd, x)
    g = torch.Generator().manual_seed(1)
    d_pert, s_pert, _, t_pert = E.hybrid_network(net.sd, x * (1 + 1e-7 * torch.randn(x.shape, generator=g)))
    for i, (got, ref, pert) in enumerate(zip(taps, t_emu, t_pert)):
        err, floor = (got - ref).abs().mean().item(), (pert - ref).abs().mean().item()
        print(f"hybrid tap {i + 1}: mean |err| vs emulation {err:.3e}, emulation noise floor {floor:.3e} (mean|tap| {ref.abs().mean().item():.3e})")
        assert err <= 3 * floor
    for name, got, ref, pert in (("depth", depth, d_emu, d_pert), ("seg", seg, s_emu, s_pert)):
        e, f = (got - ref).abs(), (pert - ref).abs()
        print(f"hybrid {name}: vs emulation max {e.max().item():.3e} mean {e.mean().item():.3e}; "
              f"noise floor max {f.max().item():.3e} mean {f.mean().item():.3e}")
        assert e.mean().item() <= 3 * f.mean().item() and e.max().item() <= 3 * f.max().item()
    # (2) bf16 storage vs the fp32 oracle
    d_ref, s_ref, _, _ = O.OracleV3(net.sd, MT).network(x)
    derr, serr = (depth - d_ref).abs(), (seg - s_ref).abs()
    print(f"hybrid_384 vs fp32 oracle: depth max-abs err {derr.max().item():.3e} (max|d| {d_ref.abs().max().item():.3e}), "
          f"seg max {serr.max().item():.3e} mean {serr.mean().item():.3e}")
    assert derr.max().item() <= 0.12 * d_ref.abs().max().item()
    assert serr.max().item() <= 0.3 and serr.mean().item() <= 4e-2
    assert out[0].shape == (1, 1080, 1920) and out[3].shape == (1, 256, 256, 32, 3)
    # the fixture written by the reference (fp16 storage): the oracle must sit on it
    gold = np.load(os.path.join(GU.GOLD, "net_hybrid_b1.npz"))
    gd = torch.from_numpy(gold["depth"].astype(np.float32))
    assert (d_ref - gd).abs().max().item() <= 2e-3 * gd.abs().max().item()
    assert (depth - gd).abs().max().item() <= 0.12 * gd.abs().max().item()
    # the occupancy grid is bit-exactly the voxeliser applied to the returned maps ...
    pts2, grid2 = net.voxelize(out[0].reshape(1, 1080, 1920).clone(), out[1].reshape(1, 3, 1080, 1920))
    assert torch.equal(grid2, out[3])
    # ... and overlaps the reference's occupied-cell list as far as the depth deviation allows
    occ = set(map(tuple, GU.occupied_list(out[3][0].cpu()).tolist()))
    ref_occ = set(map(tuple, gold["occupied"].tolist()))
    inter = len(occ & ref_occ)
    print(f"hybrid_384 occupancy: mine {len(occ)} reference {len(ref_occ)} common {inter}")
    assert inter >= 0.5 * max(1, len(ref_occ))


def test_hybrid_damped_residuals_match_fp32_oracle(net):
    """(B): well-conditioned weights -> the fp32 oracle itself is the yardstick, with the Swin-base tolerances."""
    sd = seeded_state_dict(net.state_dict(), 0, residual_gain=0.1)
    net.load_state_dict(sd, strict=True)
    try:
        x = synthetic_frames(2, 384, 3)
        depth, seg, taps = _run(net, x)
        d_ref, s_ref, _, t_ref = O.OracleV3(sd, MT).network(x)
        for i, (got, ref) in enumerate(zip(taps, t_ref)):
            err = (got - ref).abs().mean().item()
            print(f"hybrid(damped) tap {i + 1}: mean |err| {err:.3e} (mean|tap| {ref.abs().mean().item():.3e})")
            assert err <= 3e-2 * ref.abs().mean().item()
        derr, serr = (depth - d_ref).abs(), (seg - s_ref).abs()
        print(f"hybrid(damped): depth max-abs err {derr.max().item():.3e} (max|d| {d_ref.abs().max().item():.3e}), "
              f"seg max {serr.max().item():.3e} mean {serr.mean().item():.3e}")
        assert bool((derr <= 4e-2 * d_ref.abs().max() + 2e-2 * d_ref.abs()).all())
        assert serr.max().item() <= 8e-2 and serr.mean().item() <= 8e-3
    finally:
        net.load_state_dict(net.sd, strict=True)


def test_hybrid_reference_fixture_weights_fp32_trunk_meets_swin_base_tolerance(net):
    """(C): the weights of the reference-recorded fixture (plain random init) with the ResNetV2 trunk in the fp32-storage parity
    mode (engine.set_trunk_precision("fp32"), csrc/trunk_fp32.cu): the bf16 rounding the 16 GroupNorm bottlenecks amplify is gone,
    and the CUDA path meets the Swin-base tolerances against the fp32 oracle AND the fixture the unmodified reference wrote
    (vit.py:147-258).  Taps 1 / 2 are the trunk's own outputs: rounded to bf16 once."""
    eng = net.engine()
    eng.set_trunk_precision("fp32")
    try:
        x = synthetic_frames(1, 384, 0)
        depth, seg, taps = _run(net, x)
        d_ref, s_ref, _, t_ref = O.OracleV3(net.sd, MT).network(x)
        for i, (got, ref) in enumerate(zip(taps, t_ref)):
            err = (got - ref).abs().mean().item()
            print(f"hybrid(fp32 trunk) tap {i + 1}: mean |err| {err:.3e} (mean|tap| {ref.abs().mean().item():.3e})")
            assert err <= (4e-3 if i < 2 else 3e-2) * ref.abs().mean().item()
        derr, serr = (depth - d_ref).abs(), (seg - s_ref).abs()
        print(f"hybrid(fp32 trunk) vs fp32 oracle: depth max-abs err {derr.max().item():.3e} (max|d| {d_ref.abs().max().item():.3e}), "
              f"seg max {serr.max().item():.3e} mean {serr.mean().item():.3e}")
        assert bool((derr <= 4e-2 * d_ref.abs().max() + 2e-2 * d_ref.abs()).all())
        assert serr.max().item() <= 8e-2 and serr.mean().item() <= 8e-3
        gold = np.load(os.path.join(GU.GOLD, "net_hybrid_b1.npz"))
        gd = torch.from_numpy(gold["depth"].astype(np.float32))
        assert (depth - gd).abs().max().item() <= 4e-2 * gd.abs().max().item()
        with torch.no_grad():
            out = net(x.cuda())
        occ = set(map(tuple, GU.occupied_list(out[3][0].cpu()).tolist()))
        ref_occ = set(map(tuple, gold["occupied"].tolist()))
        inter = len(occ & ref_occ)
        print(f"hybrid(fp32 trunk) occupancy: mine {len(occ)} reference {len(ref_occ)} common {inter}")
        assert inter >= 0.85 * max(1, len(ref_occ))
    finally:
        eng.set_trunk_precision("bf16")
