"""GPU parity of the post-processing kernels (BASELINE config 5):
bit-exact points / clamped inverse depth / occupancy grid against the oracle and against the fixtures
produced by the unmodified reference, through the public API -> C-ABI."""
import numpy as np
import pytest
import torch

import golden_util as GU
import soccdpt_oracle as O
from soccdpt_b200 import SOccDPT
from soccdpt_b200.synthetic import write_calib_yaml

pytestmark = pytest.mark.gpu


GU_SMALL = {"Camera.fx": 104.2, "Camera.fy": 104.6, "Camera.cx": 81.5, "Camera.cy": 46.8, "Camera.k1": 0.0, "Camera.k2": 0.0,
            "Camera.p1": 0.0, "Camera.p2": 0.0, "Camera.width": 160, "Camera.height": 90}


def _bits(t):
    return (t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)).view(np.uint32)


def _net(tmp_path, calib, geom, scale, **kw):
    yml = write_calib_yaml(str(tmp_path / "calib.yaml"), calib)
    return SOccDPT(camera_intrinsics_yaml=yml, compute_occ=True, grid_size=geom.grid_size, scale=scale, **kw)


@pytest.mark.parametrize("name", GU.VOXEL_CASES)
def test_voxelize_bit_exact_vs_reference_fixture_and_oracle(name, tmp_path):
    z, calib, geom, inv, seg = GU.load_voxel_case(name)
    net = _net(tmp_path, calib, geom, tuple(float(s) for s in z["scale"]))
    inv_d, seg_d = inv.cuda().clone(), seg.cuda()
    pts, grid = net.voxelize(inv_d, seg_d)
    torch.cuda.synchronize()
    # fixtures written by the reference itself
    assert GU.sha(inv_d.cpu().numpy()) == str(z["inv_sha"])
    assert GU.sha(pts.cpu().numpy()) == str(z["points_sha"])
    for b in range(grid.shape[0]):
        assert np.array_equal(GU.occupied_list(grid[b]), z["occupied"])
    # oracle, element by element
    inv_o, pts_o, grid_o = O.voxelize(inv.numpy(), seg.numpy(), geom)
    assert np.array_equal(_bits(inv_d), inv_o.view(np.uint32))
    assert np.array_equal(_bits(pts), pts_o.view(np.uint32))
    assert torch.equal(grid.cpu(), torch.from_numpy(grid_o))


@pytest.mark.parametrize("name", ["small_b2", "ragged_b3_grid64"])
def test_per_frame_mode_matches_oracle_and_union_is_or(name, tmp_path):
    z, calib, geom, inv, seg = GU.load_voxel_case(name)
    scale = tuple(float(s) for s in z["scale"])
    per = _net(tmp_path, calib, geom, scale, occupancy_mode="per_frame")
    uni = _net(tmp_path, calib, geom, scale)
    _, g_per = per.voxelize(inv.cuda().clone(), seg.cuda())
    _, g_uni = uni.voxelize(inv.cuda().clone(), seg.cuda())
    _, _, o_per = O.voxelize(inv.numpy(), seg.numpy(), geom, per_frame=True)
    assert torch.equal(g_per.cpu(), torch.from_numpy(o_per))
    assert torch.equal(g_uni[0], g_per.max(dim=0).values)


def test_edge_cases_empty_nan_inf_zero_classes(tmp_path):
    z, calib, geom, inv, seg = GU.load_voxel_case("small_b2")
    net = _net(tmp_path, calib, geom, tuple(float(s) for s in z["scale"]))
    B, H, W = inv.shape
    for fill in (float("nan"), float("inf"), 0.0, -3.0, 1e-30):
        i = torch.full((B, H, W), fill).cuda()
        pts, grid = net.voxelize(i, seg.cuda())
        io, po, go = O.voxelize(np.full((B, H, W), fill, np.float32), seg.numpy(), geom)
        assert np.array_equal(_bits(pts), po.view(np.uint32)) and np.array_equal(_bits(i), io.view(np.uint32))
        assert torch.equal(grid.cpu(), torch.from_numpy(go))
    # every class score exactly zero / negative zero -> nothing is set (torch.nonzero semantics); NaN counts
    for s_fill, expect_any in ((0.0, False), (-0.0, False), (float("nan"), True)):
        s = torch.full_like(seg, s_fill)
        pts, grid = net.voxelize(inv.cuda().clone(), s.cuda())
        _, _, go = O.voxelize(inv.numpy(), s.numpy(), geom)
        assert torch.equal(grid.cpu(), torch.from_numpy(go)) and bool(grid.any()) == expect_any
    # compute_occ=False -> no grid, points still produced
    net.compute_occ = False
    pts, grid = net.voxelize(inv.cuda().clone(), seg.cuda())
    assert grid is None and pts.shape == (B, H, W, 3)


def test_fused_resize_path_matches_oracle(tmp_path):
    """network-resolution maps -> get_semantic_occupancy: nearest classes exact, bicubic depth within fp32
    rounding of ATen's CPU kernel, squeeze quirks, and grid == voxelize(own maps) exactly."""
    geom = O.Geometry()
    net = _net(tmp_path, O.SYNTHETIC_CALIB, geom, (2.0, 2.0, 0.666))
    g = torch.Generator().manual_seed(3)
    for B in (1, 2):
        base = torch.rand(B, 1, 16, 16, generator=g) * 0.2 + 0.02
        inv = torch.nn.functional.interpolate(base, size=(256, 256), mode="bilinear")[:, 0].contiguous()
        seg = torch.sigmoid(torch.randn(B, 3, 256, 256, generator=g))
        ref = O.get_semantic_occupancy(inv.clone(), seg.clone(), geom)
        out = net.get_semantic_occupancy(inv.cuda(), seg.cuda())
        torch.cuda.synchronize()
        for r, o in zip(ref, out):
            assert tuple(r.shape) == tuple(o.shape)         # incl. the B=1 squeeze of the segmentation
        assert torch.equal(out[1].cpu(), ref[1])            # legacy nearest: exact
        assert torch.allclose(out[0].cpu(), ref[0], rtol=2e-6, atol=2e-7)
        assert torch.allclose(out[2].cpu(), ref[2], rtol=1e-5, atol=1e-5, equal_nan=True)
        # exact self-consistency: the fused grid equals the bit-exact voxeliser applied to the fused maps
        pts2, grid2 = net.voxelize(out[0].reshape(B, 1080, 1920).clone(), out[1].reshape(B, 3, 1080, 1920))
        assert torch.equal(grid2, out[3]) and torch.equal(pts2, out[2])
        # against the oracle only a handful of boundary-straddling voxels may differ
        diff = (out[3].cpu() != ref[3]).sum().item()
        assert diff <= 0.002 * max(1, int(ref[3].sum().item())), diff


@pytest.mark.parametrize("calib,scale", [(O.SYNTHETIC_CALIB, (2.0, 2.0, 0.666)), (GU_SMALL, (2.0, 2.0, 0.666)),
                                         (dict(O.SYNTHETIC_CALIB, **{"Camera.fx": 1277.0, "Camera.fy": 911.3}), (1.0, 3.0, 0.25))])
def test_fast_exact_arithmetic_is_ieee_on_all_inputs(tmp_path, calib, scale):
    """The straight-line fast paths of the exact stage (reciprocal by MUFU.RCP + Newton step; division by the calibration
    constants as an fp64 multiply) agree with correctly rounded IEEE 1/x and x/c on EVERY fp32 input they accept: exhaustive
    sweep of the 2^32 bit patterns on the device (soccdpt_selftest_exact_math)."""
    import ctypes
    from soccdpt_b200 import _cabi
    geom = O.Geometry(calib=calib, scale=scale)
    net = _net(tmp_path, calib, geom, scale)
    g = net._geometry()
    out = (ctypes.c_ulonglong * 6)()
    _cabi.check(_cabi.load().soccdpt_selftest_exact_math(ctypes.byref(g), ctypes.byref(out), _cabi.current_stream()), "selftest")
    assert list(out) == [0] * 6, list(out)


def test_full_size_batch_properties(tmp_path):
    """BASELINE config-5 size (1080x1920, 256x256x32) at B=8: equality with the oracle, idempotence,
    union == OR of per-frame grids, never-filled planes."""
    geom = O.Geometry()
    uni = _net(tmp_path, O.SYNTHETIC_CALIB, geom, (2.0, 2.0, 0.666))
    per = _net(tmp_path, O.SYNTHETIC_CALIB, geom, (2.0, 2.0, 0.666), occupancy_mode="per_frame")
    B = 8
    inv, seg = O.config5_maps(B, 1080, 1920, 3, seed=11)
    inv_d = inv.cuda().clone()
    pts, grid = uni.voxelize(inv_d, seg.cuda())
    inv_o, pts_o, grid_o = O.voxelize(inv.numpy(), seg.numpy(), geom)
    assert np.array_equal(_bits(pts), pts_o.view(np.uint32))
    assert np.array_equal(_bits(inv_d), inv_o.view(np.uint32))
    assert torch.equal(grid.cpu(), torch.from_numpy(grid_o))
    pts2, grid2 = uni.voxelize(inv_d, seg.cuda())           # clamp is idempotent
    assert torch.equal(pts2, pts) and torch.equal(grid2, grid)
    _, gper = per.voxelize(inv.cuda().clone(), seg.cuda())
    assert torch.equal(gper.max(dim=0).values, grid[0])
    assert grid[:, 0].sum() == 0 and grid[:, :, 0].sum() == 0 and grid[:, :, :, 0].sum() == 0
    assert set(torch.unique(grid).tolist()) <= {0.0, 1.0}


@pytest.mark.parametrize("hw", [(320, 1024), (192, 640), (384, 640)])
def test_external_depth_model_wrapper_like_eval_others(tmp_path, hw):
    """the reference's OtherModelWrapper (scripts/eval_others.py:54-247) subclasses SOccDPT around a
    third-party depth network, passes an all-zero segmentation at that network's resolution (monodepth2 / manydepth 320x1024
    and 192x640, PackNet 384x640) and returns get_semantic_occupancy(inv_depth, segmentation).  Same subclass here, a stub in
    place of the third-party network: shapes / squeeze quirks, points against the oracle, an empty grid (no class is non-zero)."""
    geom = O.Geometry()
    yml = write_calib_yaml(str(tmp_path / "calib.yaml"), O.SYNTHETIC_CALIB)

    class OtherModelWrapper(SOccDPT):
        def __init__(self, model, num_classes, **kwargs):
            super().__init__(**kwargs)
            self._model, self.num_classes = model, num_classes

        def forward(self, x):
            segmentation = torch.zeros((1, self.num_classes, x.shape[2], x.shape[3]), device=x.device)
            return self.get_semantic_occupancy(self._model(x), segmentation)

    h, w = hw
    g = torch.Generator().manual_seed(11)
    coarse = torch.rand(1, 1, 12, 20, generator=g) * 0.2 + 0.02
    inv = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear")[:, 0].contiguous()
    net = OtherModelWrapper(lambda x: inv.to(x.device), 3, camera_intrinsics_yaml=yml, compute_occ=True)
    out = net(torch.zeros(1, 3, h, w, device="cuda"))
    torch.cuda.synchronize()
    ref = O.get_semantic_occupancy(inv.clone(), torch.zeros(1, 3, h, w), geom)
    for r, o in zip(ref, out):
        assert tuple(r.shape) == tuple(o.shape)
    assert torch.allclose(out[0].cpu(), ref[0], rtol=2e-6, atol=2e-7)
    assert torch.allclose(out[2].cpu(), ref[2], rtol=1e-5, atol=1e-5, equal_nan=True)
    assert float(out[1].abs().max()) == 0.0 and float(out[3].sum()) == 0.0 and float(ref[3].sum()) == 0.0
