"""GPU parity of the sparse occupancy outputs (csrc/occupancy_points.cu): occupancy_grid_to_points
bit-equal to the reference fixture / the oracle, and the voxeliser's packed output equal to its own dense grid."""
import hashlib
import os

import numpy as np
import pytest
import torch

import golden_util as GU
import soccdpt_oracle as O
from soccdpt_b200 import SOccDPT
from soccdpt_b200.occupancy import mask_words, occupancy_grid_to_points, pack_grid, packed_to_points
from soccdpt_b200.synthetic import write_calib_yaml
from test_oracle_occupancy_points import CASES, GOLD, make_grid

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_grid_to_points_bit_equal(name):
    gold = np.load(GOLD)
    g, G, scale = make_grid(name)
    pts = occupancy_grid_to_points(torch.from_numpy(g).cuda(), G, scale).cpu().numpy()
    assert pts.dtype == np.float64 and pts.shape == tuple(gold[name + "_shape"])
    assert hashlib.sha256(np.ascontiguousarray(pts).tobytes()).digest() == gold[name + "_sha256"].tobytes()   # the reference's bytes
    assert np.array_equal(pts, O.occupancy_grid_to_points(g, G, scale))


def test_empty_and_full_grids():
    G = (16, 8, 4)
    z = torch.zeros((*G, 3), device="cuda")
    assert occupancy_grid_to_points(z, G, (1.0, 1.0, 1.0)).shape == (0, 4)
    f = torch.ones((*G, 3), device="cuda")
    pts = occupancy_grid_to_points(f, G, (1.0, 1.0, 1.0)).cpu().numpy()
    assert np.array_equal(pts, O.occupancy_grid_to_points(f.cpu().numpy(), G, (1.0, 1.0, 1.0)))


@pytest.mark.parametrize("name,mode", [("small_b2", "reference_union"), ("ragged_b3_grid64", "per_frame"), ("full_b1", "reference_union")])
def test_packed_voxeliser_output_equals_dense(name, mode, tmp_path):
    """occupancy_output='packed': the 4th output is the bit mask; expanding / listing it gives exactly the dense grid."""
    z, calib, geom, inv, seg = GU.load_voxel_case(name)
    scale = tuple(float(s) for s in z["scale"])
    yml = write_calib_yaml(str(tmp_path / "calib.yaml"), calib)
    kw = dict(camera_intrinsics_yaml=yml, compute_occ=True, grid_size=geom.grid_size, scale=scale, occupancy_mode=mode)
    dense = SOccDPT(**kw)
    packed = SOccDPT(occupancy_output="packed", **kw)
    p_d, g_d = dense.voxelize(inv.cuda().clone(), seg.cuda())
    p_p, m = packed.voxelize(inv.cuda().clone(), seg.cuda())
    assert torch.equal(p_d, p_p) and m.dtype == torch.int32
    B, n = inv.shape[0], mask_words(geom.grid_size)
    assert m.shape == ((B, n) if mode == "per_frame" else (n,))
    frames = range(B) if mode == "per_frame" else [0]
    for b in frames:
        mb = m[b] if mode == "per_frame" else m
        assert torch.equal(mb, pack_grid(g_d[b]))
        pts = packed_to_points(mb.contiguous(), geom.grid_size, scale, 3).cpu().numpy()
        assert np.array_equal(pts, O.occupancy_grid_to_points(g_d[b].cpu().numpy(), geom.grid_size, scale))
    if mode == "reference_union":        # the reference fixture's occupied list, through the packed path
        occ = np.argwhere(g_d[0].cpu().numpy() != 0).astype(np.int16)
        assert np.array_equal(occ, z["occupied"])
