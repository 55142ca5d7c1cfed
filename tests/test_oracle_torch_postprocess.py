"""Pins oracle/torch_postprocess.py (the op-for-op ATen restatement of SOccDPT.py:264-463) against the fixtures the
UNMODIFIED reference wrote and against the C oracle.  CPU only."""
import numpy as np
import pytest
import torch

import golden_util as GU
import soccdpt_oracle as O
import torch_postprocess as TP


@pytest.mark.parametrize("name", ["small_b2", "small_b1_tanh", "ragged_b3_grid64"])
def test_torch_port_matches_reference_fixture_and_c_oracle(name):
    z, calib, geom, inv, seg = GU.load_voxel_case(name)
    inv_t, pts_t, grid_t = TP.voxelize(inv, seg, geom, device="cpu")
    assert GU.sha(inv_t.numpy()) == str(z["inv_sha"])
    assert GU.sha(pts_t.numpy()) == str(z["points_sha"])
    for b in range(grid_t.shape[0]):
        assert np.array_equal(GU.occupied_list(grid_t[b]), z["occupied"])
    inv_c, pts_c, grid_c = O.voxelize(inv.numpy(), seg.numpy(), geom)
    assert np.array_equal(pts_t.numpy().view(np.uint32), pts_c.view(np.uint32))
    # the reference's `+= 1` with repeated indices stores 1 (index_put_ without accumulate): same 0/1 grid
    assert torch.equal(grid_t, torch.from_numpy(grid_c))


def test_torch_port_full_tuple_matches_oracle_tuple():
    geom = O.Geometry()
    g = torch.Generator().manual_seed(3)
    inv = torch.rand(1, 64, 64, generator=g) * 0.2 + 0.02
    seg = torch.sigmoid(torch.randn(1, 3, 64, 64, generator=g))
    a = TP.get_semantic_occupancy(inv.clone(), seg.clone(), geom)
    b = O.get_semantic_occupancy(inv.clone(), seg.clone(), geom)
    for x, y in zip(a, b):
        assert x.shape == y.shape and torch.equal(x, y)
