"""Pins the oracle (oracle/soccdpt_oracle.py + oracle/voxel_oracle.c) against the fixtures the
UNMODIFIED reference produced (oracle/make_golden.py).  Runs on CPU, no reference tree needed."""
import numpy as np
import pytest
import torch

import golden_util as GU
import soccdpt_oracle as O
from soccdpt_b200.synthetic import synthetic_frames


@pytest.mark.parametrize("name", GU.VOXEL_CASES)
def test_voxel_oracle_matches_reference_fixture(name):
    z, calib, geom, inv, seg = GU.load_voxel_case(name)
    inv_c, pts, grid = O.voxelize(inv.numpy(), seg.numpy(), geom)
    assert GU.sha(inv_c) == str(z["inv_sha"])
    assert GU.sha(pts) == str(z["points_sha"])
    if "points" in z:
        assert np.array_equal(pts.view(np.uint32), z["points"].view(np.uint32))
    for b in range(grid.shape[0]):
        assert np.array_equal(GU.occupied_list(grid[b]), z["occupied"])
    assert set(np.unique(grid).tolist()) <= {0.0, 1.0}
    # planes i=0, j=0, k=0 are never written (strict 0 < ijk), SOccDPT.py:423-427
    assert grid[:, 0].sum() == 0 and grid[:, :, 0].sum() == 0 and grid[:, :, :, 0].sum() == 0


def test_voxel_oracle_thread_count_invariant():
    z, calib, geom, inv, seg = GU.load_voxel_case("small_b2")
    a = O.voxelize(inv.numpy(), seg.numpy(), geom, threads=1)
    b = O.voxelize(inv.numpy(), seg.numpy(), geom, threads=5)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_voxel_oracle_per_frame_union_property():
    """reference semantics: grid(B) == OR over frames of grid(frame)."""
    z, calib, geom, inv, seg = GU.load_voxel_case("ragged_b3_grid64")
    _, _, union = O.voxelize(inv.numpy(), seg.numpy(), geom)
    _, _, per = O.voxelize(inv.numpy(), seg.numpy(), geom, per_frame=True)
    assert np.array_equal(union[0], per.max(axis=0))
    for b in range(per.shape[0]):
        _, _, single = O.voxelize(inv[b:b + 1].numpy(), seg[b:b + 1].numpy(), geom)
        assert np.array_equal(single[0], per[b])


def test_network_oracle_matches_reference_fixture():
    z = np.load(GU.GOLD + "/net_tiny_b2.npz")
    sd = GU.tiny_state_dict(0)
    assert len(sd) == int(z["n_state_keys"])
    orc = O.OracleV3(sd)
    x = synthetic_frames(2, 256, 0)
    d, g, path_1, taps = orc.network(x)
    # same torch build -> bit-equal; other builds/CPUs -> fp32 reassociation noise only
    assert np.allclose(d.numpy(), z["depth"], rtol=1e-4, atol=1e-5)
    assert np.allclose(g.numpy(), z["seg"], rtol=1e-4, atol=1e-5)
    st = z["path1_mean_std_absmax"]
    assert np.allclose([path_1.mean(), path_1.std(), path_1.abs().max()], st, rtol=1e-4)
    out = orc(x)
    assert out[0].shape == (2, 1080, 1920) and out[3].shape == (2, 256, 256, 32, 3)
    if str(z["torch_version"]) == torch.__version__ and GU.sha(out[0].numpy()) == str(z["inv_up_sha"]):
        assert GU.sha(out[2].numpy()) == str(z["points_sha"])
        assert np.array_equal(GU.occupied_list(out[3][0]), z["occupied"])
