"""Pins oracle/count_oracle.py (restatement of bdd_helper.py:289-362, the reference's counting ground-truth voxeliser) against
the reference class itself when /root/reference exists, and against tests/golden/count_occupancy.npz everywhere.  CPU only."""
import hashlib

import numpy as np
import pytest

import count_oracle as CO
import golden_util as GU
import make_golden_count as MG
import ref_env


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.mark.parametrize("name", list(MG.CASES))
def test_count_oracle_matches_golden(name):
    n, G, scale, C, thr, seed, dt = MG.CASES[name]
    z = np.load(GU.GOLD + "/count_occupancy.npz")
    pts, sem = CO.synthetic_points(n, G, scale, C, seed, np.dtype(dt))
    res = CO.transform_points_to_occupancy_grid_vect(pts, sem, G, scale, C, thr)
    assert res["occupancy_grid"].dtype == np.bool_ and res["occupancy_points"].dtype == np.float64
    assert tuple(z[name + "_points_shape"]) == res["occupancy_points"].shape
    assert int(z[name + "_grid_set"]) == int(res["occupancy_grid"].sum()) > 0
    assert np.array_equal(_sha(res["occupancy_grid"]), z[name + "_grid_sha"])
    assert np.array_equal(_sha(res["occupancy_points"]), z[name + "_points_sha"])


@pytest.mark.skipif(not ref_env.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("name", ["f64_small", "f32_odd_grid"])
def test_count_oracle_matches_reference_live(name):
    n, G, scale, C, thr, seed, dt = MG.CASES[name]
    pts, sem = CO.synthetic_points(n, G, scale, C, seed, np.dtype(dt))
    ref = MG.reference_processor(G, scale, C, thr).transform_points_to_occupancy_grid_vect(pts, sem)
    res = CO.transform_points_to_occupancy_grid_vect(pts, sem, G, scale, C, thr)
    assert np.array_equal(ref["occupancy_grid"], res["occupancy_grid"])
    assert np.array_equal(ref["occupancy_points"].view(np.uint64), res["occupancy_points"].view(np.uint64))


def test_argmax_labels_definition():
    counts = np.zeros((2, 2, 1, 3), np.float32)
    counts[0, 0, 0] = (0, 5, 5)      # tie -> first maximum
    counts[1, 1, 0] = (2, 0, 1)
    lab = CO.argmax_labels(counts)
    assert lab.tolist() == [[[2], [0]], [[0], [1]]]
