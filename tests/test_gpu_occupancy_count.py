"""GPU parity of the counting voxeliser (csrc/occupancy_count.cu) with the reference's ground-truth generator
OccupancyProcessor.transform_points_to_occupancy_grid_vect (bdd_helper.py:289-362): bit-exact against the fixtures the
reference wrote (tests/golden/count_occupancy.npz) and against the numpy oracle, element by element."""
import hashlib

import numpy as np
import pytest
import torch

import count_oracle as CO
import golden_util as GU
import make_golden_count as MG
from soccdpt_b200.occupancy import OccupancyCounter

pytestmark = pytest.mark.gpu


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.mark.parametrize("name", list(MG.CASES))
def test_count_mode_bit_exact_vs_reference_fixture_and_oracle(name):
    n, G, scale, C, thr, seed, dt = MG.CASES[name]
    z = np.load(GU.GOLD + "/count_occupancy.npz")
    pts, sem = CO.synthetic_points(n, G, scale, C, seed, np.dtype(dt))
    oc = OccupancyCounter(grid_size=G, scale=scale, num_classes=C, point_count_threshold=thr)
    out = oc.transform_points_to_occupancy_grid_vect(torch.from_numpy(pts).cuda(), torch.from_numpy(sem).cuda())
    grid = out["occupancy_grid"].cpu().numpy()
    points = out["occupancy_points"].cpu().numpy()
    assert grid.dtype == np.bool_ and points.dtype == np.float64
    assert np.array_equal(_sha(grid), z[name + "_grid_sha"])
    assert np.array_equal(_sha(points), z[name + "_points_sha"])
    ref = CO.transform_points_to_occupancy_grid_vect(pts, sem, G, scale, C, thr)
    assert np.array_equal(out["counts"].cpu().numpy(), ref["counts"].astype(np.int32))
    assert np.array_equal(out["labels"].cpu().numpy(), CO.argmax_labels(ref["counts"]))


def test_int32_semantics_negative_ids_and_accumulation():
    G, scale, C, thr = (16, 16, 8), (2.0, 2.0, 0.666), 3, 2
    pts, sem = CO.synthetic_points(5000, G, scale, C, 9, np.float32)
    sem_neg = sem.copy()
    sem_neg[::3] -= C                                    # numpy: -1 is the last class
    oc = OccupancyCounter(grid_size=G, scale=scale, num_classes=C, point_count_threshold=thr)
    p = torch.from_numpy(pts).cuda()
    c1 = oc.count(p, torch.from_numpy(sem_neg.astype(np.int32)).cuda())
    ref = CO.count_grid(pts, sem_neg, G, scale, C)
    assert np.array_equal(c1.cpu().numpy(), ref.astype(np.int32))
    # two calls accumulate: the second half on top of the first
    h = len(pts) // 2
    c2 = oc.count(p[:h], torch.from_numpy(sem[:h]).cuda())
    c2 = oc.count(p[h:], torch.from_numpy(sem[h:]).cuda(), counts=c2)
    assert np.array_equal(c2.cpu().numpy(), CO.count_grid(pts, sem, G, scale, C).astype(np.int32))


def test_empty_input_all_nonfinite_and_out_of_range_class():
    G, scale, C = (8, 8, 4), (2.0, 2.0, 0.666), 3
    oc = OccupancyCounter(grid_size=G, scale=scale, num_classes=C, point_count_threshold=1)
    out = oc.transform_points_to_occupancy_grid_vect(torch.zeros((0, 3), dtype=torch.float64).cuda(), torch.zeros(0, dtype=torch.int64).cuda())
    assert out["occupancy_points"].shape == (0, 4) and not out["occupancy_grid"].any() and not out["labels"].any()
    bad = torch.full((64, 3), float("nan"), dtype=torch.float64).cuda()
    out = oc.transform_points_to_occupancy_grid_vect(bad, torch.zeros(64, dtype=torch.int64).cuda())
    assert out["counts"].sum().item() == 0
    ok = torch.tensor([[1.0, 1.0, 3.0]], dtype=torch.float64).repeat(40, 1).cuda()          # all in cell (2, 2, 1)
    with pytest.raises(IndexError):
        oc.transform_points_to_occupancy_grid_vect(ok, torch.full((40,), C, dtype=torch.int64).cuda())
    out = oc.transform_points_to_occupancy_grid_vect(ok, torch.full((40,), 1, dtype=torch.int64).cuda())
    assert out["counts"].max().item() == 40 and out["occupancy_points"].shape == (1, 4) and int(out["labels"].max()) == 2
