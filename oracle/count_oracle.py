"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's ground-truth voxeliser
``OccupancyProcessor.transform_points_to_occupancy_grid_vect`` (SOccDPT/datasets/bdd_helper.py:289-362): points with
integer class ids are COUNTED per (voxel, class); a cell is reported as a point when its count >= point_count_threshold and
set in the boolean grid when its count > point_count_threshold (the reference's own asymmetry, :340 vs :357).

numpy promotion rules the arithmetic: ``cam_points / occupancy_shape(float32) * grid_size(tuple -> int64)``
  * float64 points (what process_frame produces, bdd_helper.py:455-489): everything in float64;
  * float32 points: an fp32 division, then the product in float64 (float32 * int64 -> float64).
``astype(int)`` truncates towards zero.  Negative class ids index from the end (numpy), ids >= num_classes raise.

Pinned against the reference class itself (compiled in memory from /root/reference, tests/test_oracle_count.py) and against
tests/golden/count_occupancy.npz (oracle/make_golden_count.py).  Only tests may import this file.
"""
import numpy as np


def occupancy_shape(grid_size, scale):
    return np.array([float(grid_size[i] / scale[i]) for i in range(len(grid_size))], dtype=np.float32)


def voxel_indices(cam_points, grid_size, occ_shape):
    """(n,3) integer voxel indices of every point and the mask of the points that survive both filters."""
    p = np.asarray(cam_points)
    finite = ~np.isinf(p).any(axis=1) & ~np.isnan(p).any(axis=1)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        ijk = (p / occ_shape * tuple(grid_size)).astype(int)
    G = grid_size
    inside = ((0 < ijk[:, 0]) & (ijk[:, 0] < G[0]) & (0 < ijk[:, 1]) & (ijk[:, 1] < G[1]) & (0 < ijk[:, 2]) & (ijk[:, 2] < G[2]))
    return ijk, finite & inside


def count_grid(cam_points, semantics, grid_size, scale, num_classes):
    """float32 count grid (G0,G1,G2,C), bdd_helper.py:293-337."""
    occ = occupancy_shape(grid_size, scale)
    ijk, keep = voxel_indices(cam_points, grid_size, occ)
    grid = np.zeros((grid_size[0], grid_size[1], grid_size[2], num_classes), dtype=np.float32)
    ijk, sem = ijk[keep], np.asarray(semantics)[keep]
    np.add.at(grid, (ijk[:, 0], ijk[:, 1], ijk[:, 2], sem), 1)
    return grid


def transform_points_to_occupancy_grid_vect(cam_points, semantics, grid_size, scale, num_classes, point_count_threshold):
    """-> {"occupancy_grid": bool (G0,G1,G2,C), "occupancy_points": float64 (n,4)}, bdd_helper.py:339-362."""
    occ = occupancy_shape(grid_size, scale)
    grid = count_grid(cam_points, semantics, grid_size, scale, num_classes)
    idx = np.argwhere(grid >= point_count_threshold)
    rows = []
    for c in range(num_classes):
        ci = idx[idx[:, 3] == c][:, :3]
        xyz = (ci / tuple(grid_size[:3]) * occ[:3]).astype(np.float32)
        rows.append(np.concatenate([xyz, np.full((xyz.shape[0], 1), c)], axis=1))
    return {"occupancy_grid": grid > point_count_threshold, "occupancy_points": np.concatenate(rows, axis=0), "counts": grid}


def argmax_labels(counts):
    """Per-voxel semantic label from the count grid: 0 = empty, 1 + argmax_c count (first maximum, as np.argmax).
    (No reference function: BASELINE.json's north_star names a per-voxel semantic argmax; this is its definition here.)"""
    c = np.asarray(counts)
    return np.where(c.sum(axis=-1) > 0, 1 + c.argmax(axis=-1), 0).astype(np.uint8)


def synthetic_points(n, grid_size, scale, num_classes, seed, dtype=np.float64, clustered=True):
    """Points spread over (and a little beyond) the grid's metric extent, clustered so that counts reach the threshold,
    with NaN / inf / boundary / negative coordinates mixed in."""
    rng = np.random.default_rng(seed)
    occ = occupancy_shape(grid_size, scale).astype(np.float64)
    if clustered:
        centres = rng.uniform(-0.05, 1.05, size=(max(n // 40, 1), 3)) * occ
        p = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 0.2, size=(n, 3))
    else:
        p = rng.uniform(-0.1, 1.1, size=(n, 3)) * occ
    k = max(n // 100, 1)
    p[rng.integers(0, n, k), rng.integers(0, 3, k)] = np.nan
    p[rng.integers(0, n, k), rng.integers(0, 3, k)] = np.inf
    p[rng.integers(0, n, k), rng.integers(0, 3, k)] = -np.inf
    # exact cell boundaries (index / G * occ) and the first / last planes
    b = rng.integers(0, n, k)
    cells = rng.integers(0, np.array(grid_size), size=(k, 3))
    p[b] = cells / np.array(grid_size, dtype=np.float64) * occ
    sem = rng.integers(0, num_classes, n).astype(np.int64)
    return p.astype(dtype), sem
