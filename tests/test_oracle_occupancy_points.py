"""Pins the oracle's occupancy_grid_to_points against the reference function itself (compiled in memory from
/root/reference when present) and against the committed fixture."""
import hashlib
import os

import numpy as np
import pytest

import ref_env
import soccdpt_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "occupancy_points.npz")
CASES = {"g256": ((256, 256, 32), (2.0, 2.0, 0.666), 3, 0.004, 0), "g64": ((64, 64, 8), (0.5, 0.5, 0.1665), 3, 0.05, 1),
         "ragged_c2": ((37, 21, 5), (0.3, 0.7, 0.11), 2, 0.2, 2), "c4_soft": ((16, 16, 9), (1.0, 1.0, 1.0), 4, 0.3, 3)}


def make_grid(name):
    G, scale, C, density, seed = CASES[name]
    rng = np.random.default_rng(seed)
    g = (rng.random((*G, C)) < density).astype(np.float32)
    if name == "c4_soft":            # non-binary values around the 0.5 threshold, NaN
        g = rng.random((*G, C)).astype(np.float32)
        g[0, 0, 0, 0] = 0.5
        g[0, 0, 1, 1] = np.nan
        g[0, 1, 0, 2] = np.float32(0.49999997)
    return g, G, scale


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_fixture(name):
    gold = np.load(GOLD)
    g, G, scale = make_grid(name)
    pts = O.occupancy_grid_to_points(g, G, scale)
    assert pts.dtype == np.float64 and pts.shape == tuple(gold[name + "_shape"])
    assert hashlib.sha256(np.ascontiguousarray(pts).tobytes()).digest() == gold[name + "_sha256"].tobytes()


@pytest.mark.skipif(not ref_env.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_live(name):
    ref = ref_env.load_reference_function("SOccDPT/utils/__init__.py", "occupancy_grid_to_points", {"np": np})
    g, G, scale = make_grid(name)
    a, b = ref(g, grid_size=G, scale=scale), O.occupancy_grid_to_points(g, G, scale)
    assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
