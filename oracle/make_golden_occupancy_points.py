"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/occupancy_points.npz from the reference's own occupancy_grid_to_points
(SOccDPT/utils/__init__.py:532-568, compiled in memory: the module's matplotlib / wandb imports are missing here).

    python oracle/make_golden_occupancy_points.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_env  # noqa: E402
from test_oracle_occupancy_points import CASES, make_grid  # noqa: E402


def main():
    ref = ref_env.load_reference_function("SOccDPT/utils/__init__.py", "occupancy_grid_to_points", {"np": np})
    out = {}
    for name in CASES:
        g, G, scale = make_grid(name)
        pts = ref(g, grid_size=G, scale=scale)
        out[name + "_shape"] = np.array(pts.shape)
        out[name + "_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pts).tobytes()).digest(), np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "occupancy_points.npz"), **out)
    print({k: v.tolist() for k, v in out.items() if k.endswith("_shape")})


if __name__ == "__main__":
    main()
