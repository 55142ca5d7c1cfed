"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/count_occupancy.npz from the reference's own
OccupancyProcessor.transform_points_to_occupancy_grid_vect (SOccDPT/datasets/bdd_helper.py:238-362, class compiled in memory:
the module imports cv2 / pandas / PIL at the top, none of which the method needs).

    python oracle/make_golden_count.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import count_oracle as CO  # noqa: E402
import ref_env  # noqa: E402

# name: (n points, grid, scale, classes, threshold, seed, dtype)
CASES = {
    "f64_small": (20000, (32, 32, 8), (2.0, 2.0, 0.666), 3, 10, 1, "float64"),
    "f32_small": (20000, (32, 32, 8), (2.0, 2.0, 0.666), 3, 10, 2, "float32"),
    "f64_odd_grid": (50000, (50, 30, 12), (1.7, 2.3, 0.9), 4, 3, 3, "float64"),
    "f32_odd_grid": (50000, (50, 30, 12), (1.7, 2.3, 0.9), 4, 3, 4, "float32"),
    "f64_full": (400000, (256, 256, 32), (2.0, 2.0, 0.666), 3, 10, 5, "float64"),
}


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def reference_processor(grid_size, scale, num_classes, threshold):
    cls = ref_env.load_reference_class("SOccDPT/datasets/bdd_helper.py", "OccupancyProcessor", {"np": np})
    K = np.array([[1250.6, 0.0, 978.4], [0.0, 1254.8, 562.1], [0.0, 0.0, 1.0]])
    return cls(intrinsic_matrix=K, height=1080, width=1920, grid_size=grid_size, scale=scale, shift=(0.0, 0.0, 0.0),
               pc_scale=(10000.0, 50000.0, 800.0), pc_shift=(55.0, -20.0, 15.0), point_count_threshold=threshold,
               class_2_color={}, color_2_class={}, num_classes=num_classes)


def main():
    out = {}
    for name, (n, G, scale, C, thr, seed, dt) in CASES.items():
        pts, sem = CO.synthetic_points(n, G, scale, C, seed, np.dtype(dt))
        res = reference_processor(G, scale, C, thr).transform_points_to_occupancy_grid_vect(pts, sem)
        out[name + "_grid_sha"] = sha(res["occupancy_grid"])
        out[name + "_points_sha"] = sha(res["occupancy_points"])
        out[name + "_points_shape"] = np.array(res["occupancy_points"].shape)
        out[name + "_grid_set"] = np.array(int(res["occupancy_grid"].sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "count_occupancy.npz"), **out)
    print({k: v.tolist() for k, v in out.items() if k.endswith(("_shape", "_set"))})


if __name__ == "__main__":
    main()
