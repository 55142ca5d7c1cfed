"""BASELINE config 5: stand-alone depth -> occupancy voxeliser sweep (grid resolution x batch) on camera-resolution maps
(1080 x 1920, 3 classes, the synthetic maps of SURVEY.md 8d: 1 % NaN / inf / zero / negative inverse depths), CUDA events,
against the HBM roofline (algorithmic bytes: maps in + clamped inverse depth, points and dense grid out) and, for the
smallest batch of every grid, the C oracle on the host cores (bit-equal outputs are asserted while timing).

    PYTHONPATH=. python tools/bench_voxeliser.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import soccdpt_oracle as O  # noqa: E402  (checker + CPU baseline only)
from soccdpt_b200 import SOccDPT  # noqa: E402
from soccdpt_b200.synthetic import write_calib_yaml  # noqa: E402

GRIDS = [((64, 64, 8), (0.5, 0.5, 0.1665)), ((128, 128, 16), (1.0, 1.0, 0.333)), ((256, 256, 32), (2.0, 2.0, 0.666))]
BATCHES = [1, 8, 64]


def main():
    peak = 6539.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    yml = write_calib_yaml("/tmp/bench_voxeliser_calib.yaml")
    H, W, C = 1080, 1920, 3
    print(f"{'grid':>14s} {'B':>3s} {'ms':>8s} {'GB/s':>8s} {'of HBM':>7s} {'frames/s':>10s} {'CPU oracle frames/s':>20s}")
    for grid, scale in GRIDS:
        net = SOccDPT(camera_intrinsics_yaml=yml, compute_occ=True, grid_size=grid, scale=scale)
        geom = O.Geometry(grid_size=grid, scale=scale)
        for B in BATCHES:
            inv, seg = O.config5_maps(B, H, W, C, seed=B)
            inv_d, seg_d = inv.cuda(), seg.cuda()
            for _ in range(3):
                pts, g = net.voxelize(inv_d.clone(), seg_d)
            work = inv_d.clone()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            s.record()
            for _ in range(reps):
                pts, g = net.voxelize(work, seg_d)          # clamp is idempotent: same work every repetition
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / reps
            cells = grid[0] * grid[1] * grid[2] * C
            nbytes = B * (4 * H * W * (1 + C) + 4 * H * W + 12 * H * W + 4 * cells)
            cpu = ""
            if B == BATCHES[0]:
                t0 = time.time()
                inv_o, pts_o, grid_o = O.voxelize(inv.numpy(), seg.numpy(), geom)
                dt = time.time() - t0
                assert np.array_equal(pts.cpu().numpy().view(np.uint32), pts_o.view(np.uint32))
                assert np.array_equal(g.cpu().numpy(), grid_o)
                cpu = f"{B / dt:10.1f} ({os.cpu_count()} threads)"
            print(f"{str(grid):>14s} {B:3d} {ms:8.3f} {nbytes / ms / 1e6:8.0f} {nbytes / ms / 1e6 / peak:7.2f} {B / ms * 1e3:10.0f} {cpu:>20s}")


if __name__ == "__main__":
    main()
