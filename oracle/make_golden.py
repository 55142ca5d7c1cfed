"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference, imported through oracle/ref_env.py with the timm shim) on CPU.
The reference ships no tests / golden vectors of its own, so these
fixtures ARE the pin for the oracle and for the CUDA path.

    python oracle/make_golden.py            # rewrites tests/golden/

Fixtures (all inputs are regenerated from seeds by the tests; only outputs are stored):
  voxel_<case>.npz   reference SOccDPT.get_semantic_occupancy on config-5 style maps
                     (SURVEY.md 8d): points bytes (small cases) or sha256 (full size),
                     clamped inv_depth sha256, sorted occupied (i,j,k,c) list, B
  net_tiny_b2.npz    reference SOccDPT_V3 dpt_swin2_tiny_256 forward, seeded weights,
                     x = randn(2,3,256,256; seed 0): depth, seg, tap/path_1 statistics,
                     occupied-cell list of the end-to-end grid
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_env  # noqa: E402
import soccdpt_oracle as O  # noqa: E402
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name: (calib, grid_size, scale, B, seed, scaled_tanh)
FULL = dict(O.SYNTHETIC_CALIB)
SMALL = {"Camera.fx": 104.2, "Camera.fy": 104.6, "Camera.cx": 81.5, "Camera.cy": 46.8,
         "Camera.k1": 0.0, "Camera.k2": 0.0, "Camera.p1": 0.0, "Camera.p2": 0.0,
         "Camera.width": 160, "Camera.height": 90}
RAGGED = dict(SMALL, **{"Camera.width": 157, "Camera.height": 83, "Camera.cx": 77.25, "Camera.cy": 40.5})
VOXEL_CASES = {
    "small_b2": (SMALL, (256, 256, 32), (2.0, 2.0, 0.666), 2, 0, False),
    "small_b1_tanh": (SMALL, (256, 256, 32), (2.0, 2.0, 0.666), 1, 1, True),
    "ragged_b3_grid64": (RAGGED, (64, 64, 8), (0.5, 0.5, 0.1665), 3, 2, False),
    "full_b1": (FULL, (256, 256, 32), (2.0, 2.0, 0.666), 1, 0, False),
    "full_b2_grid128": (FULL, (128, 128, 16), (1.0, 1.0, 0.333), 2, 1, True),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def occupied(grid):
    """(n,4) int16 sorted list of (i,j,k,c) set in frame 0, after checking all frames agree."""
    g = grid.numpy() if isinstance(grid, torch.Tensor) else grid
    for b in range(1, g.shape[0]):
        assert np.array_equal(g[0], g[b]), "reference grid differs across the batch"
    assert set(np.unique(g).tolist()) <= {0.0, 1.0}
    return np.argwhere(g[0] != 0).astype(np.int16)


def make_voxel_cases(ref_model):
    for name, (calib, grid, scale, B, seed, tanh) in VOXEL_CASES.items():
        yml = write_calib_yaml(f"/tmp/soccdpt_golden_{name}.yaml", calib)
        net = ref_model.SOccDPT(camera_intrinsics_yaml=yml, compute_occ=True, grid_size=grid, scale=scale)
        H, W = int(calib["Camera.height"]), int(calib["Camera.width"])
        inv, seg = O.config5_maps(B, H, W, 3, seed, tanh)
        # feed maps already at camera resolution: both interpolations are then identities
        with torch.no_grad():
            inv_o, seg_o, pts, g = net.get_semantic_occupancy(inv.clone(), seg.clone())
        inv_o = inv_o.reshape(B, H, W).numpy()
        pts = pts.numpy()
        assert np.array_equal(seg_o.reshape(B, 3, H, W).numpy(), seg.numpy())
        occ = occupied(g)
        out = dict(B=B, H=H, W=W, seed=seed, scaled_tanh=int(tanh), grid_size=np.array(grid),
                   scale=np.array(scale, np.float64), calib_keys=np.array(list(calib.keys())),
                   calib_vals=np.array(list(calib.values()), np.float64),
                   occupied=occ, inv_sha=sha(inv_o), points_sha=sha(pts),
                   occupancy_shape=np.array(net.occupancy_shape, np.float32))
        if H * W <= 20000:
            out["points"] = pts
            out["inv_depth"] = inv_o
        np.savez_compressed(os.path.join(GOLD, f"voxel_{name}.npz"), **out)
        print(f"voxel_{name}: B={B} {H}x{W} grid={grid} occupied cells={len(occ)}")


def make_net_case(ref_loader, ref_model):
    yml = write_calib_yaml("/tmp/soccdpt_golden_full.yaml", FULL)
    mt = "dpt_swin2_tiny_256"
    net = ref_loader.load_model(
        arch=ref_model.SOccDPT_versions[3],
        model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                          camera_intrinsics_yaml=yml, model_type=mt),
        device=torch.device("cpu"), model_path=None, model_type=mt).eval()
    sd = seeded_state_dict(net.state_dict(), 0)
    net.load_state_dict(sd, strict=True)
    x = synthetic_frames(2, 256, 0)
    with torch.no_grad():
        depth, path_1 = net.depth_net.forward(x)
        seg = net.seg_head(path_1)
        taps = [net.pretrained.activations[str(i)] for i in (1, 2, 3, 4)]  # (B,L,C) hook outputs
        inv_up, seg_up, pts, grid = net(x)
    stats = {}
    for i, t in enumerate(taps):
        stats[f"tap{i + 1}_mean_std_absmax"] = np.array([t.mean(), t.std(), t.abs().max()], np.float64)
        stats[f"tap{i + 1}_sample"] = t[:, :: max(1, t.shape[1] // 16), :: max(1, t.shape[2] // 16)].numpy()
    stats["path1_mean_std_absmax"] = np.array([path_1.mean(), path_1.std(), path_1.abs().max()], np.float64)
    stats["path1_sample"] = path_1[:, ::16, ::8, ::8].numpy()
    np.savez_compressed(
        os.path.join(GOLD, "net_tiny_b2.npz"), depth=depth.numpy(), seg=seg.numpy().astype(np.float32),
        occupied=occupied(grid), inv_up_sha=sha(inv_up.numpy()), points_sha=sha(pts.numpy()),
        n_state_keys=len(sd), state_keys_sha=hashlib.sha256("\n".join(sorted(sd)).encode()).hexdigest(),
        torch_version=torch.__version__, **stats)
    print("net_tiny_b2: depth", tuple(depth.shape), float(depth.min()), float(depth.max()),
          "occupied", int(grid[0].sum()))
    with open(os.path.join(GOLD, "state_keys_tiny.txt"), "w") as f:
        for k in net.state_dict().keys():
            f.write(f"{k} {tuple(net.state_dict()[k].shape)}\n")


def make_hybrid_case(ref_loader, ref_model):
    """dpt_hybrid_384 through the reference with the in-memory `value` repair (ref_env.py)."""
    yml = write_calib_yaml("/tmp/soccdpt_golden_full.yaml", FULL)
    mt = "dpt_hybrid_384"
    net = ref_loader.load_model(
        arch=ref_model.SOccDPT_versions[3],
        model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                          camera_intrinsics_yaml=yml, model_type=mt),
        device=torch.device("cpu"), model_path=None, model_type=mt).eval()
    sd = seeded_state_dict(net.state_dict(), 0)
    net.load_state_dict(sd, strict=True)
    x = synthetic_frames(1, 384, 0)
    with torch.no_grad():
        depth, path_1 = net.depth_net.forward(x)
        seg = net.seg_head(path_1)
        inv_up, seg_up, pts, grid = net(x)
    np.savez_compressed(
        os.path.join(GOLD, "net_hybrid_b1.npz"), depth=depth.numpy().astype(np.float16), seg=seg.numpy().astype(np.float16),
        depth_sha=sha(depth.numpy()), seg_sha=sha(seg.numpy()), occupied=occupied(grid),
        path1_mean_std_absmax=np.array([path_1.mean(), path_1.std(), path_1.abs().max()], np.float64),
        n_state_keys=len(sd), torch_version=torch.__version__)
    print("net_hybrid_b1: depth", tuple(depth.shape), float(depth.min()), float(depth.max()), "occupied", int(grid[0].sum()))
    with open(os.path.join(GOLD, "state_keys_hybrid.txt"), "w") as f:
        for k in net.state_dict().keys():
            f.write(f"{k} {tuple(net.state_dict()[k].shape)}\n")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    ref_loader, ref_model = ref_env.import_reference()
    torch.manual_seed(0)
    make_voxel_cases(ref_model)
    make_net_case(ref_loader, ref_model)
    make_hybrid_case(ref_loader, ref_model)
