"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/net_v1_tiny_b2.npz + state_keys_v1_tiny.txt from the UNMODIFIED
reference (through oracle/timm_shim, in the build container): SOccDPT_V1 (SOccDPT.py:470-523), dpt_swin2_tiny_256, seeded
weights and frames.  Run:  python oracle/make_golden_v1.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_env  # noqa: E402
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_reference_v1(ref_loader, ref_model, yml, mt="dpt_swin2_tiny_256"):
    return ref_loader.load_model(
        arch=ref_model.SOccDPT_versions[1],
        model_kwargs=dict(load_depth=False, load_seg=False, num_classes=3, compute_occ=True,
                          camera_intrinsics_yaml=yml, model_type=mt),
        device=torch.device("cpu"), model_path=None, model_type=mt).eval()


if __name__ == "__main__":
    ref_loader, ref_model = ref_env.import_reference()
    yml = write_calib_yaml("/tmp/soccdpt_golden_full.yaml")
    net = build_reference_v1(ref_loader, ref_model, yml)
    sd = seeded_state_dict(net.state_dict(), 0)
    net.load_state_dict(sd, strict=True)
    x = synthetic_frames(2, 256, 0)
    with torch.no_grad():
        depth = net.depth_net(x)
        seg = net.seg_net(x)
        inv_up, seg_up, pts, grid = net(x)
    occupied = np.argwhere(grid[0].numpy() != 0).astype(np.int16)
    np.savez_compressed(
        os.path.join(GOLD, "net_v1_tiny_b2.npz"), depth=depth.numpy(), seg=seg.numpy().astype(np.float32),
        occupied=occupied, inv_up_sha=sha(inv_up.numpy()), points_sha=sha(pts.numpy()), n_state_keys=len(sd),
        torch_version=torch.__version__)
    print("net_v1_tiny_b2: depth", tuple(depth.shape), float(depth.min()), float(depth.max()), "seg", tuple(seg.shape),
          float(seg.min()), float(seg.max()), "occupied", len(occupied))
    with open(os.path.join(GOLD, "state_keys_v1_tiny.txt"), "w") as f:
        for k, v in net.state_dict().items():
            f.write(f"{k} {tuple(v.shape)}\n")
