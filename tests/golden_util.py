"""Helpers shared by the parity tests: load tests/golden fixtures, rebuild their seeded inputs."""
import ast
import hashlib
import os

import numpy as np
import torch

import soccdpt_oracle as O
from soccdpt_b200.synthetic import seeded_state_dict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VOXEL_CASES = ["small_b2", "small_b1_tanh", "ragged_b3_grid64", "full_b1", "full_b2_grid128"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_voxel_case(name):
    z = np.load(os.path.join(GOLD, f"voxel_{name}.npz"), allow_pickle=False)
    calib = {str(k): float(v) for k, v in zip(z["calib_keys"], z["calib_vals"])}
    geom = O.Geometry(calib, 3, tuple(int(g) for g in z["grid_size"]), tuple(float(s) for s in z["scale"]))
    inv, seg = O.config5_maps(int(z["B"]), int(z["H"]), int(z["W"]), 3, int(z["seed"]), bool(z["scaled_tanh"]))
    # the fixture went through the reference's resize (SOccDPT.py:270-282) at identical size: nearest is
    # an identity, bicubic is an identity except that it smears NaN/inf into the 4x4 neighbourhood
    # (0 * inf).  The voxeliser's input is the map AFTER that resize.
    inv = torch.nn.functional.interpolate(inv.unsqueeze(1), size=(geom.height, geom.width), mode="bicubic",
                                          align_corners=False).squeeze(1).contiguous()
    return z, calib, geom, inv, seg


def occupied_list(grid):
    g = grid.cpu().numpy() if isinstance(grid, torch.Tensor) else np.asarray(grid)
    return np.argwhere(g != 0).astype(np.int16)


def tiny_state_dict(seed=0, keys_file="state_keys_tiny.txt", encoder_prefixes=("depth_net.pretrained.model.", "pretrained.model.")):
    """The seeded weights of golden/net_tiny_b2.npz, rebuilt from the recorded key/shape list."""
    shapes = {}
    with open(os.path.join(GOLD, keys_file)) as f:
        for line in f:
            k, shp = line.rstrip("\n").split(" ", 1)
            shapes[k] = ast.literal_eval(shp)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(shp, dtype=torch.long)
        else:
            sd[k] = torch.zeros(shp)
    # attn_mask buffers are structural (not random): take them from a freshly built encoder
    enc = O.OracleV3.__new__(O.OracleV3)
    import sys
    shim = os.path.join(os.path.dirname(O.__file__), "timm_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import timm
    m = timm.create_model("swinv2_tiny_window16_256")
    for k, v in m.state_dict().items():
        if k.endswith("attn_mask"):
            for pfx in encoder_prefixes:
                sd[pfx + k] = v.clone()
    del enc
    return seeded_state_dict(sd, seed)


def v1_tiny_state_dict(seed=0):
    """The seeded SOccDPT_V1 weights of golden/net_v1_tiny_b2.npz (two DPTs: depth_net.* (+ pretrained.* alias), seg_net.*)."""
    return tiny_state_dict(seed, "state_keys_v1_tiny.txt",
                           ("depth_net.pretrained.model.", "pretrained.model.", "seg_net.pretrained.model."))
