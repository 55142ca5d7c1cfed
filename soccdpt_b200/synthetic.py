"""Synthetic inputs for benchmarks and parity tests (no datasets, no checkpoints, no network).

``seeded_state_dict`` overwrites EVERY learnable tensor of a SOccDPT state_dict from one
seeded CPU generator: timm's default SwinV2 init zeroes ``norm1/norm2`` of every block
(res-post-norm), which would turn the whole encoder into an identity and leave attention
and MLP kernels untested.  The same dict is loaded into the
reference / oracle and into this package, so both sides see identical weights.
"""
import math
import os

import torch

SYNTHETIC_CALIB = {
    # reference media/manydepth/intrinsics.json:1-5, frame size bengaluru_driving_dataset.py:118-121
    "Camera.fx": 1250.6, "Camera.fy": 1254.8, "Camera.cx": 978.4, "Camera.cy": 562.1,
    "Camera.k1": 0.0, "Camera.k2": 0.0, "Camera.p1": 0.0, "Camera.p2": 0.0,
    "Camera.width": 1920, "Camera.height": 1080,
}


def write_calib_yaml(path, calib=None):
    """Writes the camera YAML the reference constructor reads (SOccDPT.py:190-228)."""
    calib = dict(SYNTHETIC_CALIB if calib is None else calib)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        for k, v in calib.items():
            if k in ("Camera.width", "Camera.height"):
                v = int(v)
            f.write(f"{k}: {v}\n")
    return path


_ALIAS = "pretrained."
_CANON = "depth_net.pretrained."


def seeded_state_dict(state_dict, seed=0, residual_gain=1.0):
    """Returns a new dict with the same keys/shapes, values drawn deterministically (sorted key
    order, one CPU generator).  ``pretrained.*`` keys alias ``depth_net.pretrained.*`` (same module registered twice).
    ``residual_gain`` scales the last GroupNorm (``norm3``) of every ResNetV2 bottleneck of the hybrid encoder: 1.0 is
    plain random init (numerically chaotic under bf16 storage, see tests/test_gpu_network_hybrid.py), 0.1 damps the
    residual branches the way trained weights / timm's ``zero_init_last`` do."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(state_dict.keys()):
        v = state_dict[k]
        if k.startswith(_ALIAS):
            continue
        shape = tuple(v.shape)
        if k.endswith("attn_mask") or k.endswith("num_batches_tracked"):
            out[k] = v.detach().clone().cpu()
            continue
        # BatchNorms: SOccDPT_V3's seg head; SOccDPT_V1's segmentation DPT (residual conv units, head, auxlayer)
        is_norm = (".norm" in k) or k.startswith("seg_head.1.") or (
            k.startswith("seg_net.") and (".bn1." in k or ".bn2." in k or ".output_conv.1." in k or ".auxlayer.1." in k))
        if k.endswith("running_mean"):
            t = torch.randn(shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        elif k.endswith("logit_scale"):
            t = torch.rand(shape, generator=g) * (math.log(20.0) - math.log(5.0)) + math.log(5.0)
        elif is_norm and k.endswith("weight"):
            t = torch.rand(shape, generator=g) + 0.5
        elif k.endswith("bias"):
            t = torch.rand(shape, generator=g) * 0.2 - 0.1
        elif ".scratch." in k or k.startswith("seg_head."):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        elif "cpb_mlp" in k:
            t = torch.randn(shape, generator=g) * 0.5
        else:
            t = torch.randn(shape, generator=g) * 0.02
        # keep the synthetic inverse depth in ~[0.03, 0.3] (depth 3-30 m) so that unprojected points
        # land inside the 128 m x 128 m x 48 m occupancy volume instead of the never-filled k=0 plane
        if residual_gain != 1.0 and (k.endswith(".norm3.weight") or k.endswith(".norm3.bias")):
            t = t * residual_gain
        if k.endswith("scratch.output_conv.4.weight") and not k.startswith("seg_net."):
            t = t * 0.03
        if k.endswith("scratch.output_conv.4.bias") and not k.startswith("seg_net."):
            t = t * 0.1 + 0.05
        out[k] = t.to(torch.float32)
    for k in state_dict.keys():
        if k.startswith(_ALIAS):
            out[k] = out[_CANON + k[len(_ALIAS):]]
    return out


def synthetic_frames(batch, size=256, seed=0):
    """x = randn(B,3,size,size) from a seeded CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, size, size, generator=g)
