"""Host-side preparation of the constants the post-processing kernels need, computed exactly as the
reference computes them on its CPU path so that the CUDA result is bit-identical to it."""
import os

import numpy as np
import torch
import yaml

from . import _cabi


def load_calib(path):
    """Camera YAML as read by SOccDPT.__init__ (SOccDPT/model/SOccDPT.py:190-228)."""
    path = os.path.expanduser(path)
    with open(path, "r") as stream:
        cam = yaml.load(stream, Loader=yaml.FullLoader)
    for k in ("Camera.k1", "Camera.k2", "Camera.p1", "Camera.p2", "Camera.fx", "Camera.fy", "Camera.cx",
              "Camera.cy", "Camera.width", "Camera.height"):
        cam[k]  # KeyError like the reference
    return cam


def rotation_matrices(correction_angle, device="cpu"):
    """Ra, Rb, Rc of rotate_points (SOccDPT.py:74-111): fp32 deg2rad / cos / sin on tensors of ``device`` -- the reference
    builds them where the model lives, and libm's and CUDA's cos / sin may differ in the last bit."""
    a, b, c = torch.tensor(correction_angle).to(device=device, dtype=torch.float32)
    a, b, c = torch.deg2rad(a), torch.deg2rad(b), torch.deg2rad(c)
    Ra = torch.tensor([[1, 0, 0], [0, torch.cos(a), -torch.sin(a)], [0, torch.sin(a), torch.cos(a)]])
    Rb = torch.tensor([[torch.cos(b), 0, torch.sin(b)], [0, 1, 0], [-torch.sin(b), 0, torch.cos(b)]])
    Rc = torch.tensor([[torch.cos(c), -torch.sin(c), 0], [torch.sin(c), torch.cos(c), 0], [0, 0, 1]])
    return torch.stack([Ra, Rb, Rc]).to(torch.float32).reshape(-1).tolist()


def make_geometry(fx, fy, cx, cy, height, width, num_classes, grid_size, occupancy_shape, pc_scale, pc_shift,
                  correction_angle, device_convention="cpu", device=None):
    """``device_convention``: "cpu" (default) restates the reference's eager path on CPU tensors -- the convention of its
    fixtures; "cuda" restates what the same ATen ops do on a CUDA device (tensor / host scalar as a multiply by the fp32
    reciprocal, rotation matrices from the CUDA cos / sin)."""
    assert device_convention in ("cpu", "cuda")
    g = _cabi.Geometry()
    # numpy.float64 scalars meet fp32 tensors in the reference -> the scalar is cast to fp32
    g.fx, g.fy, g.cx, g.cy = (float(np.float32(v)) for v in (fx, fy, cx, cy))
    g.height, g.width, g.num_classes = int(height), int(width), int(num_classes)
    for i in range(3):
        g.grid[i] = int(grid_size[i])
        g.occ_shape[i] = float(np.float32(occupancy_shape[i]))
        g.pc_scale[i] = float(np.float32(pc_scale[i]))
        g.pc_shift[i] = float(np.float32(pc_shift[i]))
    cuda = device_convention == "cuda"
    for i, v in enumerate(rotation_matrices(correction_angle, device if cuda and device is not None else "cpu")):
        g.rot[i] = v
    g.scalar_div_by_reciprocal = 1 if cuda else 0
    # ATen's CUDA div kernel: inv_b = 1.0 / b in DOUBLE on the Python-float scalar, cast to the op-math type (fp32)
    g.rcp_fx, g.rcp_fy = float(np.float32(1.0 / np.float64(fx))), float(np.float32(1.0 / np.float64(fy)))
    return g
