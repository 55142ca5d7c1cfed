"""Sparse occupancy outputs: the reference's `occupancy_grid_to_points`
(SOccDPT/utils/__init__.py:532-568, numpy on the CPU) on the device, from the dense grid the reference passes or directly
from the voxeliser's bit-packed mask (`SOccDPT(occupancy_output="packed")`: 1 MB instead of B x 100 MB per call).

    pts = occupancy_grid_to_points(grid[0])                      # dense (G0,G1,G2,C) CUDA tensor -> (n, 4) float64 CUDA tensor
    pts = packed_to_points(mask, grid_size, scale, num_classes)  # uint32 words from the voxeliser
"""
import ctypes

import numpy as np
import torch

from . import _cabi


def _occ_shape(grid_size, scale):
    # SOccDPT/utils/__init__.py:541-548: float(grid / scale) per axis, stored as fp32
    return np.array([float(grid_size[i] / scale[i]) for i in range(len(grid_size))], dtype=np.float32)


def mask_words(grid_size):
    return (int(grid_size[0]) * int(grid_size[1]) * int(grid_size[2]) + 7) // 8


def pack_grid(occupancy_grid):
    """dense (G0,G1,G2,C) CUDA tensor -> bit-packed int32 words (cells >= 0.5)."""
    if not occupancy_grid.is_cuda:
        raise _cabi.SoccdptError("soccdpt_b200 runs on CUDA devices only (no CPU fallback)")
    assert occupancy_grid.dim() == 4, "occupancy_grid must be 3D with one channel per class"
    g = occupancy_grid.to(torch.float32).contiguous()
    G = (ctypes.c_int * 3)(*g.shape[:3])
    mask = torch.empty(mask_words(g.shape[:3]), dtype=torch.int32, device=g.device)
    with torch.cuda.device(g.device):
        _cabi.check(_cabi.load().soccdpt_grid_pack_fwd(g.data_ptr(), G, int(g.shape[3]), mask.data_ptr(), _cabi.current_stream()),
                    "grid_pack")
    return mask


def packed_to_points(mask, grid_size=(256, 256, 32), scale=(2.0, 2.0, 0.666), num_classes=3):
    """bit-packed mask (int32 words, one call / one frame) -> (n, 4) float64 rows (x, y, z, class), the reference's order."""
    if not mask.is_cuda:
        raise _cabi.SoccdptError("soccdpt_b200 runs on CUDA devices only (no CPU fallback)")
    lib = _cabi.load()
    assert mask.dtype == torch.int32 and mask.is_contiguous() and mask.numel() >= mask_words(grid_size)
    G = (ctypes.c_int * 3)(*[int(v) for v in grid_size[:3]])
    occ = (ctypes.c_float * 3)(*[float(v) for v in _occ_shape(grid_size, scale)[:3]])
    dev = mask.device
    ws = torch.empty(int(lib.soccdpt_occupancy_points_workspace_bytes(G, num_classes)), dtype=torch.uint8, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = _cabi.current_stream()
        _cabi.check(lib.soccdpt_occupancy_points_fwd(mask.data_ptr(), G, occ, num_classes, None, 0, count.data_ptr(),
                                                     ws.data_ptr(), ws.numel(), st), "occupancy_points (count)")
        n = int(count.item())                                        # the one host sync: sizes the output
        pts = torch.empty((n, 4), dtype=torch.float64, device=dev)
        if n:
            _cabi.check(lib.soccdpt_occupancy_points_fwd(mask.data_ptr(), G, occ, num_classes, pts.data_ptr(), n, count.data_ptr(),
                                                         ws.data_ptr(), ws.numel(), st), "occupancy_points")
    return pts


def occupancy_grid_to_points(occupancy_grid, grid_size=(256, 256, 32), scale=(2.0, 2.0, 0.666), shift=(0.0, 0.0, 0.0)):
    """Same signature and result as the reference function (shift is unused there as well); CUDA tensors in and out."""
    assert len(occupancy_grid.shape) == 4, "occupancy_grid must be 3D with one channel per class"
    assert tuple(occupancy_grid.shape[:3]) == tuple(int(v) for v in grid_size[:3])
    return packed_to_points(pack_grid(occupancy_grid), grid_size, scale, int(occupancy_grid.shape[3]))
