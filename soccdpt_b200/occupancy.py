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


class OccupancyCounter:
    """The reference's ground-truth voxeliser ``OccupancyProcessor.transform_points_to_occupancy_grid_vect``
    (SOccDPT/datasets/bdd_helper.py:238-362) on the device: points with integer class ids are counted per (voxel, class)
    with warp-aggregated atomics, then thresholded.  Same constructor keywords for the part of the class this path uses.

        oc = OccupancyCounter(grid_size=(256, 256, 32), scale=(2.0, 2.0, 0.666), num_classes=3, point_count_threshold=10)
        out = oc.transform_points_to_occupancy_grid_vect(cam_points, semantics)      # CUDA tensors, (n,3) f32|f64 and (n,) int
        out["occupancy_grid"]    bool (G0,G1,G2,C)   count >  threshold   (bdd_helper.py:357)
        out["occupancy_points"]  f64  (m,4)          count >= threshold   (bdd_helper.py:340-355), class-major argwhere order
        out["counts"]            i32  (G0,G1,G2,C)   (extension)  out["labels"] u8 (G0,G1,G2): 0 empty, 1 + argmax class
    """

    def __init__(self, grid_size=(256, 256, 32), scale=(2.0, 2.0, 0.666), num_classes=3, point_count_threshold=10, **_unused):
        self.grid_size = tuple(int(v) for v in grid_size)
        self.scale = tuple(scale)
        self.num_classes = int(num_classes)
        self.point_count_threshold = point_count_threshold
        self.occupancy_shape = _occ_shape(self.grid_size, self.scale)

    def count(self, cam_points, semantics, counts=None):
        """Accumulates into (or creates) the int32 count grid."""
        if not cam_points.is_cuda or not semantics.is_cuda:
            raise _cabi.SoccdptError("soccdpt_b200 runs on CUDA devices only (no CPU fallback)")
        assert cam_points.dim() == 2 and cam_points.shape[1] == 3 and cam_points.dtype in (torch.float32, torch.float64)
        assert cam_points.shape[0] == semantics.shape[0], \
            "cam_points and semantics must have the same number of points, but got {} and {}".format(cam_points.shape[0], semantics.shape[0])
        assert semantics.dtype in (torch.int32, torch.int64)
        dev = cam_points.device
        G, C = self.grid_size, self.num_classes
        if counts is None:
            counts = torch.zeros((G[0], G[1], G[2], C), dtype=torch.int32, device=dev)
        p, s = cam_points.contiguous(), semantics.contiguous()
        bad = torch.zeros(1, dtype=torch.int64, device=dev)
        Gc = (ctypes.c_int * 3)(*G)
        occ = (ctypes.c_float * 3)(*[float(v) for v in self.occupancy_shape[:3]])
        with torch.cuda.device(dev):
            _cabi.check(_cabi.load().soccdpt_voxel_count_fwd(
                p.data_ptr(), int(p.dtype == torch.float64), s.data_ptr(), int(s.dtype == torch.int64), int(p.shape[0]), Gc, occ, C,
                counts.data_ptr(), bad.data_ptr(), _cabi.current_stream()), "voxel_count")
        self._bad = bad
        return counts

    def finish(self, counts, want_points=True):
        dev = counts.device
        G, C = self.grid_size, self.num_classes
        grid_gt = torch.empty((G[0], G[1], G[2], C), dtype=torch.uint8, device=dev)
        labels = torch.empty(G, dtype=torch.uint8, device=dev)
        mask = torch.empty(mask_words(G), dtype=torch.int32, device=dev)
        Gc = (ctypes.c_int * 3)(*G)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.load().soccdpt_voxel_count_finish_fwd(
                counts.data_ptr(), Gc, C, ctypes.c_float(float(self.point_count_threshold)), grid_gt.data_ptr(), mask.data_ptr(),
                labels.data_ptr(), _cabi.current_stream()), "voxel_count_finish")
        out = {"occupancy_grid": grid_gt.view(torch.bool), "labels": labels, "counts": counts, "mask_ge": mask}
        if want_points:
            out["occupancy_points"] = packed_to_points(mask, G, self.scale, C)
        return out

    def transform_points_to_occupancy_grid_vect(self, cam_points_orig, semantics):
        counts = self.count(cam_points_orig, semantics)
        out = self.finish(counts)
        if int(self._bad.item()):          # numpy's np.add.at raises here; so do we (after the fact: one host sync, already paid by the points list)
            raise IndexError(f"{int(self._bad.item())} class ids are out of bounds for {self.num_classes} classes")
        return out
