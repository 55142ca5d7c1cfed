"""TEST INFRASTRUCTURE ONLY -- op-for-op torch restatement of the reference's post-processing
(SOccDPT/model/SOccDPT.py:60-130 rotate_points, :264-372 get_semantic_occupancy, :374-463 points_to_occupancy_grid).

Unlike oracle/voxel_oracle.c (a scalar C restatement of the arithmetic) this file issues the SAME ATen operators
in the same order as the reference, on whichever device it is given.  Two uses:
  * ``bench.py``'s second cpu_baseline leg (kind "port", the reference's own masked_select / nonzero / index_put_ cost);
  * the device-convention check (tests/test_gpu_device_convention.py): the reference divides a tensor by a host
    scalar (``X = (V - cx) * depth / fx``) -- ATen's CPU kernel divides, its CUDA kernel may multiply by the
    reciprocal.  Running this file on ``cuda`` and on ``cpu`` tells whether the reference itself is device dependent.
Pinned against the unmodified reference live in tests/test_oracle_vs_reference.py (CPU) and against tests/golden/voxel_*.npz.
"""
import numpy as np
import torch


def rotation_matrices(correction_angle, device):
    """rotate_points' three matrices, built exactly as SOccDPT.py:74-111 does (fp32 deg2rad / cos / sin on ``device``)."""
    ang = torch.tensor(correction_angle).to(device=device, dtype=torch.float32)
    a, b, c = (torch.deg2rad(t) for t in ang)
    Ra = torch.tensor([[1, 0, 0], [0, torch.cos(a), -torch.sin(a)], [0, torch.sin(a), torch.cos(a)]])
    Rb = torch.tensor([[torch.cos(b), 0, torch.sin(b)], [0, 1, 0], [-torch.sin(b), 0, torch.cos(b)]])
    Rc = torch.tensor([[torch.cos(c), -torch.sin(c), 0], [torch.sin(c), torch.cos(c), 0], [0, 0, 1]])
    return [m.to(device=device, dtype=torch.float32) for m in (Ra, Rb, Rc)]


def rotate_points(points, correction_angle, device):
    for R in rotation_matrices(correction_angle, device):
        points = torch.einsum("bnm,mj->bnj", points, R)
    return points


def points_to_occupancy_grid(points, semantics_3D, geom, device):
    """SOccDPT.py:374-463 (occupancy_conv is nn.Identity, SOccDPT.py:244)."""
    B = semantics_3D.shape[0]
    G, C = geom.grid_size, geom.num_classes
    grid = torch.zeros((B, G[0], G[1], G[2], C), dtype=torch.float32, device=device)
    mask = (~torch.isinf(points).any(dim=-1)) & (~torch.isnan(points).any(dim=-1))
    points = torch.masked_select(points, mask.unsqueeze(-1)).reshape(-1, 3)
    semantics_3D = torch.masked_select(semantics_3D, mask.unsqueeze(-1)).reshape(-1, C)
    occ = torch.tensor(geom.occupancy_shape).to(device=device, dtype=torch.float32)
    gsz = torch.tensor(G).to(device=device, dtype=torch.float32)
    ijk = (points / occ * gsz).type(torch.int64)
    mask = ((0 < ijk[..., 0]) & (ijk[..., 0] < G[0]) & (0 < ijk[..., 1]) & (ijk[..., 1] < G[1])
            & (0 < ijk[..., 2]) & (ijk[..., 2] < G[2]))
    ijk = torch.masked_select(ijk, mask.unsqueeze(-1)).reshape(-1, 3)
    semantics_3D = torch.masked_select(semantics_3D, mask.unsqueeze(-1)).reshape(-1, C)
    sem_idx = semantics_3D.nonzero(as_tuple=False)
    bi = torch.cat([ijk[sem_idx[:, 0]], sem_idx[:, 1].view(-1, 1)], dim=1)
    grid[:, bi[:, 0], bi[:, 1], bi[:, 2], bi[:, 3]] += 1
    return grid


def get_semantic_occupancy(inv_depth, segmentation, geom, compute_occ=True, device=None):
    """SOccDPT.py:264-372 with point_compute_method == "torch".  Tensors are taken on ``device`` (default: where
    ``inv_depth`` lives); returns the reference's 4-tuple."""
    device = inv_depth.device if device is None else torch.device(device)
    inv_depth, segmentation = inv_depth.to(device), segmentation.to(device)
    H, W = geom.height, geom.width
    if inv_depth.dim() == 3:
        inv_depth = inv_depth.unsqueeze(1)
    inv_depth = torch.nn.functional.interpolate(inv_depth, size=(H, W), mode="bicubic", align_corners=False).squeeze()
    segmentation = torch.nn.functional.interpolate(segmentation, size=(H, W), mode="nearest").squeeze()
    if inv_depth.dim() == 2:
        inv_depth = inv_depth.unsqueeze(0)
    return _unproject_and_voxelise(inv_depth, segmentation, geom, compute_occ, device)


def _unproject_and_voxelise(inv_depth, segmentation, geom, compute_occ, device):
    """The stage after the two resizes (SOccDPT.py:289-372); ``inv_depth`` (B,H,W) is modified in place like upstream."""
    H, W = geom.height, geom.width
    depth = inv_depth
    depth[depth < 1e-8] = 1e-8
    depth = 1.0 / depth
    depth[torch.isinf(depth)] = float("inf")
    depth[torch.isnan(depth)] = float("inf")
    U, V = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32),
                          torch.arange(W, device=device, dtype=torch.float32), indexing="ij")
    U = U.unsqueeze(0).repeat(inv_depth.shape[0], 1, 1)
    V = V.unsqueeze(0).repeat(inv_depth.shape[0], 1, 1)
    # cx, cy, fx, fy are numpy float64 scalars upstream (self.intrinsic_matrix[...]): a host scalar against an fp32 tensor
    cx, cy, fx, fy = (np.float64(v) for v in (geom.cx, geom.cy, geom.fx, geom.fy))
    X = (V - cx) * depth / fx
    Y = (U - cy) * depth / fy
    points_batched = torch.stack([X, Y, depth], dim=3)
    C = geom.num_classes
    semantics_3D = segmentation.reshape(-1, C, H * W).permute(0, 2, 1)
    points_3D = points_batched.reshape(-1, H * W, C)      # upstream reshapes with num_classes (== 3 coordinates)
    # upstream indexes dim 1 (the first three PIXELS, not the coordinates): kept, on the shared storage of points_batched
    for i in range(3):
        points_3D[:, i] = points_3D[:, i] * geom.pc_scale[i] + geom.pc_shift[i]
    points_3D = rotate_points(points_3D, geom.correction_angle, device)
    grid = points_to_occupancy_grid(points_3D, semantics_3D, geom, device) if compute_occ else None
    return inv_depth, segmentation, points_batched, grid


def voxelize(inv_depth_up, seg_up, geom, compute_occ=True, device="cpu"):
    """Camera-resolution maps in (config 5): (B,H,W) and (B,C,H,W) -> (clamped inv depth, points, grid) on ``device``."""
    dev = torch.device(device)
    inv = torch.as_tensor(inv_depth_up).to(dev).clone()
    seg = torch.as_tensor(seg_up).to(dev)
    inv_c, _, pts, grid = _unproject_and_voxelise(inv, seg, geom, compute_occ, dev)
    return inv_c, pts, grid
