"""VERDICT r1 'missing 6': is the reference's own eager path device dependent?

SOccDPT.py:311-313 divides an fp32 tensor by a host scalar (``X = (V - cx) * depth / fx``).  ATen's CPU kernel performs
a true division; ATen's CUDA kernel for tensor / scalar multiplies by the reciprocal of the scalar (computed in the
op-math type).  The product kernel follows the CPU convention (the fixtures were recorded on CPU).  This test runs the
op-for-op ATen restatement (oracle/torch_postprocess.py) on cuda and on cpu and RECORDS the outcome; it asserts
  * product kernel == CPU convention, bit for bit (the parity claim), and
  * wherever the CUDA eager path differs from the CPU path, it differs by at most two ulps in X / Y (a * (1/b) against
    a / b) and the occupancy grids differ only in voxels whose generating point sits on a cell boundary.
Measured on B200 / torch 2.11 (profiles/r2_device_convention.jsonl): 25-43 % of the coordinates differ by 1-2 ulps, the
grids are identical -- the reference IS device dependent in the last bits of its points.
"""
import json
import os

import numpy as np
import pytest
import torch

import golden_util as GU
import soccdpt_oracle as O
import torch_postprocess as TP
from soccdpt_b200 import SOccDPT
from soccdpt_b200.synthetic import write_calib_yaml

pytestmark = pytest.mark.gpu


def _ulp_diff(a, b):
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


@pytest.mark.parametrize("name", ["small_b2", "ragged_b3_grid64", "full_b1"])
def test_reference_eager_path_on_cuda_vs_cpu_vs_kernel(name, tmp_path):
    z, calib, geom, inv, seg = GU.load_voxel_case(name)
    inv_cpu, pts_cpu, grid_cpu = TP.voxelize(inv, seg, geom, device="cpu")
    inv_gpu, pts_gpu, grid_gpu = TP.voxelize(inv, seg, geom, device="cuda")
    vox = SOccDPT(camera_intrinsics_yaml=write_calib_yaml(str(tmp_path / "c.yaml"), calib), compute_occ=True,
                  grid_size=geom.grid_size, scale=tuple(float(s) for s in z["scale"]))
    inv_k = inv.cuda().clone()
    pts_k, grid_k = vox.voxelize(inv_k, seg.cuda())
    # 1) the product kernel follows the CPU convention exactly
    assert np.array_equal(pts_k.cpu().numpy().view(np.uint32), pts_cpu.numpy().view(np.uint32))
    assert torch.equal(grid_k.cpu(), grid_cpu)
    # 2) the reference's own eager path on CUDA, against its CPU path
    pc, pg = pts_cpu.numpy(), pts_gpu.cpu().numpy()
    finite = np.isfinite(pc) & np.isfinite(pg)
    same_nonfinite = np.array_equal(np.isnan(pc), np.isnan(pg)) and np.array_equal(np.isinf(pc), np.isinf(pg))
    ulps = _ulp_diff(pc[finite], pg[finite])
    n_diff = int((ulps != 0).sum())
    grid_diff = int((grid_gpu.cpu() != grid_cpu).sum().item())
    record = dict(case=name, points=int(pc.size), points_differing=n_diff, max_ulp=int(ulps.max()) if ulps.size else 0,
                  same_nonfinite_pattern=bool(same_nonfinite), grid_cells_differing=grid_diff,
                  grid_cells_set=int(grid_cpu.sum().item()), torch=torch.__version__,
                  device=torch.cuda.get_device_name(0))
    out = os.environ.get("SOCCDPT_CONVENTION_LOG", "")
    if out:
        with open(out, "a") as f:
            f.write(json.dumps(record) + "\n")
    print("device convention:", json.dumps(record))
    assert same_nonfinite
    assert record["max_ulp"] <= 2, record
    # a two-ulp move of a coordinate can only move a point across a cell boundary: a handful of cells at most
    assert grid_diff <= max(8, record["grid_cells_set"] // 1000), record
    # 3) device_convention="cuda": the product kernel restates the CUDA eager path bit for bit as well
    vox_c = SOccDPT(camera_intrinsics_yaml=write_calib_yaml(str(tmp_path / "c2.yaml"), calib), compute_occ=True,
                    grid_size=geom.grid_size, scale=tuple(float(s) for s in z["scale"]), device_convention="cuda")
    inv_c = inv.cuda().clone()
    pts_c, grid_c = vox_c.voxelize(inv_c, seg.cuda())
    assert np.array_equal(pts_c.cpu().numpy().view(np.uint32), pts_gpu.cpu().numpy().view(np.uint32))
    assert np.array_equal(inv_c.cpu().numpy().view(np.uint32), inv_gpu.cpu().numpy().view(np.uint32))
    assert torch.equal(grid_c, grid_gpu)
