"""Profiling driver: a few calls of the fused post-processing (B=8) and of the stand-alone voxeliser."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from soccdpt_b200 import SOccDPT
from soccdpt_b200.synthetic import write_calib_yaml
import soccdpt_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
net = SOccDPT(camera_intrinsics_yaml=write_calib_yaml("/tmp/prof_calib.yaml"), compute_occ=True)
g = torch.Generator().manual_seed(0)
base = torch.rand(B, 1, 16, 16, generator=g) * 0.2 + 0.02
inv = torch.nn.functional.interpolate(base, size=(256, 256), mode="bilinear")[:, 0].contiguous().cuda()
seg = torch.sigmoid(torch.randn(B, 3, 256, 256, generator=g)).cuda()
for _ in range(3):
    out = net.get_semantic_occupancy(inv, seg)
inv_up, seg_up = O.config5_maps(B, 1080, 1920, 3, seed=0)
inv_up, seg_up = inv_up.cuda(), seg_up.cuda()
for _ in range(3):
    net.voxelize(inv_up.clone(), seg_up)
torch.cuda.synchronize()
print("ok")
