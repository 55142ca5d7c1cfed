"""One stand-alone voxeliser call shape for ncu (BASELINE config 5: camera-resolution maps, 256x256x32 grid, B frames)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import soccdpt_oracle as O
from soccdpt_b200 import SOccDPT
from soccdpt_b200.synthetic import write_calib_yaml

B = int(os.environ.get("B", "64"))
net = SOccDPT(camera_intrinsics_yaml=write_calib_yaml("/tmp/prof_vox_calib.yaml"), compute_occ=True)
inv, seg = O.config5_maps(B, 1080, 1920, 3, seed=B)
inv_d, seg_d = inv.cuda(), seg.cuda()
work = inv_d.clone()
for _ in range(3):
    net.voxelize(work, seg_d)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    net.voxelize(work, seg_d)
e.record()
torch.cuda.synchronize()
print(f"B={B}: {s.elapsed_time(e) / 10:.3f} ms per call")
