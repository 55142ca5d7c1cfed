"""Depth-head tail (csrc/depth_head.cu) alone: 9 x 32 tap planes at low resolution -> depth map, at the bench shapes.
    python tools/bench_depth_tail.py            (SOCCDPT_LIB=build/variants/<name>/lib.so for an A/B build)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuda_ops as K

for B, h in ((64, 128), (32, 192)):
    g = torch.Generator().manual_seed(0)
    Ts = [torch.randn(B, h, h, 288, generator=g, dtype=torch.bfloat16).cuda() for _ in range(2)]   # 2 x 0.6 GB: never L2-resident
    b2, pw, pb = (torch.randn(32, generator=g) * 0.1).cuda(), (torch.randn(32, generator=g) * 0.2).cuda(), torch.zeros(1).cuda()
    for _ in range(3):
        K.depth_tail(Ts[0], b2, pw, pb)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    s.record()
    for i in range(reps):
        K.depth_tail(Ts[i & 1], b2, pw, pb)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / reps * 1e3
    nbytes = Ts[0].numel() * 2 + B * 4 * h * h * 4
    print(f"B={B} {h}x{h} -> {2*h}x{2*h}: {us:7.1f} us   {nbytes / us / 1e3:7.1f} GB/s algorithmic (T read once + fp32 depth written)")
