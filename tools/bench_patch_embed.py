"""Patch embed (conv 4x4 stride 4 + LayerNorm) timing: tensor-core kernel vs the fp32 CUDA-core kernel (SOCCDPT_PATCH_EMBED_FP32=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuda_ops as K

for B, S, E in ((64, 256, 96), (32, 384, 128)):
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(B, 3, S, S, generator=g).cuda() for _ in range(4)]          # 4 x 50 MB > L2 with the outputs
    w, b = (torch.randn(E, 48, generator=g) * 0.2).cuda(), (torch.randn(E, generator=g) * 0.1).cuda()
    lw, lb = (torch.rand(E, generator=g) + 0.5).cuda(), (torch.rand(E, generator=g) - 0.5).cuda()
    for _ in range(3):
        K.patch_embed(xs[0], w, b, lw, lb)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    s.record()
    for i in range(reps):
        K.patch_embed(xs[i & 3], w, b, lw, lb)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / reps * 1e3
    nbytes = B * 3 * S * S * 4 + B * (S // 4) ** 2 * E * 6
    print(f"B={B} S={S} E={E}: {us:7.1f} us   {nbytes / us / 1e3:7.1f} GB/s (in + bf16 + fp32 out; includes the wrapper's allocations)")
