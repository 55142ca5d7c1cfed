"""Window attention per encoder stage of dpt_swin2_tiny_256 at B frames (CUDA events, back-to-back launches on one buffer set per
stage -- S0's qkv is 151 MB, larger than L2).  KERNEL=tma: the TMA-fed pipelined kernel on pre-normalised operands
(soccdpt_window_attention_normed_fwd, 16x16 windows); KERNEL=old: round 1's kernel (soccdpt_window_attention_fwd); default both."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
import cuda_ops as K

B = int(os.environ.get("B", "64"))
REPS = int(os.environ.get("REPS", "10"))
KERNELS = {"tma": ("tma",), "old": ("old",)}.get(os.environ.get("KERNEL", ""), ("old", "tma"))
STAGES = [("S0", 64, 96, 3, 16), ("S1", 32, 192, 6, 16), ("S2", 16, 384, 12, 16), ("S3", 8, 768, 24, 8)]
if os.environ.get("ONLY"):
    STAGES = [s for s in STAGES if s[0] in os.environ["ONLY"].split(",")]
g = torch.Generator().manual_seed(0)
for name, res, C, heads, ws in STAGES:
    qkv = torch.randn(B, res * res, 3 * C, generator=g).bfloat16().cuda()
    bias = (torch.rand(heads, (2 * ws - 1) ** 2, generator=g) * 16).cuda()
    scale = (torch.rand(heads, generator=g) * 15 + 5).cuda()
    q, k, v = qkv.float().view(B, res * res, 3, heads, 32).unbind(2)
    qkvn = torch.stack((F.normalize(q, dim=-1) * (scale * 1.4426950408889634).view(1, 1, heads, 1), F.normalize(k, dim=-1), v),
                       dim=2).reshape(B, res * res, 3 * C).bfloat16().contiguous()
    for shift in (0, ws // 2 if res > ws else 0):
        line = f"{name} shift={shift:2d}:"
        for kern in KERNELS:
            if kern == "tma" and ws != 16:
                continue
            fn = ((lambda: K.window_attention_normed(qkvn, bias, scale, B, res, res, C, heads, shift)) if kern == "tma" else
                  (lambda: K.window_attention(qkv, bias, scale, B, res, res, C, heads, ws, shift)))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(REPS):
                fn()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / REPS
            line += f"  {kern} {ms * 1e3:7.1f} us"
        exps = B * res * res * heads * ws * ws
        print(line + f"   (MUFU floor {exps / (148 * 16 * 1.965e9) * 1e6:5.1f} us)")
