"""Window attention (soccdpt_window_attention_fwd) per encoder stage of dpt_swin2_tiny_256 at B frames (CUDA events).
SOCCDPT_ATTN_WS=0 selects round 1's one-CTA-per-(window, head) kernel, the default is the warp-specialised persistent one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuda_ops as K

B = int(os.environ.get("B", "64"))
STAGES = [("S0", 64, 96, 3, 16), ("S1", 32, 192, 6, 16), ("S2", 16, 384, 12, 16), ("S3", 8, 768, 24, 8)]
if os.environ.get("ONLY"):
    STAGES = [s for s in STAGES if s[0] == os.environ["ONLY"]]
g = torch.Generator().manual_seed(0)
for name, res, C, heads, ws in STAGES:
    qkv = torch.randn(B, res * res, 3 * C, generator=g).bfloat16().cuda()
    bias = (torch.rand(heads, (2 * ws - 1) ** 2, generator=g) * 16).cuda()
    scale = (torch.rand(heads, generator=g) * 15 + 5).cuda()
    for shift in (0, ws // 2 if res > ws else 0):
        for _ in range(3):
            K.window_attention(qkv, bias, scale, B, res, res, C, heads, ws, shift)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            K.window_attention(qkv, bias, scale, B, res, res, C, heads, ws, shift)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        exps = B * res * res * heads * ws * ws
        print(f"{name} shift={shift:2d}: {ms * 1e3:8.1f} us   {exps / ms / 1e6:8.1f} G exp/s  (MUFU floor {exps / (148 * 16 * 1.9e9) * 1e6:6.1f} us)")
