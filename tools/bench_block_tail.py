"""Fused Swin block tail (soccdpt_swin_block_tail_fwd) against the un-fused launch sequence it replaces, per encoder stage
(dpt_swin2_tiny_256 shapes at B frames; CUDA events; each variant cycles through enough independent buffer sets to exceed L2)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuda_ops as K

B = int(os.environ.get("B", "64"))
STAGES = [("S0", B * 4096, 96), ("S1", B * 1024, 192)] + ([("S2", B * 256, 384)] if os.environ.get("S2") else [])
if os.environ.get("ONLY"):
    STAGES = [st for st in STAGES if st[0] == os.environ["ONLY"]]
PEAK_HBM = 6539.9
g = torch.Generator().manual_seed(0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


print(f"{'case':34s} {'fused ms':>9s} {'unfused ms':>11s} {'GB/s (fused, algorithmic)':>26s}")
for name, M, C in STAGES:
    HID = 4 * C
    x = torch.randn(M, C, generator=g).bfloat16().cuda()
    w1 = (torch.randn(HID, C, generator=g) / math.sqrt(C)).bfloat16().cuda()
    b1 = torch.randn(HID, generator=g).cuda()
    w2 = (torch.randn(C, HID, generator=g) / math.sqrt(HID)).bfloat16().cuda()
    wp = (torch.randn(C, C, generator=g) / math.sqrt(C)).bfloat16().cuda()
    b2 = torch.randn(C, generator=g).cuda()
    ga, be = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.rand(C, generator=g) - 0.5).cuda()
    master = torch.randn(M, C, generator=g).cuda()
    y = torch.empty(M, C, dtype=torch.bfloat16, device="cuda")
    x4, w1c, w2c, wpc = x.view(1, 1, M, C), w1.view(HID, 1, C), w2.view(C, 1, HID), wp.view(C, 1, C)
    bytes_alg = M * C * (2 + 4 + 4 + 2)

    if C <= 256:
        f_mlp = timed(lambda: K.swin_block_tail(x, w2, b2, ga, be, master, w1, b1, y=y))
    else:       # wide rows: fc1 + GELU as a plain GEMM, then fc2 -> norm -> residual fused
        def fused_wide():
            h, _, _ = K.conv(x4, w1c, bias=b1, act=2)
            K.swin_block_tail(h.view(M, HID), w2, b2, ga, be, master, y=y)
        f_mlp = timed(fused_wide)
    f_proj = timed(lambda: K.swin_block_tail(x, wp, b2, ga, be, master, y=y))

    def unfused_mlp():
        h, _, _ = K.conv(x4, w1c, bias=b1, act=2)
        t, _, _ = K.conv(h, w2c, bias=b2)
        K.layernorm_master(t.view(M, C), master, True, ga, be)

    def unfused_proj():
        t, _, _ = K.conv(x4, wpc, bias=b2)
        K.layernorm_master(t.view(M, C), master, True, ga, be)

    u_mlp, u_proj = timed(unfused_mlp), timed(unfused_proj)
    print(f"{name + ' mlp  C=' + str(C) + ' M=' + str(M):34s} {f_mlp:9.4f} {u_mlp:11.4f} {bytes_alg / f_mlp / 1e6:14.0f} ({bytes_alg / f_mlp / 1e6 / PEAK_HBM:.2f} of peak)")
    print(f"{name + ' proj C=' + str(C) + ' M=' + str(M):34s} {f_proj:9.4f} {u_proj:11.4f} {bytes_alg / f_proj / 1e6:14.0f} ({bytes_alg / f_proj / 1e6 / PEAK_HBM:.2f} of peak)")
