"""Per-shape timing of soccdpt_conv_fwd (CUDA events, L2 flushed between runs by cycling large buffers)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuda_ops as K

B = int(os.environ.get("B", "64"))
# name, N,H,W,Cin,Cout,k, act, nres, relu_copy, proj_n
SHAPES = [
    ("rcu64  3x3 256->256 relu", B, 64, 64, 256, 256, 3, 1, 0, False, 0),
    ("rcu64  3x3 256->256 +res", B, 64, 64, 256, 256, 3, 0, 1, False, 0),
    ("seg    3x3 256->256 @128 proj3", B, 128, 128, 256, 256, 3, 1, 0, False, 3),
    ("dep0   3x3 256->128 @128", B, 128, 128, 256, 128, 3, 0, 0, False, 0),
    ("dep2   3x3 128->32 @256 proj1", B, 256, 256, 128, 32, 3, 1, 0, False, 1),
    ("outc   1x1 256->256 @64", B, 64, 64, 256, 256, 1, 0, 0, False, 0),
    ("rn1    3x3  96->256 @64 +relucopy", B, 64, 64, 96, 256, 3, 0, 0, True, 0),
    ("rcu32  3x3 256->256 relu", B, 32, 32, 256, 256, 3, 1, 0, False, 0),
    ("S0 qkv  96->288", 1, 1, B * 4096, 96, 288, 1, 0, 0, False, 0),
    ("S0 fc1  96->384 gelu", 1, 1, B * 4096, 96, 384, 1, 2, 0, False, 0),
    ("S0 fc2 384->96", 1, 1, B * 4096, 384, 96, 1, 0, 0, False, 0),
    ("S2 qkv 384->1152", 1, 1, B * 256, 384, 1152, 1, 0, 0, False, 0),
    ("S2 fc1 384->1536 gelu", 1, 1, B * 256, 384, 1536, 1, 2, 0, False, 0),
    ("S2 fc2 1536->384", 1, 1, B * 256, 1536, 384, 1, 0, 0, False, 0),
    ("S3 fc1 768->3072 gelu", 1, 1, B * 64, 768, 3072, 1, 2, 0, False, 0),
]
only = os.environ.get("ONLY")
g = torch.Generator().manual_seed(0)
print(f"{'shape':38s} {'ms':>8s} {'TFLOP/s':>9s}")
for name, N, H, W, Cin, Cout, k, act, nres, rc, pn in SHAPES:
    if only and only not in name:
        continue
    x = (torch.randn(N, H, W, Cin, generator=g) * 0.5).bfloat16().cuda()
    w = K.pack_conv_weight(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    res = [torch.randn(N, H, W, Cout).bfloat16().cuda() for _ in range(nres)]
    proj = ((torch.randn(pn, Cout) * 0.1).cuda(), torch.randn(pn).cuda(), True) if pn else None
    kw = dict(bias=b, act=act, res1=res[0] if nres else None, want_y=pn == 0, want_relu=rc, proj=proj)
    for _ in range(3):
        K.conv(x, w, **kw)
    torch.cuda.synchronize()
    reps = 10
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        K.conv(x, w, **kw)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    fl = 2.0 * N * H * W * Cout * k * k * Cin
    print(f"{name:38s} {ms:8.3f} {fl / ms / 1e9:9.1f}")
