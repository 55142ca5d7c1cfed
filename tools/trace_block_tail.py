"""Hand-off trace of the fused Swin block tail's CTA 0 (trace build only):
    SOCCDPT_NVCC_FLAGS=-DSOCCDPT_TAIL_TRACE python tools/build_variant.py trace
    SOCCDPT_LIB=build/variants/trace/lib.so python tools/trace_block_tail.py
Events: GELU warp (per hidden chunk g): 0 waits for D1, 1 has it, 2 TMEM loaded, 3 H handed to the MMA warp; MMA warp: 4 / 5 before / after
the wait for H of chunk g, 6 / 7 before / after the wait for the activation tile, 8 / 9 before / after the wait for a free D2;
LayerNorm warp (per tile k): 10 / 11 before / after the wait for D2, 12 statistics done, 13 tile written; 14-19 the phases of the tile's
second 32-column block (TMEM load, normalise + transposition stores, cp.async wait, residual add + global stores, next request)."""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuda_ops as K
from soccdpt_b200 import _cabi

lib = ctypes.CDLL(_cabi.lib_path())
EV, IDS = 24, 64
B = 64
g = torch.Generator().manual_seed(0)
for name, M, C, mlp in (("S0 mlp", B * 4096, 96, True), ("S1 mlp", B * 1024, 192, True), ("S0 proj", B * 4096, 96, False)):
    HID = 4 * C
    x = torch.randn(M, C, generator=g).bfloat16().cuda()
    w1 = (torch.randn(HID, C, generator=g) / math.sqrt(C)).bfloat16().cuda()
    b1 = torch.randn(HID, generator=g).cuda()
    w2 = (torch.randn(C, HID if mlp else C, generator=g) / math.sqrt(HID)).bfloat16().cuda()
    b2 = torch.randn(C, generator=g).cuda()
    ga, be = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.rand(C, generator=g) - 0.5).cuda()
    master = torch.randn(M, C, generator=g).cuda()
    y = torch.empty(M, C, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        if mlp:
            K.swin_block_tail(x, w2, b2, ga, be, master, w1, b1, y=y)
        else:
            K.swin_block_tail(x, w2, b2, ga, be, master, y=y)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (EV * IDS))()
    lib.soccdpt_block_tail_trace_read(buf)
    t = [[buf[e * IDS + i] for i in range(IDS)] for e in range(EV)]
    nch = HID // 128 if mlp else 0
    print(f"== {name}  (M = {M}, C = {C}, {nch} hidden chunks per tile)")
    t0 = min(v for v in (t[10][1], t[6][1], t[0][nch] if mlp else t[10][1]) if v)
    for k in range(1, 7):
        line = f" tile {k}: LN wait_d2={t[10][k]-t0} got={t[11][k]-t0} (+{t[11][k]-t[10][k]}) stats=+{t[12][k]-t[11][k]} blocks=+{t[13][k]-t[12][k]} | period {t[13][k]-t[13][k-1]}"
        line += f" | block 1: tmem_ld +{t[15][k]-t[14][k]} math+sts +{t[16][k]-t[15][k]} cp_wait +{t[17][k]-t[16][k]} add+stores +{t[18][k]-t[17][k]} fetch_next +{t[19][k]-t[18][k]}"
        if mlp:
            line += f" | MMA a_full wait +{t[7][k]-t[6][k]}, d2_empty wait +{t[9][k]-t[8][k]}"
            for j in range(nch):
                gi = k * nch + j
                if gi < IDS:
                    line += f" | g{gi}: gelu wait_d1 +{t[1][gi]-t[0][gi]} ld +{t[2][gi]-t[1][gi]} math+st +{t[3][gi]-t[2][gi]} (at {t[3][gi]-t0}); mma wait_h +{t[5][gi]-t[4][gi]}"
        print(line)
