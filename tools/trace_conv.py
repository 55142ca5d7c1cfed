"""Hand-off trace of the conv kernel's CTA 0 (trace build only):
    SOCCDPT_NVCC_FLAGS=-DSOCCDPT_CONV_TRACE python tools/build_variant.py trace
    SOCCDPT_LIB=build/variants/trace/lib.so python tools/trace_conv.py
Events per tile: 0/1 MMA warp before/after the acc_empty wait, 2 first k block issued, 3 accumulator committed; 8/9 epilogue thread 0
before/after the acc_full wait, then per 32-column chunk: start, TMEM load done, math done, staging tile free (store warp's barrier),
staged + proxy fence + arrive; 30 accumulator handed back."""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuda_ops as K
from soccdpt_b200 import _cabi

lib = ctypes.CDLL(_cabi.lib_path())
T, E = 24, 32
names = {0: "mma:wait_empty", 1: "mma:got_empty", 2: "mma:kb0_issued", 3: "mma:committed", 8: "epi:wait_full", 9: "epi:got_full", 30: "epi:released"}
chunk_ev = ["start", "tmem_ld", "math", "tile_free", "staged+fence+arrive"]
NE = len(chunk_ev)
B = 64
SHAPES = [("S0 qkv 96->288 (qk epilogue, N blocks of 96)", B * 4096, 96, 288, 0, 3),
          ("S2 qkv 384->1152 (qk epilogue)", B * 256, 384, 1152, 0, 12),
          ("S2 fc1 384->1536 gelu", B * 256, 384, 1536, 2, 0),
          ("out_conv 256->256 @64^2", B * 4096, 256, 256, 0, 0)]
g = torch.Generator().manual_seed(0)
for name, M, Cin, Cout, act, heads in SHAPES:
    x = torch.randn(1, 1, M, Cin, generator=g).bfloat16().cuda()
    w = K.pack_conv_weight(torch.randn(Cout, Cin, 1, 1, generator=g) / math.sqrt(Cin)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    qk = (torch.ones(heads).cuda(), heads) if heads else None
    for _ in range(3):
        K.conv(x, w, bias=b, act=act, qk=qk)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (T * E))()
    lib.soccdpt_conv_trace_read(buf)
    tr = [[buf[t * E + e] for e in range(E)] for t in range(T)]
    print(f"== {name}")
    t0 = tr[2][0]
    for t in range(2, 10):
        r = tr[t]
        if r[9] == 0:
            break
        mma = " ".join(f"{names[e]}={r[e] - t0}" for e in (0, 1, 2, 3))
        epi = f"epi:wait_full={r[8] - t0} got_full={r[9] - t0}"
        chunks = []
        for c in range(3):
            base = 10 + NE * c
            if r[base] == 0 or r[base] < r[9]:
                break
            chunks.append(" ".join(f"{chunk_ev[i]}+{r[base + i] - (r[base + i - 1] if i else r[9] if c == 0 else r[base - 1])}" for i in range(NE)))
        print(f" tile {t}: {mma} | {epi} | " + " || ".join(chunks) + f" | released={r[30] - t0} (tile period {r[30] - tr[t - 1][30]})")
