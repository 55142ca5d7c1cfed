"""Timeline of one CTA of the TMA-fed window attention kernel (debug build with -DSOCCDPT_ATTN_TRACE, see csrc/attention_tma.cu):
    SOCCDPT_NVCC_FLAGS=-DSOCCDPT_ATTN_TRACE python tools/build_variant.py trace      (writes build/variants/trace/lib.so)
    SOCCDPT_LIB=build/variants/trace/lib.so python tools/trace_attention.py [S0|S1|S2] [shift]
Prints, per work unit, the SM-clock intervals of the hand-offs between the MMA issuer and the softmax groups."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.nn.functional as F
import cuda_ops as K
from soccdpt_b200 import _cabi

stage = sys.argv[1] if len(sys.argv) > 1 else "S0"
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 0
res, C, heads = {"S0": (64, 96, 3), "S1": (32, 192, 6), "S2": (16, 384, 12)}[stage]
B = 64
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, res * res, 3 * C, generator=g)
bias = (torch.rand(heads, 961, generator=g) * 16).cuda()
scale = (torch.rand(heads, generator=g) * 15 + 5).cuda()
q, k, v = qkv.view(B, res * res, 3, heads, 32).unbind(2)
qkvn = torch.stack((F.normalize(q, dim=-1) * (scale.cpu() * 1.4426950408889634).view(1, 1, heads, 1), F.normalize(k, dim=-1), v),
                   dim=2).reshape(B, res * res, 3 * C).bfloat16().cuda()
lib = _cabi.load()
fn = lib.soccdpt_debug_attention_trace
fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p]
for _ in range(3):
    K.window_attention_normed(qkvn, bias, scale, B, res, res, C, heads, shift)
torch.cuda.synchronize()
KINDS, IDS = 11, 384
buf = np.zeros(KINDS * IDS, dtype=np.uint32)
assert fn(buf.ctypes.data) == 0
tr = buf.reshape(KINDS, IDS).astype(np.int64)
t0 = tr[tr > 0].min()
ev = [(k_, i, (tr[k_, i] - t0) & 0xFFFFFFFF) for k_ in range(KINDS) for i in range(IDS) if tr[k_, i] > 0]
n = len(ev)
ev = np.array(ev, dtype=np.int64)
names = {1: "tma_req", 2: "mma_full_wait", 3: "mma_full_ok", 4: "S_issued", 5: "P_seen", 6: "PV_issued", 7: "sm_wait", 8: "sm_go",
         9: "sm_done", 10: "O_seen"}
by = {}
for kind, idx, t in ev:
    by[(int(kind), int(idx))] = int(t)
U = max(i for (k_, i) in by if k_ == 9) + 1
print(f"{stage} shift {shift}: {n} events, {U} units in CTA 0, total {ev[:, 2].max()} cycles")
print("unit  S_issued  sm_wait   sm_go  sm_done |  wait(go-wait)  softmax(done-go)  S->go  done->P_seen  P_seen->PV_issued  PV_issued->S(u+6)  S(u+6)->go(u+6)")
tot = dict(wait=0, soft=0)
for u in range(U):
    g_ = lambda k_, i=u: by.get((k_, i), -1)
    w, go, dn, si, ps, pv = g_(7), g_(8), g_(9), g_(4), g_(5), g_(6)
    s3, go3 = g_(4, u + 6), g_(8, u + 6)
    tot["wait"] += go - w
    tot["soft"] += dn - go
    if u < 24 or u > U - 6:
        print(f"{u:4d} {si:9d} {w:8d} {go:7d} {dn:8d} | {go - w:8d} {dn - go:12d} {go - si:12d} {ps - dn:10d} {pv - ps:14d} {s3 - pv if s3 >= 0 else -1:16d} {go3 - s3 if go3 >= 0 else -1:14d}")
print(f"mean wait {tot['wait'] / U:.0f} cycles, mean softmax {tot['soft'] / U:.0f} cycles per unit")
for (k_, i), t in sorted(by.items()):
    if k_ in (1, 2, 3) and i < 8:
        print(names[k_], i, t)
