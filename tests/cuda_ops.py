"""Thin tensor-level wrappers over the C-ABI used by the GPU tests (pointers in, pointers out)."""
import ctypes

import torch

from soccdpt_b200 import _cabi


def _s():
    return _cabi.current_stream()


def conv(x, w, bias=None, act=0, res1=None, res2=None, want_y=True, want_relu=False, proj=None, impl="tcgen05",
         stride=1, pad_trim=0, qk=None, up=None):
    """x (N,H,W,Cin) bf16; w (Cout,KH*KW,Cin) bf16 packed. Returns (y, y_relu, proj_out)."""
    lib = _cabi.load()
    N, H, W, Cin = x.shape
    Cout, taps, _ = w.shape
    K = {9: 3, 4: 2, 1: 1}[taps]
    c = _cabi.Conv()
    c.x, c.wgt = x.data_ptr(), w.data_ptr()
    c.bias = bias.data_ptr() if bias is not None else None
    c.res1 = res1.data_ptr() if res1 is not None else None
    c.res2 = res2.data_ptr() if res2 is not None else None
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    c.stride, c.pad_trim = stride, pad_trim
    y = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device) if want_y else None
    yr = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device) if want_relu else None
    c.y = y.data_ptr() if y is not None else None
    c.y_relu = yr.data_ptr() if yr is not None else None
    c.N, c.H, c.W, c.Cin, c.Cout, c.KH, c.KW, c.act = N, H, W, Cin, Cout, K, K, act
    po = None
    if proj is not None:
        pw, pb, relu = proj
        po = torch.empty((N, Ho, Wo, pw.shape[0]), dtype=torch.float32, device=x.device)
        c.proj_w, c.proj_b, c.proj_out, c.proj_n, c.proj_relu = pw.data_ptr(), pb.data_ptr(), po.data_ptr(), pw.shape[0], int(relu)
    if qk is not None:          # (qk_scale f32 [heads], heads): cosine-attention epilogue of a qkv linear
        c.qk_scale, c.qk_heads = qk[0].data_ptr(), qk[1]
    if up is not None:          # (N, H/2, W/2, Cout) bf16: residual added through a bilinear x2 (align_corners=True) in the epilogue
        c.up_src, c.up_h, c.up_w = up.data_ptr(), up.shape[1], up.shape[2]
    fn = lib.soccdpt_conv_fwd if impl == "tcgen05" else lib.soccdpt_conv_ref_fwd
    _cabi.check(fn(ctypes.byref(c), _s()), "conv")
    return y, yr, po


def pack_conv_weight(w):
    """(Cout,Cin,KH,KW) f32 -> (Cout,KH*KW,Cin) bf16"""
    Cout, Cin, KH, KW = w.shape
    return w.permute(0, 2, 3, 1).reshape(Cout, KH * KW, Cin).to(torch.bfloat16).contiguous()


def layernorm(t, res, g, b, eps=1e-5):
    lib = _cabi.load()
    y = torch.empty_like(t)
    rows, C = t.shape
    _cabi.check(lib.soccdpt_layernorm_fwd(t.data_ptr(), res.data_ptr() if res is not None else None, g.data_ptr(),
                                          b.data_ptr(), y.data_ptr(), rows, C, eps, _s()), "layernorm")
    return y


def patch_embed(x, w, b, g, be, want_f32=False):
    lib = _cabi.load()
    B, _, H, W = x.shape
    E = w.shape[0]
    out = torch.empty((B, (H // 4) * (W // 4), E), dtype=torch.bfloat16, device=x.device)
    out32 = torch.empty((B, (H // 4) * (W // 4), E), dtype=torch.float32, device=x.device)
    _cabi.check(lib.soccdpt_patch_embed_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), g.data_ptr(), be.data_ptr(),
                                            out.data_ptr(), out32.data_ptr(), B, H, W, E, _s()), "patch_embed")
    assert torch.equal(out32.bfloat16(), out)
    return (out, out32) if want_f32 else out


def layernorm_master(t, master, accumulate, g, b, eps=1e-5):
    lib = _cabi.load()
    y = torch.empty_like(t)
    rows, C = t.shape
    _cabi.check(lib.soccdpt_layernorm_master_fwd(t.data_ptr(), master.data_ptr(), int(accumulate), g.data_ptr(), b.data_ptr(),
                                                 y.data_ptr(), rows, C, eps, _s()), "layernorm_master")
    return y


def merge_gather(x):
    lib = _cabi.load()
    B, H, W, C = x.shape
    y = torch.empty((B, H // 2, W // 2, 4 * C), dtype=torch.bfloat16, device=x.device)
    _cabi.check(lib.soccdpt_patch_merge_gather_fwd(x.data_ptr(), y.data_ptr(), B, H, W, C, _s()), "merge")
    return y


def upsample(x, H, W):
    lib = _cabi.load()
    N, h, w, C = x.shape
    y = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=x.device)
    _cabi.check(lib.soccdpt_upsample_bilinear_fwd(x.data_ptr(), y.data_ptr(), N, h, w, H, W, C, _s()), "upsample")
    return y


def seg_finish(logits, act):
    lib = _cabi.load()
    N, h, w, P = logits.shape
    y = torch.empty((N, P, 2 * h, 2 * w), dtype=torch.float32, device=logits.device)
    _cabi.check(lib.soccdpt_seg_finish_fwd(logits.data_ptr(), y.data_ptr(), N, h, w, P, act, _s()), "seg_finish")
    return y


def window_attention(qkv, biasT, scale, B, Hs, Ws, C, heads, ws, shift):
    lib = _cabi.load()
    out = torch.empty((B, Hs * Ws, C), dtype=torch.bfloat16, device=qkv.device)
    _cabi.check(lib.soccdpt_window_attention_fwd(qkv.data_ptr(), biasT.data_ptr(), scale.data_ptr(), out.data_ptr(), B, Hs,
                                                 Ws, C, heads, ws, shift, _s()), "window_attention")
    return out


def window_attention_normed(qkvn, biasT, scale, B, Hs, Ws, C, heads, shift):
    lib = _cabi.load()
    out = torch.empty((B, Hs * Ws, C), dtype=torch.bfloat16, device=qkvn.device)
    _cabi.check(lib.soccdpt_window_attention_normed_fwd(qkvn.data_ptr(), biasT.data_ptr(), scale.data_ptr(), out.data_ptr(), B, Hs,
                                                        Ws, C, heads, shift, _s()), "window_attention_normed")
    return out


def depth_tail(T, b2, pw, pb):
    lib = _cabi.load()
    N, h, w, _ = T.shape
    out = torch.empty((N, 2 * h, 2 * w), dtype=torch.float32, device=T.device)
    _cabi.check(lib.soccdpt_depth_tail_fwd(T.data_ptr(), b2.data_ptr(), pw.data_ptr(), pb.data_ptr(), out.data_ptr(), N, h, w,
                                           _s()), "depth_tail")
    return out


# ---- ViT-hybrid encoder kernels
def stem_conv7(x, w):
    lib = _cabi.load()
    B, _, H, W = x.shape
    y = torch.empty((B, (H + 1) // 2, (W + 1) // 2, 64), dtype=torch.bfloat16, device=x.device)
    _cabi.check(lib.soccdpt_stem_conv7_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), B, H, W, _s()), "stem_conv7")
    return y


def groupnorm(x, gamma, beta, shortcut=None, relu=True, eps=1e-5):
    lib = _cabi.load()
    B, H, W, C = x.shape
    y = torch.empty_like(x)
    scratch = torch.empty(B * 64, dtype=torch.float64, device=x.device)
    _cabi.check(lib.soccdpt_groupnorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                          shortcut.data_ptr() if shortcut is not None else None, y.data_ptr(), B, H * W, C,
                                          eps, int(relu), scratch.data_ptr(), _s()), "groupnorm")
    return y


def maxpool3s2(x):
    lib = _cabi.load()
    B, H, W, C = x.shape
    y = torch.empty((B, (H + 1) // 2, (W + 1) // 2, C), dtype=torch.bfloat16, device=x.device)
    _cabi.check(lib.soccdpt_maxpool3s2_fwd(x.data_ptr(), y.data_ptr(), B, H, W, C, _s()), "maxpool")
    return y


def vit_tokens(patches, cls, pos):
    lib = _cabi.load()
    B, L, D = patches.shape
    t = torch.empty((B, L + 1, D), dtype=torch.bfloat16, device=patches.device)
    t32 = torch.empty((B, L + 1, D), dtype=torch.float32, device=patches.device)
    _cabi.check(lib.soccdpt_vit_tokens_fwd(patches.data_ptr(), cls.data_ptr(), pos.data_ptr(), t.data_ptr(), t32.data_ptr(),
                                           B, L, D, _s()), "vit_tokens")
    assert torch.equal(t32.bfloat16(), t)
    return t


def prenorm(t, master, gamma, beta, want_y=True, want_stream=False, eps=1e-6):
    """master (rows,C) f32 updated in place; returns (y, stream_bf16)."""
    lib = _cabi.load()
    rows, C = master.shape
    y = torch.empty((rows, C), dtype=torch.bfloat16, device=master.device) if want_y else None
    sb = torch.empty((rows, C), dtype=torch.bfloat16, device=master.device) if want_stream else None
    _cabi.check(lib.soccdpt_prenorm_fwd(t.data_ptr() if t is not None else None, master.data_ptr(),
                                        gamma.data_ptr() if gamma is not None else None,
                                        beta.data_ptr() if beta is not None else None,
                                        y.data_ptr() if y is not None else None, sb.data_ptr() if sb is not None else None,
                                        rows, C, eps, _s()), "prenorm")
    return y, sb


def readout_concat(tokens):
    lib = _cabi.load()
    B, L1, D = tokens.shape
    f = torch.empty((B, L1 - 1, 2 * D), dtype=torch.bfloat16, device=tokens.device)
    _cabi.check(lib.soccdpt_readout_concat_fwd(tokens.data_ptr(), f.data_ptr(), B, L1 - 1, D, _s()), "readout_concat")
    return f


def global_attention(qkv, heads):
    lib = _cabi.load()
    B, N, C3 = qkv.shape
    out = torch.empty((B, N, C3 // 3), dtype=torch.bfloat16, device=qkv.device)
    _cabi.check(lib.soccdpt_global_attention_fwd(qkv.data_ptr(), out.data_ptr(), B, N, heads, C3 // 3 // heads, _s()), "global_attention")
    return out


def swin_block_tail(x, w2, b2, gamma, beta, master, w1=None, b1=None, eps=1e-5, y=None):
    """x (M,K1) bf16; w1 (HID,K1) bf16 | None; w2 (C,HID) | (C,K1) bf16; master (M,C) f32 updated in place. Returns y bf16."""
    lib = _cabi.load()
    M, K1 = x.shape
    C = w2.shape[0]
    a = _cabi.BlockTail()
    a.x, a.w2, a.b2, a.gamma, a.beta = x.data_ptr(), w2.data_ptr(), b2.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    a.w1 = w1.data_ptr() if w1 is not None else None
    a.b1 = b1.data_ptr() if b1 is not None else None
    if y is None:
        y = torch.empty((M, C), dtype=torch.bfloat16, device=x.device)
    a.master, a.y = master.data_ptr(), y.data_ptr()
    a.M, a.K1, a.HID, a.C, a.eps = M, K1, (w1.shape[0] if w1 is not None else 0), C, eps
    _cabi.check(lib.soccdpt_swin_block_tail_fwd(ctypes.byref(a), _s()), "swin_block_tail")
    return y
