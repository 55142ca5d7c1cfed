"""GPU parity of the kernels only the ViT-hybrid encoder uses, each against the same operator in
plain fp32 torch on bf16-rounded inputs (the timm operators restated in oracle/timm_shim)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import cuda_ops as K  # noqa: E402


def _same_pad(x, k, s, value=0.0):
    ih, iw = x.shape[-2:]
    ph = max((math.ceil(ih / s) - 1) * s + k - ih, 0)
    pw = max((math.ceil(iw / s) - 1) * s + k - iw, 0)
    return F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2], value=value)


@pytest.mark.parametrize("hw", [(64, 64), (50, 38), (384, 384)])
def test_stem_conv7(hw):
    torch.manual_seed(0)
    x = torch.randn(2, 3, *hw, device="cuda")
    w = torch.randn(64, 3, 7, 7, device="cuda") * 0.1
    y = K.stem_conv7(x, w).float().permute(0, 3, 1, 2)
    ref = F.conv2d(_same_pad(x, 7, 2), w, None, 2)
    assert y.shape == ref.shape
    assert (y - ref).abs().max() <= 8e-3 * ref.abs().max()


@pytest.mark.parametrize("C", [64, 128, 256, 1024, 768])
@pytest.mark.parametrize("shortcut,relu", [(False, True), (True, True), (False, False)])
def test_groupnorm(C, shortcut, relu):
    torch.manual_seed(1)
    x = (torch.randn(3, 13, 11, C, device="cuda") * 2 + 0.5).bfloat16()
    g = torch.randn(C, device="cuda")
    b = torch.randn(C, device="cuda")
    sc = torch.randn(3, 13, 11, C, device="cuda").bfloat16() if shortcut else None
    y = K.groupnorm(x, g, b, sc, relu).float()
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), 32, g, b, 1e-5).permute(0, 2, 3, 1)
    if shortcut:
        ref = ref + sc.float()
    if relu:
        ref = ref.relu()
    assert (y - ref).abs().max() <= 1e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("hw", [(192, 192), (33, 17)])
def test_maxpool(hw):
    torch.manual_seed(2)
    x = torch.randn(2, *hw, 64, device="cuda").bfloat16()
    y = K.maxpool3s2(x).float().permute(0, 3, 1, 2)
    ref = F.max_pool2d(_same_pad(x.float().permute(0, 3, 1, 2), 3, 2, -float("inf")), 3, 2)
    assert torch.equal(y, ref)


def test_vit_tokens_and_readout():
    torch.manual_seed(3)
    B, L, D = 2, 576, 768
    p = torch.randn(B, L, D, device="cuda").bfloat16()
    cls = torch.randn(D, device="cuda")
    pos = torch.randn(L + 1, D, device="cuda")
    t = K.vit_tokens(p, cls, pos)
    ref = (torch.cat([cls.expand(B, 1, D), p.float()], 1) + pos).bfloat16()
    assert torch.equal(t, ref)
    f = K.readout_concat(t)
    assert torch.equal(f, torch.cat([t[:, 1:], t[:, :1].expand(B, L, D)], -1))


@pytest.mark.parametrize("C", [768, 256, 1024, 136])
def test_prenorm(C):
    torch.manual_seed(6)
    rows = 1154 + 3
    m = torch.randn(rows, C, device="cuda") * 3
    t = torch.randn(rows, C, device="cuda").bfloat16()
    g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    m0 = m.clone()
    y, sb = K.prenorm(t, m, g, b, want_stream=True)
    ref_m = m0 + t.float()
    assert torch.equal(m, ref_m) and torch.equal(sb, ref_m.bfloat16())
    ref = F.layer_norm(ref_m, (C,), g, b, 1e-6)
    assert (y.float() - ref).abs().max() <= 1e-2 * max(1.0, ref.abs().max().item())
    m2 = m.clone()
    y2, sb2 = K.prenorm(None, m2, g, b)                     # no branch: LN only, master untouched
    assert torch.equal(m2, m) and torch.equal(y2, y) and sb2 is None
    _, sb3 = K.prenorm(t, m2, None, None, want_y=False, want_stream=True)
    assert torch.equal(m2, m + t.float()) and torch.equal(sb3, m2.bfloat16())


@pytest.mark.parametrize("B,N,heads", [(2, 577, 12), (1, 100, 3), (3, 128, 2), (1, 640, 2), (2, 257, 4), (1, 7, 1),
                                       (1, 700, 2)])   # 700 > 640 tokens: CUDA-core fallback
def test_global_attention(B, N, heads):
    torch.manual_seed(4)
    qkv = (torch.randn(B, N, 3 * heads * 64, device="cuda") * 1.5).bfloat16()
    out = K.global_attention(qkv, heads).float()
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-2, -1)) * 0.125).softmax(-1) @ v
    ref = ref.transpose(1, 2).reshape(B, N, heads * 64)
    # P is rounded to bf16 for the tensor core and the output to bf16: |err| <~ 2^-8 |v|max
    assert (out - ref).abs().max() <= 2.5e-2 and (out - ref).abs().mean() <= 2e-3


@pytest.mark.parametrize("impl", ["tcgen05", "ref"])
@pytest.mark.parametrize("case", [
    # N, H, W, Cin, Cout, K, stride, pad_trim
    (2, 96, 96, 128, 128, 3, 2, 1),    # ResNetV2 stage-1 conv2 (TF SAME: 0 before, 1 after)
    (2, 48, 48, 256, 256, 3, 2, 1),    # stage-2 conv2
    (2, 96, 96, 256, 512, 1, 2, 0),    # stage-1 downsample 1x1 stride 2
    (3, 24, 24, 768, 768, 3, 2, 0),    # act_postprocess4 Conv2d(768, 768, 3, stride 2, padding 1)
    (1, 25, 19, 64, 64, 3, 2, 0),      # odd sizes
])
def test_strided_conv(case, impl):
    N, H, W, Cin, Cout, Kk, s, trim = case
    torch.manual_seed(5)
    x = torch.randn(N, H, W, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, Kk, Kk, device="cuda") / math.sqrt(Cin * Kk * Kk)).bfloat16()
    b = torch.randn(Cout, device="cuda")
    y, _, _ = K.conv(x, K.pack_conv_weight(w.float()), b, impl=impl, stride=s, pad_trim=trim)
    xin = x.float().permute(0, 3, 1, 2)
    pad = Kk // 2
    xin = F.pad(xin, [pad - trim, pad, pad - trim, pad])
    ref = F.conv2d(xin, w.float(), b, s)[:, :, : (H + s - 1) // s, : (W + s - 1) // s].permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert (y.float() - ref).abs().max() <= 2e-2 * max(1.0, ref.abs().max().item())
