"""The tcgen05/TMEM/TMA implicit-GEMM kernel against (a) a PyTorch fp32 reference of the same op and
(b) the CUDA-core reference kernel, over every tile geometry the network uses (incl. M/K tails)."""
import math

import pytest
import torch
import torch.nn.functional as F

import cuda_ops as K

pytestmark = pytest.mark.gpu

# (N, H, W, Cin, Cout, k, act, bias, nres, relu_out, proj_n)
CASES = [
    (1, 1, 256, 64, 64, 1, 0, False, 0, False, 0),       # smallest GEMM, 2 full tiles
    (1, 1, 4096, 96, 288, 1, 0, True, 0, False, 0),      # S0 qkv: K tail (96 = 64 + 32), N = 2 x 144
    (1, 1, 64, 768, 2304, 1, 0, True, 0, False, 0),      # S3 qkv at B=1: M tail (64 < 128)
    (1, 1, 200, 3072, 768, 1, 0, True, 0, False, 0),     # fc2, M tail, long K
    (1, 1, 512, 384, 1536, 1, 2, True, 0, False, 0),     # fc1 + GELU
    (1, 1, 128, 1536, 768, 1, 0, False, 0, False, 0),    # patch-merge reduction (no bias)
    (2, 8, 8, 768, 256, 3, 0, False, 0, True, 0),        # layer4_rn: 2 images per tile
    (3, 8, 8, 256, 256, 3, 1, True, 0, False, 0),        # 8x8 level, odd batch
    (2, 16, 16, 256, 256, 3, 0, True, 2, True, 0),       # RCU conv2 with two residuals + relu copy
    (2, 32, 32, 256, 256, 3, 1, True, 0, False, 0),
    (1, 64, 64, 96, 256, 3, 0, False, 0, True, 0),       # layer1_rn: Cin = 96
    (1, 64, 64, 256, 256, 1, 0, True, 0, False, 0),      # out_conv 1x1
    (1, 128, 128, 256, 128, 3, 0, True, 0, False, 0),    # depth head conv 0
    (1, 256, 256, 128, 32, 3, 1, True, 0, False, 1),     # depth head conv 2 + fused 32->1
    (1, 128, 128, 256, 256, 3, 1, True, 0, False, 3),    # seg head + fused 256->3
    (2, 12, 12, 1024, 256, 3, 0, False, 0, False, 0),    # swin2_base_384 levels (non power of two)
    (1, 24, 24, 256, 256, 3, 1, True, 1, False, 0),
    (1, 48, 48, 256, 256, 3, 0, True, 0, False, 0),
    (1, 96, 96, 128, 256, 3, 0, False, 0, False, 0),
    (1, 192, 192, 256, 128, 3, 0, True, 0, False, 0),
    (1, 7, 150, 64, 48, 3, 0, True, 0, False, 0),        # W tail inside a row (150 = 128 + 22)
    # >= 2 x 148 tiles: the RESIDENT N BLOCK path of the 1x1 / linear layers (weights of a CTA's N block fetched once,
    # activation-only ring) -- the bench-size shapes, which the small cases above never reach
    (1, 1, 40000, 96, 288, 1, 0, True, 0, False, 0),     # S0 qkv: 313 M tiles (M tail 64) x 2 N blocks, K tail, 8-slot A ring
    (1, 1, 8192, 384, 1152, 1, 0, True, 0, False, 0),    # S2 qkv: 6 N blocks (grid 144), 6 k blocks through a 3-slot ring
    (1, 1, 20000, 192, 768, 1, 2, True, 0, False, 0),    # S1 fc1 + GELU: 3 N blocks (grid 147)
    (1, 1, 8192, 384, 1536, 1, 2, True, 0, False, 0),    # S2 fc1 + GELU: 6 N blocks of 256, too large to stay resident (plain ring)
    (1, 1, 40000, 384, 96, 1, 0, True, 1, False, 0),     # S0 fc2 + residual: one N block of 96
    (12, 64, 64, 256, 256, 1, 0, True, 0, True, 0),      # out_conv 1x1 (NHWC boxes 64 x 2) + ReLU copy
    (3, 128, 128, 128, 288, 1, 0, False, 0, False, 0),   # depth-head tap GEMM: 2 N blocks of 144
    (5, 128, 128, 256, 256, 1, 1, True, 0, False, 3),    # fused projection epilogue on the resident path
    # ROW-PAIR halo path (3x3, W % 128 == 0, N <= 128, even H): two M tiles per weight tile
    (2, 4, 128, 64, 128, 3, 0, True, 0, False, 0),       # one channel block, pad rows at both image edges inside one box
    (1, 6, 256, 96, 64, 3, 1, True, 2, True, 0),         # K tail (96), N = 64 (second accumulator at TMEM column 128), residuals + ReLU copy
    (3, 128, 128, 256, 128, 3, 0, True, 0, False, 0),    # depth head conv 0 at > 148 pair tiles (persistent loop, accumulator ring)
    (1, 5, 128, 128, 128, 3, 0, True, 0, False, 0),      # odd H: falls back to single-row halo tiles
]


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,act,bias,nres,relu_out,proj_n", CASES)
def test_tcgen05_conv(N, H, W, Cin, Cout, k, act, bias, nres, relu_out, proj_n):
    g = torch.Generator().manual_seed(N * 1000 + H * 10 + Cin + Cout)
    x = torch.randn(N, H, W, Cin, generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).bfloat16()
    b = (torch.randn(Cout, generator=g) * 0.2) if bias else None
    rs = [torch.randn(N, H, W, Cout, generator=g).bfloat16() for _ in range(nres)]
    proj = None
    if proj_n:
        proj = ((torch.randn(proj_n, Cout, generator=g) * 0.1).cuda(), torch.randn(proj_n, generator=g).cuda(), proj_n == 1)
    args = dict(bias=b.cuda() if bias else None, act=act, res1=rs[0].cuda() if nres > 0 else None,
                res2=rs[1].cuda() if nres > 1 else None, want_y=proj_n == 0, want_relu=relu_out, proj=proj)
    wp = K.pack_conv_weight(w.float()).cuda()
    y, yr, po = K.conv(x.cuda(), wp, impl="tcgen05", **args)
    y2, yr2, po2 = K.conv(x.cuda(), wp, impl="ref", **args)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=k // 2).permute(0, 2, 3, 1)
    ref = F.relu(ref) if act == 1 else (F.gelu(ref) if act == 2 else ref)
    for r in rs:
        ref = ref + r.float()
    tol = dict(rtol=2e-2, atol=2e-2)
    if y is not None:
        assert torch.allclose(y.float().cpu(), ref, **tol), (y.float().cpu() - ref).abs().max()
        assert torch.allclose(y.float(), y2.float(), rtol=1e-2, atol=1e-2)
    if yr is not None:
        assert torch.allclose(yr.float().cpu(), F.relu(ref), **tol)
    if po is not None:
        pref = ref @ proj[0].cpu().t() + proj[1].cpu()
        pref = F.relu(pref) if proj[2] else pref
        assert torch.allclose(po.cpu(), pref, rtol=1e-3, atol=2e-3), (po.cpu() - pref).abs().max()
        assert torch.allclose(po, po2, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("N,H,W,C,k,relu_out", [(2, 16, 16, 256, 3, True), (1, 32, 32, 256, 3, True), (3, 64, 64, 256, 3, False),
                                                (1, 2, 2, 64, 1, False), (1, 24, 48, 128, 3, True), (1, 128, 128, 128, 3, False)])
def test_upsampled_residual_in_the_epilogue(N, H, W, C, k, relu_out):
    """soccdpt_conv_t.up_src: y = conv(x) + bias + res1 + bilinear_x2(up_src) (FeatureFusionBlock_custom.forward, blocks.py:476-487,
    without writing the up-sampled tensor) == the same with F.interpolate(scale_factor=2, align_corners=True) as a plain residual;
    the last shape takes the row-pair halo tiles."""
    g = torch.Generator().manual_seed(H * 7 + C)
    x = torch.randn(N, H, W, C, generator=g).bfloat16()
    w = (torch.randn(C, C, k, k, generator=g) / math.sqrt(C * k * k)).bfloat16()
    b = torch.randn(C, generator=g) * 0.2
    r1 = torch.randn(N, H, W, C, generator=g).bfloat16()
    low = torch.randn(N, H // 2, W // 2, C, generator=g).bfloat16()
    wp = K.pack_conv_weight(w.float()).cuda()
    args = dict(bias=b.cuda(), res1=r1.cuda(), up=low.cuda(), want_relu=relu_out)
    y, yr, _ = K.conv(x.cuda(), wp, impl="tcgen05", **args)
    y2, yr2, _ = K.conv(x.cuda(), wp, impl="ref", **args)
    torch.cuda.synchronize()
    up = F.interpolate(low.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=k // 2).permute(0, 2, 3, 1) + r1.float() + up
    tol = dict(rtol=2e-2, atol=2e-2)
    assert torch.allclose(y.float().cpu(), ref, **tol), (y.float().cpu() - ref).abs().max()
    assert torch.allclose(y.float(), y2.float(), rtol=1e-2, atol=1e-2)
    if relu_out:
        assert torch.allclose(yr.float().cpu(), F.relu(ref), **tol)
        assert torch.allclose(yr.float(), yr2.float(), rtol=1e-2, atol=1e-2)
    # the interpolation itself, isolated: zero weights and bias leave res1 + up
    z, _, _ = K.conv(x.cuda(), torch.zeros_like(wp), impl="tcgen05", res1=r1.cuda(), up=low.cuda())
    d = (z.float().cpu() - (r1.float() + up)).abs()
    assert d.max().item() <= 2 ** -7 * (r1.float() + up).abs().max().item() + 1e-6      # one bf16 rounding of the sum


def test_tcgen05_conv_is_deterministic_and_reentrant():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 32, 32, 256, generator=g).bfloat16().cuda()
    w = K.pack_conv_weight(torch.randn(256, 256, 3, 3, generator=g) / 48).cuda()
    a = K.conv(x, w)[0]
    for _ in range(3):
        assert torch.equal(K.conv(x, w)[0], a)


@pytest.mark.parametrize("B,Hs,Ws,C", [(2, 64, 64, 96), (3, 16, 16, 384), (1, 96, 96, 128), (2, 24, 24, 512)])
def test_patch_merging_as_2x2_stride2_conv(B, Hs, Ws, C):
    """timm PatchMerging (cat of the four parity slices -> Linear(4C, 2C, bias=False)) evaluated as a 2x2 stride-2 implicit GEMM:
    the gather lives in the TMA box (element strides 2) -- against the torch restatement and the CUDA-core kernel."""
    g = torch.Generator().manual_seed(B * 100 + C)
    x = torch.randn(B, Hs, Ws, C, generator=g).bfloat16()
    rw = (torch.randn(2 * C, 4 * C, generator=g) / math.sqrt(4 * C)).bfloat16()
    xf = x.float()
    cat = torch.cat([xf[:, 0::2, 0::2], xf[:, 1::2, 0::2], xf[:, 0::2, 1::2], xf[:, 1::2, 1::2]], -1)      # timm order
    ref = cat @ rw.float().t()
    w_taps = torch.stack([rw[:, q * C:(q + 1) * C] for q in (0, 2, 1, 3)], dim=1).contiguous().cuda()      # (2C, 4, C)
    y, _, _ = K.conv(x.cuda(), w_taps, impl="tcgen05", stride=2, pad_trim=1)
    y2, _, _ = K.conv(x.cuda(), w_taps, impl="ref", stride=2, pad_trim=1)
    assert y.shape == (B, Hs // 2, Ws // 2, 2 * C)
    assert torch.allclose(y.float().cpu(), ref, rtol=2e-2, atol=2e-2), (y.float().cpu() - ref).abs().max()
    assert torch.allclose(y.float(), y2.float(), rtol=1e-2, atol=1e-2)
