"""GPU numerics of the glue kernels against plain PyTorch fp32 references of the same op
(bf16 storage -> tolerances are a couple of bf16 ulps = 2^-8 relative)."""
import math

import pytest
import torch
import torch.nn.functional as F

import cuda_ops as K
import ref_env

pytestmark = pytest.mark.gpu
BF = dict(rtol=2e-2, atol=2e-2)


def _g(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("rows,C,res", [(1000, 96, True), (77, 768, True), (64, 1536, False), (5, 2048, False)])
def test_layernorm_residual(rows, C, res):
    g = _g(1)
    t = torch.randn(rows, C, generator=g) * 2 + 0.3
    r = torch.randn(rows, C, generator=g) if res else None
    w, b = torch.rand(C, generator=g) + 0.5, torch.rand(C, generator=g) - 0.5
    tb, rb = t.bfloat16(), (r.bfloat16() if res else None)
    y = K.layernorm(tb.cuda(), rb.cuda() if res else None, w.cuda(), b.cuda())
    ref = F.layer_norm(tb.float(), (C,), w, b, 1e-5) + (rb.float() if res else 0)
    assert torch.allclose(y.float().cpu(), ref, **BF)


@pytest.mark.parametrize("rows,C,acc", [(300, 96, True), (64, 768, True), (33, 384, False)])
def test_layernorm_fp32_master_stream(rows, C, acc):
    g = _g(11)
    t = (torch.randn(rows, C, generator=g) * 2 + 0.3).bfloat16()
    m = torch.randn(rows, C, generator=g) * 5
    w, b = torch.rand(C, generator=g) + 0.5, torch.rand(C, generator=g) - 0.5
    md = m.cuda().clone()
    y = K.layernorm_master(t.cuda(), md, acc, w.cuda(), b.cuda())
    ref = F.layer_norm(t.float(), (C,), w, b, 1e-5) + (m if acc else 0)
    assert torch.allclose(md.cpu(), ref, rtol=1e-5, atol=1e-5)
    assert torch.equal(y, md.bfloat16())


# W % 64 == 0 with E in (96, 128): the tensor-core kernel (bf16 hi/lo split of image and weights, fp32 accumulate);
# other shapes: the fp32 CUDA-core kernel.  Both must reproduce the fp32 convolution + LayerNorm to fp32-level accuracy.
@pytest.mark.parametrize("B,S,E", [(2, 64, 96), (1, 32, 128), (3, 128, 96), (2, 192, 128), (1, 256, 96)])
def test_patch_embed(B, S, E):
    g = _g(2)
    x = torch.randn(B, 3, S, S, generator=g)
    w, b = torch.randn(E, 3, 4, 4, generator=g) * 0.2, torch.randn(E, generator=g) * 0.1
    lw, lb = torch.rand(E, generator=g) + 0.5, torch.rand(E, generator=g) - 0.5
    y, y32 = K.patch_embed(x.cuda(), w.reshape(E, -1).contiguous().cuda(), b.cuda(), lw.cuda(), lb.cuda(), want_f32=True)
    ref = F.layer_norm(F.conv2d(x.double(), w.double(), b.double(), stride=4).flatten(2).transpose(1, 2), (E,),
                       lw.double(), lb.double(), 1e-5)
    assert torch.allclose(y.float().cpu(), ref.float(), **BF)
    err = (y32.cpu().double() - ref).abs().max().item()
    assert err <= 2e-4, err          # fp32 master stream: the dropped lo*lo term is ~2^-17 relative per product


def test_patch_merge_gather():
    x = torch.randn(2, 8, 12, 16, generator=_g(3)).bfloat16()
    y = K.merge_gather(x.cuda())
    ref = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    assert torch.equal(y.cpu(), ref)


# exact x2 shapes run the row-marching kernel (segments of 8 row groups: 2, 3, 9, 17, 33 row groups cover one segment, a ragged
# last segment and many segments); the last case is the general kernel
@pytest.mark.parametrize("h,w,H,W,C", [(8, 8, 16, 16, 32), (16, 16, 32, 32, 32), (12, 12, 24, 24, 32), (2, 2, 4, 4, 8), (32, 20, 64, 40, 256),
                                       (7, 9, 14, 18, 64), (5, 7, 9, 20, 32)])
def test_upsample_bilinear_align_corners(h, w, H, W, C):
    x = torch.randn(3, h, w, C, generator=_g(4)).bfloat16()
    y = K.upsample(x.cuda(), H, W)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    assert torch.allclose(y.float().cpu(), ref, rtol=1e-2, atol=1e-2)
    assert (y.float().cpu() - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()      # one bf16 rounding of the result


@pytest.mark.parametrize("act", [0, 1])
def test_seg_finish(act):
    lg = torch.randn(2, 16, 16, 3, generator=_g(5)) * 3
    y = K.seg_finish(lg.cuda(), act)
    up = F.interpolate(lg.permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True)
    ref = torch.sigmoid(up) if act == 0 else 0.5 * torch.tanh(up) + 0.5
    assert torch.allclose(y.cpu(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("res,ws,target_ws,heads,shift_block,hi_scale", [
    (32, 16, 16, 3, True, False), (16, 16, 16, 12, True, False), (8, 8, 16, 24, False, False), (24, 12, 12, 4, True, False),
    (32, 8, 8, 2, False, False), (32, 8, 8, 3, True, False), (16, 8, 8, 24, True, True),
    # un-shifted 8x8 / 12x12 windows: the warp-MMA kernel (attention_small.cu); last stage of tiny (8x8 map) and base_384 (12x12 map)
    (12, 12, 24, 4, False, False), (24, 12, 12, 3, False, False), (8, 8, 16, 24, False, True), (12, 12, 24, 32, False, True),
    # 24x24 windows of swin2_base_384 (attention_tc24.cu): shifted 2x2 windows, un-shifted, one window == the whole stage
    (48, 24, 24, 4, True, False), (48, 24, 24, 2, False, False), (24, 24, 24, 3, True, False),
    # logit scales 40..100 (clamp): the exact row-max pre-pass of both tensor-core kernels
    (32, 16, 16, 3, True, True), (48, 24, 24, 2, True, True)])
def test_window_attention_vs_timm_restatement(res, ws, target_ws, heads, shift_block, hi_scale):
    """against the oracle's SwinTransformerBlock._attn minus the proj (roll, partition, cosine attention,
    cpb bias, shift mask, softmax, PV, reverse)."""
    ref_env.enable_shim()
    from timm.models.swin_transformer_v2 import SwinTransformerBlock
    from soccdpt_b200.model.encoder import relative_position_bias_table
    C = heads * 32
    torch.manual_seed(0)
    blk = SwinTransformerBlock(C, (res, res), heads, target_ws, target_ws // 2 if shift_block else 0, 4.0, 0).eval()
    g = _g(6)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.5 if p.dim() > 1 else 0.2))
        blk.attn.logit_scale.copy_(torch.rand(heads, 1, 1, generator=g) * 1.4 + 1.6)   # exp -> 5 .. 20
        if hi_scale:
            blk.attn.logit_scale.copy_(torch.rand(heads, 1, 1, generator=g) * 1.2 + 3.7)   # exp -> 40 .. 134, clamped at 100
        blk.attn.proj.weight.copy_(torch.eye(C))
        blk.attn.proj.bias.zero_()
    assert blk.window_size[0] == ws
    B = 2
    x = torch.randn(B, res * res, C, generator=g).bfloat16().float()
    a = blk.attn
    with torch.no_grad():
        qkv = F.linear(x, a.qkv.weight, torch.cat((a.q_bias, a.k_bias, a.v_bias))).bfloat16()
        # reference path on the SAME bf16-rounded qkv: temporarily make qkv an identity on a packed input
        ref = _attn_from_qkv(blk, qkv.float(), B)
        bias = relative_position_bias_table(a, ws, 0)
        scale = torch.clamp(a.logit_scale, max=math.log(100.0)).exp().reshape(-1)
    out = K.window_attention(qkv.cuda(), bias.contiguous().cuda(), scale.contiguous().cuda(), B, res, res, C,
                             heads, ws, blk.shift_size[0])
    # bf16 operands: the tcgen05 kernel rounds the normalised, scale-multiplied q (|q| up to the logit scale) and
    # the probabilities to bf16 -> logit noise ~ 2^-9 * scale; a layout bug would give O(1) relative errors.
    err = (out.float().cpu() - ref).abs()
    mx = ref.abs().max().item()
    assert torch.isfinite(out.float()).all()
    if hi_scale:    # logit noise ~ 2^-9 * 100 = 0.2 on a peaky softmax: looser, still far from a layout / underflow bug (O(1))
        assert err.max().item() <= 0.35 * mx and err.mean().item() <= 3e-2 * mx, (err.max().item(), err.mean().item(), mx)
    else:
        assert err.max().item() <= 4e-2 * mx and err.mean().item() <= 4e-3 * mx, (err.max().item(), err.mean().item(), mx)


@pytest.mark.parametrize("res,heads,shift_block,B", [
    (64, 3, True, 12), (64, 3, False, 2),       # stage 0 of swin2_tiny_256: 4x4 windows, every mask case (edge rows, columns, corner)
    (32, 6, True, 3), (32, 6, False, 1),        # stage 1: 2x2 windows, each one a different mask case
    (16, 12, True, 5),                          # stage 2: one window per frame (timm drops the shift)
    (48, 2, True, 1), (16, 3, False, 70)])      # 3x3 windows (not a power of two); more items than CTAs x stages
def test_window_attention_tma_vs_timm_restatement(res, heads, shift_block, B):
    """The TMA-fed pipelined kernel on operands normalised by the qkv GEMM's epilogue (csrc/attention_tma.cu) against the oracle's
    SwinTransformerBlock._attn, and against the round-1 kernel on the same raw qkv."""
    ref_env.enable_shim()
    from timm.models.swin_transformer_v2 import SwinTransformerBlock
    from soccdpt_b200.model.encoder import relative_position_bias_table
    C, ws = heads * 32, 16
    torch.manual_seed(0)
    blk = SwinTransformerBlock(C, (res, res), heads, ws, ws // 2 if shift_block else 0, 4.0, 0).eval()
    g = _g(16)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.5 if p.dim() > 1 else 0.2))
        blk.attn.logit_scale.copy_(torch.rand(heads, 1, 1, generator=g) * 1.4 + 1.6)   # exp -> 5 .. 20
    a = blk.attn
    x = torch.randn(B, res * res, C, generator=g).bfloat16()
    with torch.no_grad():
        wq = a.qkv.weight.bfloat16()
        bq = torch.cat((a.q_bias, a.k_bias, a.v_bias)).float()
        qkv32 = F.linear(x.float(), wq.float(), bq)                     # what the GEMM accumulates (fp32)
        ref = _attn_from_qkv(blk, qkv32, B)
        bias = relative_position_bias_table(a, ws, 0).contiguous().cuda()
        scale = torch.clamp(a.logit_scale, max=math.log(100.0)).exp().reshape(-1).contiguous()
    qscale = (scale * math.log2(math.e)).cuda()
    # qkv GEMM with the cosine-attention epilogue, on both conv kernels
    y_tc, _, _ = K.conv(x.view(1, 1, B * res * res, C).cuda(), wq.view(3 * C, 1, C).cuda().contiguous(), bias=bq.cuda(), qk=(qscale, heads))
    y_ref, _, _ = K.conv(x.view(1, 1, B * res * res, C).cuda(), wq.view(3 * C, 1, C).cuda().contiguous(), bias=bq.cuda(), qk=(qscale, heads),
                         impl="ref")
    q, k, v = qkv32.view(B, res * res, 3, heads, 32).unbind(2)
    want = torch.stack((F.normalize(q, dim=-1) * qscale.cpu().view(1, 1, heads, 1), F.normalize(k, dim=-1), v), dim=2).reshape(B * res * res, 3 * C)
    for y in (y_tc, y_ref):
        e = (y.float().cpu().view(-1, 3 * C) - want).abs()
        # one bf16 rounding of the fp32 result; the absolute term covers the accumulation-order noise of a K = C sum (|terms| ~ 10)
        assert (e <= 2.0 ** -8 * want.abs() + 1e-3).all(), (e - 2.0 ** -8 * want.abs()).max().item()
    out = K.window_attention_normed(y_tc.view(B, res * res, 3 * C), bias, scale.cuda(), B, res, res, C, heads, blk.shift_size[0])
    assert torch.isfinite(out.float()).all()
    err = (out.float().cpu() - ref).abs()
    mx = ref.abs().max().item()
    assert err.max().item() <= 4e-2 * mx and err.mean().item() <= 4e-3 * mx, (err.max().item(), err.mean().item(), mx)
    # the round-1 kernel on the raw bf16 qkv of the same GEMM: both are within bf16 noise of the reference, and of each other
    raw, _, _ = K.conv(x.view(1, 1, B * res * res, C).cuda(), wq.view(3 * C, 1, C).cuda().contiguous(), bias=bq.cuda())
    old = K.window_attention(raw.view(B, res * res, 3 * C), bias, scale.cuda(), B, res, res, C, heads, ws, blk.shift_size[0])
    d = (out.float() - old.float()).abs().cpu()
    assert d.max().item() <= 5e-2 * mx and d.mean().item() <= 5e-3 * mx, (d.max().item(), d.mean().item(), mx)


def _attn_from_qkv(blk, qkv, B):
    """SwinTransformerBlock._attn with the qkv projection already applied (qkv: (B, L, 3C))."""
    from timm.models.swin_transformer_v2 import window_partition, window_reverse
    H, W = blk.input_resolution
    C3 = qkv.shape[-1]
    C = C3 // 3
    a = blk.attn
    x = qkv.view(B, H, W, C3)
    if any(blk.shift_size):
        x = torch.roll(x, shifts=(-blk.shift_size[0], -blk.shift_size[1]), dims=(1, 2))
    xw = window_partition(x, blk.window_size).view(-1, blk.window_area, C3)
    B_, N, _ = xw.shape
    q, k, v = xw.reshape(B_, N, 3, a.num_heads, -1).permute(2, 0, 3, 1, 4).unbind(0)
    attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
    attn = attn * torch.clamp(a.logit_scale, max=math.log(100.0)).exp()
    attn = attn + a.relative_position_bias().unsqueeze(0)
    if blk.attn_mask is not None:
        nW = blk.attn_mask.shape[0]
        attn = attn.view(B_ // nW, nW, a.num_heads, N, N) + blk.attn_mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, a.num_heads, N, N)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(B_, N, C)
    o = window_reverse(o.view(-1, blk.window_size[0], blk.window_size[1], C), blk.window_size, blk.input_resolution)
    if any(blk.shift_size):
        o = torch.roll(o, shifts=blk.shift_size, dims=(1, 2))
    return o.reshape(B, H * W, C)


@pytest.mark.parametrize("N,H,W,Cin,Cout,K3", [(2, 8, 8, 64, 32, True), (1, 5, 9, 24, 16, True), (1, 1, 300, 96, 288, False)])
def test_conv_ref_kernel_vs_torch(N, H, W, Cin, Cout, K3):
    g = _g(7)
    k = 3 if K3 else 1
    x = torch.randn(N, H, W, Cin, generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)).bfloat16()
    b = torch.randn(Cout, generator=g) * 0.1
    r1 = torch.randn(N, H, W, Cout, generator=g).bfloat16()
    pw, pb = torch.randn(3, Cout, generator=g) * 0.1, torch.randn(3, generator=g)
    proj = (pw.cuda(), pb.cuda(), True) if Cout <= 256 else None
    y, yr, po = K.conv(x.cuda(), K.pack_conv_weight(w.float()).cuda(), b.cuda(), act=1, res1=r1.cuda(), want_relu=True,
                       proj=proj, impl="ref")
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=k // 2)).permute(0, 2, 3, 1) + r1.float()
    assert torch.allclose(y.float().cpu(), ref, **BF)
    assert torch.allclose(yr.float().cpu(), F.relu(ref), **BF)
    if proj is not None:
        assert torch.allclose(po.cpu(), F.relu(ref @ pw.t() + pb), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("N,h,w", [(2, 16, 16), (1, 24, 40), (1, 33, 50), (1, 2, 2), (2, 70, 17), (1, 3, 129)])
def test_depth_tail_vs_fp32_gather_of_the_same_taps(N, h, w):
    """The tail kernel alone, on a given bf16 tap tensor, against upsample -> shift -> sum in fp32 (tolerance: fp32 rounding order).
    Shapes cover ragged column segments (w % 16), strips of one row (h % 32 == 1), the 2x2 minimum and several strips."""
    g = _g(21)
    T = torch.randn(N, h, w, 288, generator=g).bfloat16()
    b2 = torch.randn(32, generator=g) * 0.1
    pw, pb = torch.randn(32, generator=g) * 0.2, torch.randn(1, generator=g) * 0.1
    out = K.depth_tail(T.cuda(), b2.cuda(), pw.cuda(), pb.cuda()).cpu()
    up = F.interpolate(T.double().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True)   # (N, 288, 2h, 2w)
    up = F.pad(up, (1, 1, 1, 1))
    acc = b2.double().view(1, 32, 1, 1).expand(N, 32, 2 * h, 2 * w).clone()
    for dy in range(3):
        for dx in range(3):
            t = dy * 3 + dx
            acc += up[:, t * 32:(t + 1) * 32, dy:dy + 2 * h, dx:dx + 2 * w]
    ref = F.relu((F.relu(acc) * pw.double().view(1, 32, 1, 1)).sum(1) + pb.double()).float()
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item()
    assert err <= 1e-4 * max(ref.abs().max().item(), 1.0), err


@pytest.mark.parametrize("N,h,w", [(2, 16, 16), (1, 24, 40)])
def test_depth_head_tail_equals_upsample_conv_relu_proj(N, h, w):
    """dpt.py:209-219 restructured: tap matrices at low resolution (tcgen05 GEMM) + gather == the reference order."""
    g = _g(12)
    d0 = (torch.randn(N, h, w, 128, generator=g)).bfloat16()
    w2 = torch.randn(32, 128, 3, 3, generator=g) / math.sqrt(128 * 9)
    b2 = torch.randn(32, generator=g) * 0.1
    pw, pb = torch.randn(1, 32, generator=g) * 0.2, torch.randn(1, generator=g) * 0.1
    w2t = K.pack_conv_weight(w2.permute(2, 3, 0, 1).reshape(288, 128, 1, 1)).cuda()
    T = K.conv(d0.cuda(), w2t)[0]
    out = K.depth_tail(T, b2.cuda(), pw.reshape(-1).contiguous().cuda(), pb.cuda())
    up = F.interpolate(d0.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True)
    ref = F.relu(F.conv2d(F.relu(F.conv2d(up, w2.bfloat16().float(), b2, padding=1)), pw.view(1, 32, 1, 1), pb)).squeeze(1)
    err = (out.cpu() - ref).abs()
    assert err.max().item() <= 2e-2 * ref.abs().max().item() + 1e-3, (err.max().item(), ref.abs().max().item())
