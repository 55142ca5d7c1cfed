"""Execution engine: packs the weights of a DPT module tree once (bf16, BN folded, attention bias tables baked) and
replays the image -> (inverse depth, segmentation) network as a fixed list of C-ABI kernel launches on the current CUDA
stream.  PyTorch only owns the device buffers.  One engine per network: SOccDPT_V3 has one (shared trunk, both heads),
SOccDPT_V1 two (depth DPT; segmentation DPT with BatchNorm in its residual conv units).

Kernel schedule per frame batch (reference call sites in include/soccdpt_b200.h):
  patch_embed -> for every Swin block: qkv GEMM, window attention, proj GEMM, LN+residual,
  fc1 GEMM (+GELU), fc2 GEMM, LN+residual -> patch-merge gather + GEMM + LN between stages ->
  decoder: layerN_rn conv3x3, residual conv units (bias/ReLU/residual fused in the GEMM epilogue),
  out_conv 1x1 evaluated BEFORE the bilinear upsample (they commute, 4x fewer FLOPs) ->
  depth head (32->1 projection fused in the epilogue) and seg head (BN folded, 256->3 fused).
ViT-hybrid encoder (dpt_hybrid_384) instead of the Swin part:
  stem conv7x7/2 -> GroupNorm+ReLU -> max-pool -> 16 ResNetV2 bottlenecks (weight-standardised convs as implicit GEMMs,
  GroupNorm (+shortcut) (+ReLU) kernels) -> 1x1 patch projection -> cls/pos tokens -> 12 pre-norm ViT blocks
  (fp32 residual stream, global attention) -> ProjectReadout GEMM (+GELU) + 1x1 / stride-2 3x3 convs on blocks 8 and 11.
"""
import ctypes
import math

import os

import torch

from . import _cabi
from .model.encoder import SwinV2Params, relative_position_bias_table


class _Launch:
    """One enqueued C-ABI call with its arguments frozen (pointers stay valid: the plan owns the buffers)."""
    __slots__ = ("fn", "args", "name")

    def __init__(self, name, fn, *args):
        self.name, self.fn, self.args = name, fn, args

    def __call__(self, stream):
        rc = self.fn(*self.args, stream)
        if rc != 0:
            _cabi.check(rc, self.name)


def _bf16(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).to(torch.bfloat16).contiguous()


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _pack_conv(w, dev):
    """(Cout, Cin, KH, KW) fp32 -> [Cout][KH*KW][Cin] bf16."""
    return _bf16(w.detach().float().permute(0, 2, 3, 1), dev)


class NetworkEngine:
    def __init__(self, net, conv_impl="tcgen05", dpt=None, seg_head=None, heads=("depth", "seg")):
        """``dpt``: the DPT parameter container to run (default ``net.depth_net``); ``seg_head``: the segmentation head
        Sequential (default ``net.seg_head``; SOccDPT_V1's segmentation network passes its own ``scratch.output_conv``);
        ``heads``: which outputs the plan produces -- ("depth", "seg") for SOccDPT_V3's shared trunk, ("depth",) / ("seg",)
        for the two networks of SOccDPT_V1 (SOccDPT.py:470-523)."""
        self.net = net
        self.dpt = dpt if dpt is not None else net.depth_net
        self.heads = tuple(heads)
        assert self.heads and set(self.heads) <= {"depth", "seg"}
        self.seg_head = seg_head if seg_head is not None else (getattr(net, "seg_head", None) if "seg" in self.heads else None)
        self.conv_impl = conv_impl        # "tcgen05" (product) | "ref" (CUDA-core cross-check, tests only)
        self._weights = None
        self._plans = {}                 # (batch, device) -> plan, least recently used first; at most `max_plans` are kept
        self.max_plans = int(os.environ.get("SOCCDPT_MAX_PLANS", "4"))
        self.lib = _cabi.load()
        self.use_graphs = os.environ.get("SOCCDPT_CUDA_GRAPH", "0") == "1"
        # parity mode of dpt_hybrid_384: ResNetV2 trunk with fp32 storage and fp32 CUDA-core convolutions (csrc/trunk_fp32.cu)
        self.trunk_fp32 = os.environ.get("SOCCDPT_HYBRID_TRUNK", "") == "fp32"

    def set_trunk_precision(self, precision):
        """"bf16" (product path) or "fp32" (parity mode of the hybrid encoder's ResNetV2 trunk: the 16 GroupNorm bottlenecks amplify
        bf16 storage rounding ~50x on random-init weights; everything after the trunk stays on the product path)."""
        assert precision in ("bf16", "fp32")
        if self.trunk_fp32 != (precision == "fp32"):
            self.trunk_fp32 = precision == "fp32"
            self.invalidate()
        return self

    def enable_graphs(self, on=True):
        """Replay each batch size's launch list as a CUDA graph (opt-in; SOCCDPT_CUDA_GRAPH=1 sets it at construction)."""
        self.use_graphs = bool(on)
        if not on:
            for p in self._plans.values():
                p.pop("graph", None)
        return self

    def invalidate(self):
        self._weights = None
        self._plans = {}

    # ------------------------------------------------------------------ weight packing
    def _pack(self, dev):
        enc = self.dpt.pretrained.model
        W = {"stages": []}
        self.hybrid = not isinstance(enc, SwinV2Params)
        if self.hybrid:
            W["hy"] = self._pack_hybrid(dev)
        else:
            self._pack_swin(W, enc, dev)
        self._pack_decoder(W, dev)
        return W

    @staticmethod
    def _std_weight(conv):
        """timm StdConv2dSame: per-output-channel standardisation with the biased variance, eps 1e-8 (fp32)."""
        w = conv.weight.detach().float()
        flat = w.reshape(w.shape[0], -1)
        mean = flat.mean(1, keepdim=True)
        var = flat.var(1, unbiased=False, keepdim=True)
        return ((flat - mean) * torch.rsqrt(var + conv.eps)).reshape_as(w)

    def _pack_hybrid(self, dev):
        pre = self.dpt.pretrained
        enc = pre.model
        bb = enc.patch_embed.backbone
        gn = lambda n: (_f32(n.weight, dev), _f32(n.bias, dev))
        H = dict(stem_w=_f32(self._std_weight(bb.stem.conv), dev), stem_n=gn(bb.stem.norm), stages=[])
        f32w = (lambda c: _f32(self._std_weight(c).permute(0, 2, 3, 1).contiguous(), dev)) if self.trunk_fp32 else (lambda c: None)
        H["stem_w32"] = f32w(bb.stem.conv)
        for stage in bb.stages:
            blocks = []
            for blk in stage.blocks:
                d = dict(stride=blk.stride, cin=blk.conv1.weight.shape[1], mid=blk.conv1.weight.shape[0],
                         cout=blk.conv3.weight.shape[0],
                         w1=_pack_conv(self._std_weight(blk.conv1), dev), n1=gn(blk.norm1),
                         w2=_pack_conv(self._std_weight(blk.conv2), dev), n2=gn(blk.norm2),
                         w3=_pack_conv(self._std_weight(blk.conv3), dev), n3=gn(blk.norm3), down=None)
                if blk.downsample is not None:
                    d["down"] = (_pack_conv(self._std_weight(blk.downsample.conv), dev), gn(blk.downsample.norm))
                    d["down32"] = f32w(blk.downsample.conv)
                d["w32"] = (f32w(blk.conv1), f32w(blk.conv2), f32w(blk.conv3))
                blocks.append(d)
            H["stages"].append(blocks)
        H["proj"] = (_pack_conv(enc.patch_embed.proj.weight, dev), _f32(enc.patch_embed.proj.bias, dev))
        H["cls"] = _f32(enc.cls_token.reshape(-1), dev)
        H["pos"] = _f32(enc.pos_embed[0], dev)
        H["heads"] = enc.num_heads
        H["blocks"] = [dict(n1=gn(b.norm1), wqkv=_bf16(b.attn.qkv.weight, dev), bqkv=_f32(b.attn.qkv.bias, dev),
                            wproj=_bf16(b.attn.proj.weight, dev), bproj=_f32(b.attn.proj.bias, dev), n2=gn(b.norm2),
                            w1=_bf16(b.mlp.fc1.weight, dev), b1=_f32(b.mlp.fc1.bias, dev),
                            w2=_bf16(b.mlp.fc2.weight, dev), b2=_f32(b.mlp.fc2.bias, dev)) for b in enc.blocks]
        H["hooks"] = tuple(enc.hooks)
        for idx in (3, 4):
            pp = getattr(pre, f"act_postprocess{idx}")
            d = dict(wr=_bf16(pp[0].project[0].weight, dev), br=_f32(pp[0].project[0].bias, dev),
                     w1=_pack_conv(pp[3].weight, dev), b1=_f32(pp[3].bias, dev))
            if idx == 4:
                d["w2"], d["b2"] = _pack_conv(pp[4].weight, dev), _f32(pp[4].bias, dev)
            H[f"pp{idx}"] = d
        return H

    def _pack_swin(self, W, enc, dev):
        pe = enc.patch_embed
        W["pe"] = (_f32(pe.proj.weight.reshape(pe.proj.weight.shape[0], -1), dev), _f32(pe.proj.bias, dev),
                   _f32(pe.norm.weight, dev), _f32(pe.norm.bias, dev))
        for layer in enc.layers:
            blocks = []
            for blk in layer.blocks:
                a = blk.attn
                C = a.qkv.weight.shape[1]
                bias = relative_position_bias_table(a, blk.window_size, layer.pretrained_window)
                sc = torch.clamp(a.logit_scale.detach().float(), max=math.log(1.0 / 0.01)).exp().reshape(-1)
                blocks.append(dict(
                    # TMA-fed attention on operands the qkv GEMM already normalised (csrc/attention_tma.cu): 16x16 windows whose
                    # logit scales admit the one-pass softmax bound (2.01 * scale + 16 < 80: every random-init head, most trained ones)
                    qscale=_f32(sc * math.log2(math.e), dev), one_pass=bool((2.01 * sc.max() + 16.0 < 80.0).item()),
                    ws=blk.window_size, shift=blk.shift_size, heads=blk.num_heads,
                    wqkv=_bf16(a.qkv.weight, dev),
                    bqkv=_f32(torch.cat([a.q_bias.detach().float(), torch.zeros(C, device=a.q_bias.device),
                                         a.v_bias.detach().float()]), dev),
                    biasT=_f32(bias, dev),                          # [h][(2ws-1)^2] relative-position table
                    scale=_f32(torch.clamp(a.logit_scale.detach().float(), max=math.log(1.0 / 0.01)).exp().reshape(-1), dev),
                    wproj=_bf16(a.proj.weight, dev), bproj=_f32(a.proj.bias, dev),
                    n1=(_f32(blk.norm1.weight, dev), _f32(blk.norm1.bias, dev)),
                    w1=_bf16(blk.mlp.fc1.weight, dev), b1=_f32(blk.mlp.fc1.bias, dev),
                    w2=_bf16(blk.mlp.fc2.weight, dev), b2=_f32(blk.mlp.fc2.bias, dev),
                    n2=(_f32(blk.norm2.weight, dev), _f32(blk.norm2.bias, dev))))
            ds = None
            if not isinstance(layer.downsample, torch.nn.Identity):
                # PatchMerging = cat(x[0::2,0::2], x[1::2,0::2], x[0::2,1::2], x[1::2,1::2]) -> Linear(4C, 2C): a 2x2 stride-2 conv
                # whose tap (kh, kw) takes the weight columns of the slice with row parity kh and column parity kw
                rw = layer.downsample.reduction.weight                       # (2C, 4C)
                Cd = rw.shape[1] // 4
                order = (0, 2, 1, 3)                                         # taps (0,0) (0,1) (1,0) (1,1) <- x0 x2 x1 x3
                rw_taps = torch.stack([rw[:, q * Cd:(q + 1) * Cd] for q in order], dim=1).contiguous()   # (2C, 4, C)
                ds = dict(w=_bf16(layer.downsample.reduction.weight, dev), w_taps=_bf16(rw_taps, dev),
                          n=(_f32(layer.downsample.norm.weight, dev), _f32(layer.downsample.norm.bias, dev)))
            W["stages"].append(dict(dim=layer.dim, res=layer.input_resolution, blocks=blocks, down=ds))

    @staticmethod
    def _fold_bn(w, b, bn):
        """eval-mode BatchNorm2d after a conv -> (scaled weight, bias): bn(conv(x) + b) = conv(x) * g + (b - mean) * g + beta."""
        g = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
        b0 = b.detach().float() if b is not None else torch.zeros_like(g)
        return w.detach().float() * g.view(-1, 1, 1, 1), (b0 - bn.running_mean.detach().float()) * g + bn.bias.detach().float()

    def _pack_rcu(self, rcu, dev):
        """ResidualConvUnit_custom (blocks.py:348-414); with ``use_bn`` (DPTSegmentationModel, dpt.py:240) each conv is followed
        by a BatchNorm2d, folded into the conv here."""
        out = []
        for conv, bn in ((rcu.conv1, getattr(rcu, "bn1", None)), (rcu.conv2, getattr(rcu, "bn2", None))):
            w, b = conv.weight, conv.bias
            if bn is not None:
                w, b = self._fold_bn(w, b, bn)
            out += [_pack_conv(w, dev), _f32(b, dev)]
        return tuple(out)

    def _pack_decoder(self, W, dev):
        sc = self.dpt.scratch
        W["rn"] = [_pack_conv(getattr(sc, f"layer{i}_rn").weight, dev) for i in (1, 2, 3, 4)]
        W["fusion"] = {}
        for i in (1, 2, 3, 4):
            f = getattr(sc, f"refinenet{i}")
            W["fusion"][i] = dict(
                out_w=_pack_conv(f.out_conv.weight, dev), out_b=_f32(f.out_conv.bias, dev),
                rcu1=self._pack_rcu(f.resConfUnit1, dev), rcu2=self._pack_rcu(f.resConfUnit2, dev))
        if "depth" in self.heads:
            oc = sc.output_conv
            # conv 2 of the head acts on a bilinear upsample: apply its nine tap matrices at low resolution instead
            # (rows = tap*32 + c), the gather kernel interpolates and sums them (csrc/depth_head.cu)
            w2 = oc[2].weight.detach().float()                       # (32, 128, 3, 3)
            w2t = w2.permute(2, 3, 0, 1).reshape(9 * w2.shape[0], w2.shape[1], 1, 1)
            W["dh"] = dict(w0=_pack_conv(oc[0].weight, dev), b0=_f32(oc[0].bias, dev),
                           w2t=_pack_conv(w2t, dev), b2=_f32(oc[2].bias, dev),
                           pw=_f32(oc[4].weight.reshape(1, -1), dev), pb=_f32(oc[4].bias, dev))
        if "seg" in self.heads:
            sh = self.seg_head
            w0, b0 = self._fold_bn(sh[0].weight, sh[0].bias, sh[1])
            W["sh"] = dict(w0=_pack_conv(w0, dev), b0=_f32(b0, dev),
                           pw=_f32(sh[4].weight.reshape(sh[4].weight.shape[0], -1), dev), pb=_f32(sh[4].bias, dev))
            W["num_classes"] = sh[4].weight.shape[0]
            W["seg_act"] = 0 if isinstance(sh[6], torch.nn.Sigmoid) else 1

    # ------------------------------------------------------------------ plan construction
    def _attn_tma(self, b, Hs, Ws):
        """16x16 windows on the tcgen05 engine: the qkv GEMM normalises q / k in its epilogue and the TMA-fed pipelined kernel
        (csrc/attention_tma.cu) consumes them.  SOCCDPT_ATTN_TMA=0 keeps the round-1 kernel (A/B runs); heads whose logit scale is too
        large for the one-pass softmax bound keep it too (it has an exact row-max pre-pass)."""
        return (self.conv_impl == "tcgen05" and b["ws"] == 16 and b["shift"] in (0, 8) and Hs % 16 == 0 and Ws % 16 == 0
                and b["one_pass"] and os.environ.get("SOCCDPT_ATTN_TMA", "1") != "0")

    def _conv(self, plan, x, w, N, H, Wd, Cin, Cout, K, bias=None, act=_cabi.ACT_NONE, res1=None, res2=None, y=None,
              y_relu=None, proj=None, stride=1, pad_trim=0, qk=None, up=None):
        c = _cabi.Conv()
        if up is not None:      # (low-resolution map, h, w): residual added through a bilinear x2 in the epilogue
            c.up_src, c.up_h, c.up_w = up[0].data_ptr(), up[1], up[2]
        if qk is not None:
            c.qk_scale, c.qk_heads = qk[0].data_ptr(), qk[1]
        c.stride, c.pad_trim = stride, pad_trim
        c.x, c.wgt = x.data_ptr(), w.data_ptr()
        c.bias = bias.data_ptr() if bias is not None else None
        c.res1 = res1.data_ptr() if res1 is not None else None
        c.res2 = res2.data_ptr() if res2 is not None else None
        c.y = y.data_ptr() if y is not None else None
        c.y_relu = y_relu.data_ptr() if y_relu is not None else None
        c.N, c.H, c.W, c.Cin, c.Cout, c.KH, c.KW, c.act = N, H, Wd, Cin, Cout, K, K, act
        if proj is not None:
            pw, pb, out, relu = proj
            c.proj_w, c.proj_b, c.proj_out = pw.data_ptr(), pb.data_ptr(), out.data_ptr()
            c.proj_n, c.proj_relu = pw.shape[0], int(relu)
        fn = self.lib.soccdpt_conv_fwd if self.conv_impl == "tcgen05" else self.lib.soccdpt_conv_ref_fwd
        plan["keep"].append(c)
        plan["ops"].append(_Launch("conv", fn, ctypes.byref(c)))

    def _fuse_tail(self, C, which):
        """Fused Swin block tails (csrc/swin_block_tail.cu): "mlp" (fc1 -> GELU -> fc2 -> norm2 -> residual, C <= 256), "proj"
        (proj -> norm1 -> residual, C <= 512) and "fc2" (fc2 -> norm2 -> residual behind the un-fused fc1 GEMM, for the stages
        whose rows are too wide for the two-GEMM form: 256 < C <= 512).  SOCCDPT_FUSED_TAIL = comma list or 0 switches them
        per kind for A/B runs (default: all).  The CUDA-core cross-check engine keeps the un-fused ops."""
        if self.conv_impl != "tcgen05" or C % 32:
            return False
        if C > (256 if which == "mlp" else 512) or (which == "fc2" and C <= 256):
            return False
        sel = os.environ.get("SOCCDPT_FUSED_TAIL", "mlp,proj,fc2")
        return which in sel.split(",")

    def _block_tail(self, plan, x, w1, b1, w2, b2, norm, master, y, M, C, HID, K1=None):
        a = _cabi.BlockTail()
        a.x, a.w2, a.b2 = x.data_ptr(), w2.data_ptr(), b2.data_ptr()
        a.w1 = w1.data_ptr() if w1 is not None else None
        a.b1 = b1.data_ptr() if b1 is not None else None
        a.gamma, a.beta, a.master, a.y = norm[0].data_ptr(), norm[1].data_ptr(), master.data_ptr(), y.data_ptr()
        a.M, a.K1, a.HID, a.C, a.eps = M, (K1 if K1 is not None else C), HID, C, 1e-5
        plan["keep"].append(a)
        plan["ops"].append(_Launch("block_tail_mlp" if w1 is not None else "block_tail_proj", self.lib.soccdpt_swin_block_tail_fwd,
                                   ctypes.byref(a)))

    def _plan_swin(self, plan, buf, x_in, B, img):
        Wt, lib, ops = self._weights, self.lib, plan["ops"]
        stages = Wt["stages"]
        E = stages[0]["dim"]
        g0 = img // 4
        maxLC = max(st["res"][0] * st["res"][1] * st["dim"] for st in stages)
        qkv = buf(B * maxLC * 3)
        att = buf(B * maxLC)
        tmp = buf(B * maxLC)
        hid = buf(B * maxLC * 4)
        cur = buf(B * g0 * g0, E)                               # bf16 copy of the residual stream (GEMM operand)
        master = buf(B * g0 * g0, E, dtype=torch.float32)       # fp32 residual stream
        ops.append(_Launch("patch_embed", lib.soccdpt_patch_embed_fwd, plan["x_arg"], *(t.data_ptr() for t in Wt["pe"]),
                           cur.data_ptr(), master.data_ptr(), B, img, img, E))
        taps = []
        for si, st in enumerate(stages):
            Hs, Ws = st["res"]
            C, L = st["dim"], Hs * Ws
            M = B * L
            for b in st["blocks"]:
                if self._attn_tma(b, Hs, Ws):
                    self._conv(plan, cur, b["wqkv"], 1, 1, M, C, 3 * C, 1, bias=b["bqkv"], y=qkv, qk=(b["qscale"], b["heads"]))
                    ops.append(_Launch("window_attention", lib.soccdpt_window_attention_normed_fwd, qkv.data_ptr(),
                                       b["biasT"].data_ptr(), b["scale"].data_ptr(), att.data_ptr(), B, Hs, Ws, C, b["heads"], b["shift"]))
                else:
                    self._conv(plan, cur, b["wqkv"], 1, 1, M, C, 3 * C, 1, bias=b["bqkv"], y=qkv)
                    ops.append(_Launch("window_attention", lib.soccdpt_window_attention_fwd, qkv.data_ptr(), b["biasT"].data_ptr(),
                                       b["scale"].data_ptr(), att.data_ptr(), B, Hs, Ws, C, b["heads"], b["ws"], b["shift"]))
                if self._fuse_tail(C, "proj"):
                    self._block_tail(plan, att, None, None, b["wproj"], b["bproj"], b["n1"], master, cur, M, C, 0)
                else:
                    self._conv(plan, att, b["wproj"], 1, 1, M, C, C, 1, bias=b["bproj"], y=tmp)
                    ops.append(_Launch("ln_res", lib.soccdpt_layernorm_master_fwd, tmp.data_ptr(), master.data_ptr(), 1,
                                       b["n1"][0].data_ptr(), b["n1"][1].data_ptr(), cur.data_ptr(), M, C, ctypes.c_float(1e-5)))
                if self._fuse_tail(C, "mlp"):
                    # fc1 -> GELU -> fc2 -> norm2 -> residual in one kernel; y aliases x (a tile's rows are read before they are written)
                    self._block_tail(plan, cur, b["w1"], b["b1"], b["w2"], b["b2"], b["n2"], master, cur, M, C, 4 * C)
                elif self._fuse_tail(C, "fc2"):
                    # wide rows (256 < C <= 512): fc1 + GELU stays a plain GEMM, fc2 -> norm2 -> residual is one kernel
                    self._conv(plan, cur, b["w1"], 1, 1, M, C, 4 * C, 1, bias=b["b1"], act=_cabi.ACT_GELU, y=hid)
                    self._block_tail(plan, hid, None, None, b["w2"], b["b2"], b["n2"], master, cur, M, C, 0, K1=4 * C)
                else:
                    self._conv(plan, cur, b["w1"], 1, 1, M, C, 4 * C, 1, bias=b["b1"], act=_cabi.ACT_GELU, y=hid)
                    self._conv(plan, hid, b["w2"], 1, 1, M, 4 * C, C, 1, bias=b["b2"], y=tmp)
                    ops.append(_Launch("ln_res", lib.soccdpt_layernorm_master_fwd, tmp.data_ptr(), master.data_ptr(), 1,
                                       b["n2"][0].data_ptr(), b["n2"][1].data_ptr(), cur.data_ptr(), M, C, ctypes.c_float(1e-5)))
            taps.append((cur, Hs, Ws, C))   # hooks sit on the last block of every stage (dpt.py:61-72)
            if st["down"] is not None:
                M2 = B * L // 4
                # the 2x2 gather lives in the TMA box of the implicit GEMM (element strides 2, tap offsets (kh, kw))
                self._conv(plan, cur, st["down"]["w_taps"], B, Hs, Ws, C, 2 * C, 2, y=tmp, stride=2, pad_trim=1)
                nxt = buf(M2, 2 * C)
                master = buf(M2, 2 * C, dtype=torch.float32)
                ops.append(_Launch("ln", lib.soccdpt_layernorm_master_fwd, tmp.data_ptr(), master.data_ptr(), 0,
                                   st["down"]["n"][0].data_ptr(), st["down"]["n"][1].data_ptr(), nxt.data_ptr(), M2, 2 * C,
                                   ctypes.c_float(1e-5)))
                cur = nxt
        return taps

    def _plan_hybrid(self, plan, buf, x_in, B, img):
        """timm vit_base_resnet50_384 forward_flex + the reference's tap post-processing (vit.py:44-85, 179-219)."""
        Hy, lib, ops = self._weights["hy"], self.lib, plan["ops"]
        gn_scratch = buf(B * 64, dtype=torch.float64)

        def gn(x, n, HW, C, relu, shortcut=None, y=None):
            ops.append(_Launch("groupnorm", lib.soccdpt_groupnorm_fwd, x.data_ptr(), n[0].data_ptr(), n[1].data_ptr(),
                               shortcut.data_ptr() if shortcut is not None else None, (y if y is not None else x).data_ptr(),
                               B, HW, C, ctypes.c_float(1e-5), int(relu), gn_scratch.data_ptr()))

        if self.trunk_fp32:
            stage_out = self._plan_trunk_fp32(plan, buf, B, img)
            return self._plan_vit(plan, buf, B, stage_out)

        # ---- ResNetV2 stem: StdConv 7x7/2 (SAME) -> GroupNorm + ReLU -> MaxPool 3x3/2 (SAME)
        H1 = (img + 1) // 2
        s0 = buf(B, H1, H1, 64)
        ops.append(_Launch("stem_conv7", lib.soccdpt_stem_conv7_fwd, plan["x_arg"], Hy["stem_w"].data_ptr(), s0.data_ptr(), B, img, img))
        gn(s0, Hy["stem_n"], H1 * H1, 64, True)
        Hc = (H1 + 1) // 2
        cur = buf(B, Hc, Hc, 64)
        ops.append(_Launch("maxpool", lib.soccdpt_maxpool3s2_fwd, s0.data_ptr(), cur.data_ptr(), B, H1, H1, 64))

        # ---- bottleneck stages (non pre-activation): out = relu(gn3(conv3(gn2(conv2(gn1(conv1 x))))) + shortcut)
        stage_out = []
        for blocks in Hy["stages"]:
            for b in blocks:
                st, cin, mid, cout = b["stride"], b["cin"], b["mid"], b["cout"]
                Ho = (Hc + st - 1) // st
                shortcut = cur
                if b["down"] is not None:
                    shortcut = buf(B, Ho, Ho, cout)
                    self._conv(plan, cur, b["down"][0], B, Hc, Hc, cin, cout, 1, y=shortcut, stride=st)
                    gn(shortcut, b["down"][1], Ho * Ho, cout, False)
                a = buf(B, Hc, Hc, mid)
                self._conv(plan, cur, b["w1"], B, Hc, Hc, cin, mid, 1, y=a)
                gn(a, b["n1"], Hc * Hc, mid, True)
                c2 = buf(B, Ho, Ho, mid)
                # TF "SAME" on an even input: stride 2 pads 0 in front / 1 behind, stride 1 pads 1 / 1
                assert st == 1 or Hc % 2 == 0
                self._conv(plan, a, b["w2"], B, Hc, Hc, mid, mid, 3, y=c2, stride=st, pad_trim=1 if st == 2 else 0)
                gn(c2, b["n2"], Ho * Ho, mid, True)
                c3 = buf(B, Ho, Ho, cout)
                self._conv(plan, c2, b["w3"], B, Ho, Ho, mid, cout, 1, y=c3)
                gn(c3, b["n3"], Ho * Ho, cout, True, shortcut=shortcut)
                cur, Hc = c3, Ho
            stage_out.append((cur, Hc, Hc, blocks[-1]["cout"]))
        return self._plan_vit(plan, buf, B, stage_out)

    def _plan_trunk_fp32(self, plan, buf, B, img):
        """The ResNetV2 trunk of _plan_hybrid in the fp32-storage parity mode: same operator sequence, fp32 NHWC activations,
        fp32 standardised weights, CUDA-core kernels (csrc/trunk_fp32.cu); the three stage outputs are rounded to bf16 once."""
        Hy, lib, ops = self._weights["hy"], self.lib, plan["ops"]
        f32 = torch.float32
        eps = ctypes.c_float(1e-5)

        def conv(x, w, H, cin, cout, k, st):
            Ho = (H + st - 1) // st
            y = buf(B, Ho, Ho, cout, dtype=f32)
            ops.append(_Launch("conv_f32", lib.soccdpt_conv_f32_fwd, x.data_ptr(), w.data_ptr(), y.data_ptr(), B, H, H, cin, cout, k, st))
            return y, Ho

        def gn(x, n, HW, C, relu, shortcut=None):
            ops.append(_Launch("groupnorm_f32", lib.soccdpt_groupnorm_f32_fwd, x.data_ptr(), n[0].data_ptr(), n[1].data_ptr(),
                               shortcut.data_ptr() if shortcut is not None else None, x.data_ptr(), B, HW, C, eps, int(relu)))

        xh = buf(B, img, img, 3, dtype=f32)
        ops.append(_Launch("nchw_to_nhwc", lib.soccdpt_nchw_to_nhwc_f32, plan["x_arg"], xh.data_ptr(), B, 3, img * img))
        s0, H1 = conv(xh, Hy["stem_w32"], img, 3, 64, 7, 2)
        gn(s0, Hy["stem_n"], H1 * H1, 64, True)
        Hc = (H1 + 1) // 2
        cur = buf(B, Hc, Hc, 64, dtype=f32)
        ops.append(_Launch("maxpool_f32", lib.soccdpt_maxpool3s2_f32_fwd, s0.data_ptr(), cur.data_ptr(), B, H1, H1, 64))
        stage_out = []
        for blocks in Hy["stages"]:
            for b in blocks:
                st, cin, mid, cout = b["stride"], b["cin"], b["mid"], b["cout"]
                shortcut = cur
                if b["down"] is not None:
                    shortcut, Ho = conv(cur, b["down32"], Hc, cin, cout, 1, st)
                    gn(shortcut, b["down"][1], Ho * Ho, cout, False)
                a, _ = conv(cur, b["w32"][0], Hc, cin, mid, 1, 1)
                gn(a, b["n1"], Hc * Hc, mid, True)
                c2, Ho = conv(a, b["w32"][1], Hc, mid, mid, 3, st)
                gn(c2, b["n2"], Ho * Ho, mid, True)
                c3, _ = conv(c2, b["w32"][2], Ho, mid, cout, 1, 1)
                gn(c3, b["n3"], Ho * Ho, cout, True, shortcut=shortcut)
                cur, Hc = c3, Ho
            o16 = buf(B, Hc, Hc, blocks[-1]["cout"])
            ops.append(_Launch("f32_to_bf16", lib.soccdpt_f32_to_bf16, cur.data_ptr(), o16.data_ptr(), cur.numel()))
            stage_out.append((o16, Hc, Hc, blocks[-1]["cout"]))
        return stage_out

    def _plan_vit(self, plan, buf, B, stage_out):
        """Patch projection, the 12 ViT blocks and the tap post-processing of _plan_hybrid (vit.py:44-85, 179-219)."""
        Hy, lib, ops = self._weights["hy"], self.lib, plan["ops"]

        # ---- patch projection (1x1 conv == linear), cls token + position embedding
        feat, g, _, Cf = stage_out[2]
        D, L = Hy["proj"][0].shape[0], g * g
        assert Hy["pos"].shape[0] == L + 1, "dpt_hybrid_384: the position embedding is used at its native grid (384x384 frames)"
        patches = buf(B, L, D)
        self._conv(plan, feat, Hy["proj"][0], 1, 1, B * L, Cf, D, 1, bias=Hy["proj"][1], y=patches)
        N = L + 1
        M = B * N
        xn = buf(M, D)                                  # bf16 tokens, then LayerNorm outputs (GEMM operand)
        master = buf(M, D, dtype=torch.float32)         # fp32 residual stream
        ops.append(_Launch("vit_tokens", lib.soccdpt_vit_tokens_fwd, patches.data_ptr(), Hy["cls"].data_ptr(), Hy["pos"].data_ptr(),
                           xn.data_ptr(), master.data_ptr(), B, L, D))
        qkv, att, tmp, hid = buf(M, 3 * D), buf(M, D), buf(M, D), buf(M, 4 * D)
        heads = Hy["heads"]
        hooked = {}
        pending = None                                  # branch output not yet added to the residual stream
        eps = ctypes.c_float(1e-6)

        def prenorm(t, n, y, stream_copy=None):
            ops.append(_Launch("prenorm", lib.soccdpt_prenorm_fwd, t.data_ptr() if t is not None else None, master.data_ptr(),
                               n[0].data_ptr() if n is not None else None, n[1].data_ptr() if n is not None else None,
                               y.data_ptr() if y is not None else None,
                               stream_copy.data_ptr() if stream_copy is not None else None, M, D, eps))

        nblk = len(Hy["blocks"])
        for i, b in enumerate(Hy["blocks"]):
            hook_prev = hooked.get(i - 1)               # the previous block's output is complete after this add
            prenorm(pending, b["n1"], xn, hook_prev)
            self._conv(plan, xn, b["wqkv"], 1, 1, M, D, 3 * D, 1, bias=b["bqkv"], y=qkv)
            ops.append(_Launch("global_attention", lib.soccdpt_global_attention_fwd, qkv.data_ptr(), att.data_ptr(), B, N, heads, D // heads))
            self._conv(plan, att, b["wproj"], 1, 1, M, D, D, 1, bias=b["bproj"], y=tmp)
            prenorm(tmp, b["n2"], xn)
            self._conv(plan, xn, b["w1"], 1, 1, M, D, 4 * D, 1, bias=b["b1"], act=_cabi.ACT_GELU, y=hid)
            self._conv(plan, hid, b["w2"], 1, 1, M, 4 * D, D, 1, bias=b["b2"], y=tmp)
            pending = tmp
            if i in Hy["hooks"][2:]:
                hooked[i] = buf(B, N, D)
            if i == nblk - 1:                           # last add; the final LayerNorm is computed-and-discarded upstream
                prenorm(pending, None, None, hooked.get(i))

        # ---- tap post-processing: ProjectReadout (cat with cls, Linear + GELU) -> 1x1 conv [-> 3x3 stride-2 conv]
        def readout(tok, w):
            feats = buf(B, L, 2 * D)
            ops.append(_Launch("readout_concat", lib.soccdpt_readout_concat_fwd, tok.data_ptr(), feats.data_ptr(), B, L, D))
            r = buf(B, g, g, D)
            self._conv(plan, feats, w["wr"], 1, 1, B * L, 2 * D, D, 1, bias=w["br"], act=_cabi.ACT_GELU, y=r)
            o = buf(B, g, g, w["w1"].shape[0])
            self._conv(plan, r, w["w1"], B, g, g, D, w["w1"].shape[0], 1, bias=w["b1"], y=o)
            return o

        h3, h4 = Hy["hooks"][2], Hy["hooks"][3]
        l3 = readout(hooked[h3], Hy["pp3"])
        l4a = readout(hooked[h4], Hy["pp4"])
        C4 = Hy["pp4"]["w2"].shape[0]
        g4 = (g + 1) // 2
        l4 = buf(B, g4, g4, C4)
        self._conv(plan, l4a, Hy["pp4"]["w2"], B, g, g, l4a.shape[-1], C4, 3, bias=Hy["pp4"]["b2"], y=l4, stride=2)
        return [stage_out[0], stage_out[1], (l3, g, g, l3.shape[-1]), (l4, g4, g4, C4)]

    def _build_plan(self, B, dev):
        Wt = self._weights
        lib = self.lib
        plan = {"ops": [], "keep": [], "B": B}
        ops = plan["ops"]

        def buf(*shape, dtype=torch.bfloat16):
            t = torch.empty(shape, device=dev, dtype=dtype)
            plan["keep"].append(t)
            return t

        enc = self.dpt.pretrained.model
        img = enc.img_size
        plan["img"] = img
        x_in = buf(B, 3, img, img, dtype=torch.float32)
        plan["x_in"] = x_in
        # the network input is read through this slot: the caller's tensor when it is contiguous fp32 (no copy), else / under
        # CUDA-graph replay the plan's own x_in
        plan["x_arg"] = ctypes.c_void_p(x_in.data_ptr())
        taps = self._plan_hybrid(plan, buf, x_in, B, img) if self.hybrid else self._plan_swin(plan, buf, x_in, B, img)
        plan["taps"] = taps

        # ---------------- decoder (dpt.py:152-172)
        F = self.dpt.features
        lv = []
        for i, (t, Hs, Ws, C) in enumerate(taps):
            y, yr = buf(B, Hs, Ws, F), buf(B, Hs, Ws, F)
            self._conv(plan, t, Wt["rn"][i], B, Hs, Ws, C, F, 3, y=y, y_relu=yr)
            lv.append((y, yr, Hs, Ws))

        def rcu(w, x, xr, H, Wd, res2=None, want_relu=False, up=None):
            c1 = buf(B, H, Wd, F)
            self._conv(plan, xr, w[0], B, H, Wd, F, F, 3, bias=w[1], act=_cabi.ACT_RELU, y=c1)
            o = buf(B, H, Wd, F)
            orl = buf(B, H, Wd, F) if want_relu else None
            self._conv(plan, c1, w[2], B, H, Wd, F, F, 3, bias=w[3], res1=x, res2=res2, up=up, y=o, y_relu=orl)
            return o, orl

        # the x2 upsample between two fusion blocks is folded into the consumer's epilogue (soccdpt_conv_t.up_src): its only reader
        # is the residual add `xs[0] + resConfUnit1(xs[1])`; the last one (path_1) feeds the heads' 3x3 convs and stays a kernel
        fold_up = os.environ.get("SOCCDPT_FOLD_UPSAMPLE", "1") != "0" and F % 32 == 0
        path = up_low = None
        for i in (4, 3, 2, 1):
            fw = Wt["fusion"][i]
            y, yr, H, Wd = lv[i - 1]
            if path is None and up_low is None:
                s, sr = y, yr
            else:   # output = xs[0] + resConfUnit1(xs[1])   (blocks.py:476-479)
                s, sr = rcu(fw["rcu1"], y, yr, H, Wd, res2=path, up=up_low, want_relu=True)
            o, _ = rcu(fw["rcu2"], s, sr, H, Wd)
            low = buf(B, H, Wd, F)
            self._conv(plan, o, fw["out_w"], B, H, Wd, F, F, 1, bias=fw["out_b"], y=low)   # out_conv before the upsample
            TH, TW = (lv[i - 2][2], lv[i - 2][3]) if i > 1 else (2 * H, 2 * Wd)
            if fold_up and i > 1 and TH == 2 * H and TW == 2 * Wd:
                path, up_low = None, (low, H, Wd)
                continue
            path, up_low = buf(B, TH, TW, F), None
            ops.append(_Launch("upsample", lib.soccdpt_upsample_bilinear_fwd, low.data_ptr(), path.data_ptr(), B, H, Wd, TH, TW, F))
        PH, PW = 2 * lv[0][2], 2 * lv[0][3]
        plan["path_1"] = path

        # ---------------- heads
        plan["depth"] = plan["seg"] = None
        if "depth" in self.heads:
            dh = Wt["dh"]
            d0 = buf(B, PH, PW, F // 2)
            self._conv(plan, path, dh["w0"], B, PH, PW, F, F // 2, 3, bias=dh["b0"], y=d0)
            assert dh["w2t"].shape[0] == 9 * 32, "depth head: head_features_2 must be 32"
            taps = buf(B, PH, PW, 9 * 32)
            self._conv(plan, d0, dh["w2t"], B, PH, PW, F // 2, 9 * 32, 1, y=taps)
            depth = buf(B, 2 * PH, 2 * PW, dtype=torch.float32)
            ops.append(_Launch("depth_tail", lib.soccdpt_depth_tail_fwd, taps.data_ptr(), dh["b2"].data_ptr(), dh["pw"].data_ptr(),
                               dh["pb"].data_ptr(), depth.data_ptr(), B, PH, PW))
            plan["depth"] = depth
        if "seg" in self.heads:
            sh = Wt["sh"]
            P = Wt["num_classes"]
            logits = buf(B, PH, PW, P, dtype=torch.float32)
            self._conv(plan, path, sh["w0"], B, PH, PW, F, F, 3, bias=sh["b0"], act=_cabi.ACT_RELU,
                       proj=(sh["pw"], sh["pb"], logits, False))
            plan["seg_logits"] = logits      # (B, PH, PW, P) fp32: the head's output before the x2 upsample and the activation
            seg = buf(B, P, 2 * PH, 2 * PW, dtype=torch.float32)
            ops.append(_Launch("seg_finish", lib.soccdpt_seg_finish_fwd, logits.data_ptr(), seg.data_ptr(), B, PH, PW, P, Wt["seg_act"]))
            plan["seg"] = seg
        return plan

    # ------------------------------------------------------------------ run
    def plan_for(self, B, dev):
        dev = _cabi.normalize_device(dev)
        if self._weights is None:
            with torch.cuda.device(dev):
                self._weights = self._pack(dev)
        key = (B, dev.index)
        plan = self._plans.pop(key, None)
        if plan is None:
            with torch.cuda.device(dev):
                plan = self._build_plan(B, dev)
            while len(self._plans) >= max(self.max_plans, 1):       # ragged last batches must not pile up activation sets
                self._plans.pop(next(iter(self._plans)))
        self._plans[key] = plan                                      # most recently used last
        return plan

    def bind_input(self, plan, x=None):
        """Points the plan's first kernel at ``x`` (contiguous fp32: read in place) or at the plan's own copy of it."""
        if x is not None and x.dtype == torch.float32 and x.is_contiguous() and not self.use_graphs:
            plan["x_arg"].value = x.data_ptr()
            plan["x_ref"] = x                                        # keeps the caller's tensor alive while the launches are queued
        else:
            if x is not None:
                plan["x_in"].copy_(x, non_blocking=True)
            plan["x_arg"].value = plan["x_in"].data_ptr()
            plan["x_ref"] = None

    def run(self, x):
        """x: (B,3,S,S) fp32 CUDA tensor -> (inverse depth (B,S,S) f32, segmentation (B,C,S,S) f32).
        The returned tensors are the plan's static output buffers (overwritten by the next call of this batch size)."""
        if not x.is_cuda:
            raise _cabi.SoccdptError("soccdpt_b200 runs on CUDA devices only (no CPU fallback)")
        B = x.shape[0]
        plan = self.plan_for(B, x.device)
        if tuple(x.shape[1:]) != (3, plan["img"], plan["img"]):
            raise AssertionError("Input image size doesn't match model")
        with torch.cuda.device(x.device):      # the C ABI launches on the current device / its current stream
            self.bind_input(plan, x)
            if self.use_graphs:
                # CUDA-graph replay of the plan's launch list (all buffers are static, every entry point only enqueues on the
                # current stream): the first call of a plan runs eagerly (sets the kernels' attributes), the second is captured.
                # What it buys is the host side: ~130 ctypes launches per forward cost more than the kernels at small batches.
                g = plan.get("graph")
                if g is None and plan.get("warm", False):
                    g = torch.cuda.CUDAGraph()
                    torch.cuda.synchronize(x.device)
                    with torch.cuda.graph(g):
                        stream = _cabi.current_stream(x.device)
                        for op in plan["ops"]:
                            op(stream)
                    plan["graph"] = g
                if g is not None:
                    g.replay()
                    return plan["depth"], plan["seg"]
                plan["warm"] = True
            stream = _cabi.current_stream(x.device)
            for op in plan["ops"]:
                op(stream)
        return plan["depth"], plan["seg"]

    def launches_per_forward(self, B, dev):
        return len(self.plan_for(B, dev)["ops"])
