"""TEST INFRASTRUCTURE ONLY -- the oracle's dpt_hybrid_384 network with the PRODUCT'S STORAGE ROUNDING emulated.

The CUDA path computes every contraction with bf16 operands and fp32 accumulation and stores activations in bf16
(fp32 for the ViT residual stream, the logits and the outputs).  With random-init weights the 16 GroupNorm bottlenecks of
the ResNetV2 trunk amplify that operand rounding roughly x3 every four blocks (measured: tools/debug_hybrid.py), so the
distance to the pure-fp32 oracle says little about whether the ALGORITHM is right.  This module evaluates the same
algorithm as oracle/soccdpt_oracle.py (same state_dict, same operators, fp32 math on CPU) but rounds to bf16 exactly
where the product stores bf16: conv / linear weights, every conv, GroupNorm, LayerNorm, attention and up-sample output.
The CUDA path must sit on this emulation to within accumulation-order noise; the fp32 oracle then bounds how far the
storage format moves the result.  Reference call sites: see soccdpt_oracle.OracleV3._hybrid_taps / decoder / heads.
"""
import math

import torch
import torch.nn.functional as F


ROUND = True      # False: no rounding anywhere -> must reproduce the fp32 oracle (tests/test_oracle_golden.py checks that)


def r(t):
    """bf16 round-to-nearest-even, kept in fp32."""
    return t.to(torch.bfloat16).to(torch.float32) if ROUND else t


def _same(x, k, s, value=0.0):
    ih, iw = x.shape[-2:]
    ph = max((math.ceil(ih / s) - 1) * s + k - ih, 0)
    pw = max((math.ceil(iw / s) - 1) * s + k - iw, 0)
    return F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2], value=value) if (ph or pw) else x


def _std(w, eps=1e-8):
    flat = w.reshape(w.shape[0], -1)
    m = flat.mean(1, keepdim=True)
    v = flat.var(1, unbiased=False, keepdim=True)
    return ((flat - m) * torch.rsqrt(v + eps)).reshape_as(w)


def _stdconv(x, w, stride=1, round_w=True):
    w = _std(w)
    return r(F.conv2d(_same(x, w.shape[-1], stride), r(w) if round_w else w, None, stride))


def _gn(sd, p, x, relu, shortcut=None):
    y = F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-5)
    if shortcut is not None:
        y = y + shortcut
    return r(F.relu(y) if relu else y)


def _lin(x, w, b=None):
    return F.linear(x, r(w), b)


@torch.no_grad()
def hybrid_taps(sd, x, heads=12, hooks=(8, 11)):
    P = "depth_net.pretrained."
    bb = P + "model.patch_embed.backbone."
    t = _stdconv(x, sd[bb + "stem.conv.weight"], 2, round_w=False)          # the stem kernel keeps fp32 weights
    t = _gn(sd, bb + "stem.norm", t, True)
    t = F.max_pool2d(_same(t, 3, 2, -float("inf")), 3, 2)
    outs = []
    for s, depth in enumerate((3, 4, 9)):
        for b in range(depth):
            p = f"{bb}stages.{s}.blocks.{b}."
            st = 2 if (b == 0 and s > 0) else 1
            sc = t
            if b == 0:
                sc = _gn(sd, p + "downsample.norm", _stdconv(t, sd[p + "downsample.conv.weight"], st), False)
            a = _gn(sd, p + "norm1", _stdconv(t, sd[p + "conv1.weight"]), True)
            a = _gn(sd, p + "norm2", _stdconv(a, sd[p + "conv2.weight"], st), True)
            t = _gn(sd, p + "norm3", _stdconv(a, sd[p + "conv3.weight"]), True, shortcut=sc)
        outs.append(t)
    m = P + "model."
    B, _, gh, gw = t.shape
    pt = r(F.conv2d(t, r(sd[m + "patch_embed.proj.weight"]), sd[m + "patch_embed.proj.bias"]))
    pt = pt.flatten(2).transpose(1, 2)
    master = torch.cat((sd[m + "cls_token"].expand(B, -1, -1), pt), 1) + sd[m + "pos_embed"]
    D = master.shape[-1]
    N = master.shape[1]
    hooked = {}
    for i in range(12):
        q = f"{m}blocks.{i}."
        y = r(F.layer_norm(master, (D,), sd[q + "norm1.weight"], sd[q + "norm1.bias"], 1e-6))
        qkv = r(_lin(y, sd[q + "attn.qkv.weight"], sd[q + "attn.qkv.bias"])).reshape(B, N, 3, heads, D // heads).permute(2, 0, 3, 1, 4)
        att = ((qkv[0] * (D // heads) ** -0.5) @ qkv[1].transpose(-2, -1)).softmax(-1) @ qkv[2]
        att = r(att.transpose(1, 2).reshape(B, N, D))
        master = master + r(_lin(att, sd[q + "attn.proj.weight"], sd[q + "attn.proj.bias"]))
        y = r(F.layer_norm(master, (D,), sd[q + "norm2.weight"], sd[q + "norm2.bias"], 1e-6))
        h = r(F.gelu(_lin(y, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"])))
        master = master + r(_lin(h, sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"]))
        if i in hooks:
            hooked[i] = r(master)

    def readout(tk, idx):
        pp = f"{P}act_postprocess{idx}."
        feats = torch.cat((tk[:, 1:], tk[:, :1].expand(-1, N - 1, -1)), -1)
        y = r(F.gelu(_lin(feats, sd[pp + "0.project.0.weight"], sd[pp + "0.project.0.bias"])))
        y = y.transpose(1, 2).unflatten(2, (gh, gw))
        return r(F.conv2d(y, r(sd[pp + "3.weight"]), sd[pp + "3.bias"]))

    l3 = readout(hooked[hooks[0]], 3)
    l4 = readout(hooked[hooks[1]], 4)
    pp = P + "act_postprocess4.4."
    l4 = r(F.conv2d(l4, r(sd[pp + "weight"]), sd[pp + "bias"], stride=2, padding=1))
    return outs[0], outs[1], l3, l4


def _c3(sd, name, x, k=3):
    return F.conv2d(x, r(sd[name + ".weight"]), sd.get(name + ".bias"), padding=k // 2)


@torch.no_grad()
def decoder(sd, taps):
    s = "depth_net.scratch."
    lv = [_c3(sd, f"{s}layer{i + 1}_rn", t) for i, t in enumerate(taps)]       # fp32 accumulators
    path = None
    for i in (4, 3, 2, 1):
        p = f"{s}refinenet{i}."
        y = lv[i - 1]
        if path is None:
            st = y                                                                # fp32 value before storage
        else:
            c1 = r(F.relu(_c3(sd, p + "resConfUnit1.conv1", r(F.relu(y)))))
            st = _c3(sd, p + "resConfUnit1.conv2", c1) + r(y) + path
        c1 = r(F.relu(_c3(sd, p + "resConfUnit2.conv1", r(F.relu(st)))))
        o = r(_c3(sd, p + "resConfUnit2.conv2", c1) + r(st))
        low = r(_c3(sd, p + "out_conv", o, 1))                                    # out_conv before the up-sample
        kw = {"size": lv[i - 2].shape[2:]} if i > 1 else {"scale_factor": 2}
        path = r(F.interpolate(low, **kw, mode="bilinear", align_corners=True))
    return path


@torch.no_grad()
def heads(sd, path, sigmoid=True):
    h = "depth_net.scratch.output_conv."
    d0 = r(_c3(sd, h + "0", path))
    w2 = sd[h + "2.weight"]                                                       # (32, 128, 3, 3)
    T = r(F.conv2d(d0, r(w2.permute(2, 3, 0, 1).reshape(9 * w2.shape[0], w2.shape[1], 1, 1))))
    Tu = F.interpolate(T, scale_factor=2, mode="bilinear", align_corners=True)
    H, W = Tu.shape[2:]
    Tp = F.pad(Tu, [1, 1, 1, 1])
    acc = sd[h + "2.bias"].view(1, -1, 1, 1).expand(T.shape[0], -1, H, W).clone()
    for dy in range(3):
        for dx in range(3):
            tap = dy * 3 + dx
            acc = acc + Tp[:, tap * 32:(tap + 1) * 32, dy:dy + H, dx:dx + W]
    d = F.relu(F.conv2d(F.relu(acc), sd[h + "4.weight"], sd[h + "4.bias"])).squeeze(1)
    inv_std = torch.rsqrt(sd["seg_head.1.running_var"] + 1e-5)
    g = sd["seg_head.1.weight"] * inv_std
    w0 = r(sd["seg_head.0.weight"] * g.view(-1, 1, 1, 1))
    b0 = sd["seg_head.1.bias"] - sd["seg_head.1.running_mean"] * g
    s = F.relu(F.conv2d(path, w0, b0, padding=1))
    s = F.conv2d(s, sd["seg_head.4.weight"], sd["seg_head.4.bias"])
    s = F.interpolate(s, scale_factor=2, mode="bilinear", align_corners=True)
    return d, (torch.sigmoid(s) if sigmoid else 0.5 * torch.tanh(s) + 0.5)


@torch.no_grad()
def hybrid_network(sd, x, sigmoid=True):
    """image -> (inverse depth, segmentation, path_1, taps) with the product's storage rounding."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    taps = hybrid_taps(sd, x)
    path = decoder(sd, taps)
    d, g = heads(sd, path, sigmoid)
    return d, g, path, taps
