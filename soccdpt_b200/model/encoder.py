"""Parameter containers for the SwinV2 encoders the reference pulls from ``timm==0.6.12``
(SOccDPT/model/backbones/swin2.py:15-30, swin_common.py:12-54).

These modules only OWN the weights, under exactly timm's attribute names
so that MiDaS / SOccDPT checkpoints keyed ``pretrained.model.*`` load and ``state_dict()`` /
``parameters()`` / the freeze helpers of the reference keep working.  They have no forward(): the
arithmetic runs in the CUDA kernels driven by ``soccdpt_b200.engine``.
"""
import math

import torch
import torch.nn as nn

SWIN_CONFIGS = {
    # backbone: (timm name, img, window, embed, depths, heads, pretrained_window_sizes, hooks)
    "swin2t16_256": ("swinv2_tiny_window16_256", 256, 16, 96, (2, 2, 6, 2), (3, 6, 12, 24), (0, 0, 0, 0), (1, 1, 5, 1)),
    "swin2b24_384": ("swinv2_base_window12to24_192to384_22kft1k", 384, 24, 128, (2, 2, 18, 2), (4, 8, 16, 32),
                     (12, 12, 12, 6), (1, 1, 17, 1)),
}


def _no_forward(self, *a, **k):
    raise RuntimeError("parameter container: the forward pass runs in soccdpt_b200.engine (CUDA only)")


class _WindowAttentionParams(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((heads, 1, 1))))
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.cpb_mlp = nn.Sequential(nn.Linear(2, 512, bias=True), nn.ReLU(inplace=True), nn.Linear(512, heads, bias=False))
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.proj = nn.Linear(dim, dim)

    forward = _no_forward


class _MlpParams(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    forward = _no_forward


def shift_attn_mask(res, ws, shift):
    """timm's {0,-100} attn_mask buffer of a shifted block: (nW, N, N)."""
    H, W = res
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = img.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0))


class _BlockParams(nn.Module):
    def __init__(self, dim, res, heads, window, shift):
        super().__init__()
        self.input_resolution = res
        self.window_size = min(window, res[0])
        self.shift_size = 0 if res[0] <= window else shift
        self.num_heads = heads
        self.attn = _WindowAttentionParams(dim, heads)
        self.norm1 = nn.LayerNorm(dim)
        self.mlp = _MlpParams(dim, dim * 4)
        self.norm2 = nn.LayerNorm(dim)
        self.register_buffer(
            "attn_mask", shift_attn_mask(res, self.window_size, self.shift_size) if self.shift_size > 0 else None)

    forward = _no_forward


class _PatchMergingParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(2 * dim)

    forward = _no_forward


class _LayerParams(nn.Module):
    def __init__(self, dim, res, depth, heads, window, downsample, pretrained_window):
        super().__init__()
        self.dim, self.input_resolution, self.pretrained_window = dim, res, pretrained_window
        self.blocks = nn.ModuleList(
            [_BlockParams(dim, res, heads, window, 0 if i % 2 == 0 else window // 2) for i in range(depth)])
        self.downsample = _PatchMergingParams(dim) if downsample else nn.Identity()

    forward = _no_forward


class _PatchEmbedParams(nn.Module):
    def __init__(self, embed):
        super().__init__()
        self.proj = nn.Conv2d(3, embed, kernel_size=4, stride=4)
        self.norm = nn.LayerNorm(embed)

    forward = _no_forward


class SwinV2Params(nn.Module):
    """``pretrained.model``: timm SwinTransformerV2 attribute tree (weights only)."""

    def __init__(self, backbone):
        super().__init__()
        (self.timm_name, self.img_size, self.window, self.embed_dim, self.depths, self.heads,
         self.pretrained_windows, self.hooks) = SWIN_CONFIGS[backbone]
        E, g = self.embed_dim, self.img_size // 4
        self.patch_grid = (g, g)
        self.patch_embed = _PatchEmbedParams(E)
        self.layers = nn.ModuleList([
            _LayerParams(E * 2 ** i, (g // 2 ** i, g // 2 ** i), self.depths[i], self.heads[i], self.window,
                         i < len(self.depths) - 1, self.pretrained_windows[i])
            for i in range(len(self.depths))])
        self.num_features = E * 2 ** (len(self.depths) - 1)
        self.norm = nn.LayerNorm(self.num_features)       # computed-and-discarded by the reference (utils.py:65)
        self.head = nn.Linear(self.num_features, 1000)    # never called; present in timm's state_dict
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        for layer in self.layers:                          # timm's _init_respostnorm
            for blk in layer.blocks:
                for n in (blk.norm1, blk.norm2):
                    nn.init.constant_(n.weight, 0)
                    nn.init.constant_(n.bias, 0)

    forward = _no_forward


def relative_position_bias_table(attn, window, pretrained_window):
    """16*sigmoid(cpb_mlp(log-spaced coords table)) -> (heads, (2*window-1)**2) fp32: timm's continuous
    relative-position bias BEFORE the relative_position_index expansion.  bias[h, i, j] of timm's
    WindowAttention.forward is table[h, (qy-ky+window-1)*(2*window-1) + (qx-kx+window-1)] for query i=(qy,qx),
    key j=(ky,kx).  Input independent: evaluated once when weights are packed."""
    dev = attn.qkv.weight.device
    rh = torch.arange(-(window - 1), window, dtype=torch.float32)
    table = torch.stack(torch.meshgrid([rh, rh], indexing="ij")).permute(1, 2, 0).contiguous().unsqueeze(0)
    table = table / ((pretrained_window - 1) if pretrained_window > 0 else (window - 1))
    table = table * 8
    table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
    heads = attn.logit_scale.shape[0]
    with torch.no_grad():
        tab = attn.cpb_mlp(table.to(dev)).view(-1, heads).float()       # ((2w-1)^2, heads)
        return (16 * torch.sigmoid(tab)).t().contiguous()
