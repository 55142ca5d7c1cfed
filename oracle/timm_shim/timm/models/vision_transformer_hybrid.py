"""TEST INFRASTRUCTURE ONLY -- restatement of the published algorithm of timm==0.6.12's
``vit_base_resnet50_384`` (= ``vit_base_r50_s16_384``): ``timm/models/vision_transformer_hybrid.py``,
``vision_transformer.py``, ``resnetv2.py``, ``layers/std_conv.py``, ``layers/norm_act.py``,
``layers/pool2d_same.py`` -- the pinned third-party dependency behind the reference's
``SOccDPT/model/backbones/vit.py:245-258`` (module tree and parameter names as in SURVEY.md Appendix A.2).

Parity status: "unpinned" by the reference (no weights / vectors shipped; its own constructor for this
model raises NameError); restated from the published timm source from memory.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _same_pad(x, k, s, value=0.0):
    ih, iw = x.shape[-2:]
    ph = max((math.ceil(ih / s) - 1) * s + k - ih, 0)
    pw = max((math.ceil(iw / s) - 1) * s + k - iw, 0)
    if ph > 0 or pw > 0:
        x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2], value=value)
    return x


class StdConv2dSame(nn.Conv2d):
    """Weight-standardised conv with TF 'SAME' dynamic padding (eps 1e-8 in this model)."""

    def __init__(self, in_chs, out_chs, kernel_size, stride=1, eps=1e-8):
        super().__init__(in_chs, out_chs, kernel_size, stride=stride, padding=0, bias=False)
        self.eps = eps

    def forward(self, x):
        x = _same_pad(x, self.kernel_size[0], self.stride[0])
        w = F.batch_norm(self.weight.reshape(1, self.out_channels, -1), None, None, training=True, momentum=0.0,
                         eps=self.eps).reshape_as(self.weight)
        return F.conv2d(x, w, None, self.stride, (0, 0))


class GroupNormAct(nn.GroupNorm):
    def __init__(self, num_channels, num_groups=32, eps=1e-5, apply_act=True):
        super().__init__(num_groups, num_channels, eps=eps)
        self.act = nn.ReLU(inplace=True) if apply_act else nn.Identity()

    def forward(self, x):
        return self.act(F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps))


class MaxPool2dSame(nn.MaxPool2d):
    def __init__(self, kernel_size, stride):
        super().__init__(kernel_size, stride, (0, 0))

    def forward(self, x):
        x = _same_pad(x, self.kernel_size, self.stride, value=-float("inf"))
        return F.max_pool2d(x, self.kernel_size, self.stride, (0, 0))


class DownsampleConv(nn.Module):
    def __init__(self, in_chs, out_chs, stride):
        super().__init__()
        self.conv = StdConv2dSame(in_chs, out_chs, 1, stride=stride)
        self.norm = GroupNormAct(out_chs, apply_act=False)

    def forward(self, x):
        return self.norm(self.conv(x))


class Bottleneck(nn.Module):
    """Non pre-activation bottleneck (ResNetV2 preact=False)."""

    def __init__(self, in_chs, out_chs, stride, has_proj):
        super().__init__()
        mid = out_chs // 4
        self.downsample = DownsampleConv(in_chs, out_chs, stride) if has_proj else None
        self.conv1 = StdConv2dSame(in_chs, mid, 1)
        self.norm1 = GroupNormAct(mid)
        self.conv2 = StdConv2dSame(mid, mid, 3, stride=stride)
        self.norm2 = GroupNormAct(mid)
        self.conv3 = StdConv2dSame(mid, out_chs, 1)
        self.norm3 = GroupNormAct(out_chs, apply_act=False)
        self.drop_path = nn.Identity()
        self.act3 = nn.ReLU(inplace=True)

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        x = self.norm1(self.conv1(x))
        x = self.norm2(self.conv2(x))
        x = self.norm3(self.conv3(x))
        return self.act3(self.drop_path(x) + shortcut)


class ResNetStage(nn.Module):
    def __init__(self, in_chs, out_chs, stride, depth):
        super().__init__()
        self.blocks = nn.Sequential(*[
            Bottleneck(in_chs if i == 0 else out_chs, out_chs, stride if i == 0 else 1, i == 0) for i in range(depth)])

    def forward(self, x):
        return self.blocks(x)


class ResNetV2(nn.Module):
    def __init__(self, layers=(3, 4, 9), channels=(256, 512, 1024), stem_chs=64):
        super().__init__()
        self.stem = nn.Sequential()
        self.stem.add_module("conv", StdConv2dSame(3, stem_chs, 7, stride=2))
        self.stem.add_module("norm", GroupNormAct(stem_chs))
        self.stem.add_module("pool", MaxPool2dSame(3, 2))
        stages, prev = [], stem_chs
        for i, (d, c) in enumerate(zip(layers, channels)):
            stages.append(ResNetStage(prev, c, 1 if i == 0 else 2, d))
            prev = c
        self.stages = nn.Sequential(*stages)
        self.num_features = prev
        self.norm = nn.Identity()
        self.head = nn.Identity()

    def forward_features(self, x):
        return self.norm(self.stages(self.stem(x)))

    def forward(self, x):
        return self.forward_features(x)


class HybridEmbed(nn.Module):
    def __init__(self, backbone, img_size=384, patch_size=1, feature_size=24, embed_dim=768):
        super().__init__()
        self.backbone = backbone
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.grid_size = (feature_size, feature_size)
        self.num_patches = feature_size * feature_size
        self.proj = nn.Conv2d(backbone.num_features, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        x = self.backbone(x)
        if isinstance(x, (list, tuple)):
            x = x[-1]
        return self.proj(x).flatten(2).transpose(1, 2)


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        x = (self.attn_drop(attn) @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.fc2 = nn.Linear(hidden, dim)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = nn.Identity()
        self.drop_path2 = nn.Identity()

    def forward(self, x):
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x))))
        x = x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))
        return x


class VisionTransformer(nn.Module):
    def __init__(self, backbone, img_size=384, embed_dim=768, depth=12, num_heads=12, num_classes=1000):
        super().__init__()
        self.num_classes, self.embed_dim, self.num_features = num_classes, embed_dim, embed_dim
        self.num_prefix_tokens = 1
        self.no_embed_class = False
        self.patch_embed = HybridEmbed(backbone, img_size=img_size, patch_size=1, feature_size=img_size // 16,
                                       embed_dim=embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches + 1, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(0.0)
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head = nn.Linear(embed_dim, num_classes)

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        return self.norm(self.blocks(self.pos_drop(x)))

    def forward(self, x):
        return self.head(self.fc_norm(self.forward_features(x)[:, 0]))


def vit_base_resnet50_384(**kwargs):
    return VisionTransformer(ResNetV2(layers=(3, 4, 9)), img_size=384, embed_dim=768, depth=12, num_heads=12)
