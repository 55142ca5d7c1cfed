"""Parameter containers for the ViT-hybrid encoder of ``dpt_hybrid_384``: timm==0.6.12 ``vit_base_resnet50_384``
(ResNetV2-50 stem + stages (3, 4, 9) feeding a 12-block ViT-B) as wired by the reference's
``_make_vit_b_rn50_backbone`` (SOccDPT/model/backbones/vit.py:147-258: hooks [0, 1, 8, 11], readout "project",
features [256, 512, 768, 768]).  Attribute names follow timm / the reference exactly, so
``pretrained.model.*`` / ``pretrained.act_postprocess{3,4}.*`` checkpoints load.  Weights only: the arithmetic runs in
``soccdpt_b200.engine``.

The reference's own constructor for this model raises NameError (``value`` undefined, vit.py:233-242);
the module tree below is what that constructor builds once the discarded ``nn.Sequential`` is bound to ``value``.
"""
import torch
import torch.nn as nn

from .encoder import _no_forward

VIT_HYBRID_CONFIGS = {
    # backbone: (timm name, img, embed, depth, heads, resnet layers, resnet channels, hooks, tap channels)
    "vitb_rn50_384": ("vit_base_resnet50_384", 384, 768, 12, 12, (3, 4, 9), (256, 512, 1024), (0, 1, 8, 11),
                      (256, 512, 768, 768)),
}


class _StdConv(nn.Conv2d):
    """timm StdConv2dSame: weight-standardised (eps 1e-8), TF-"SAME" padding, no bias."""

    def __init__(self, cin, cout, k, stride=1):
        super().__init__(cin, cout, k, stride=stride, padding=0, bias=False)
        self.eps = 1e-8

    forward = _no_forward


class _GroupNorm(nn.GroupNorm):
    def __init__(self, c, apply_act=True):
        super().__init__(32, c, eps=1e-5)
        self.apply_act = apply_act

    forward = _no_forward


class _DownsampleParams(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = _StdConv(cin, cout, 1, stride)
        self.norm = _GroupNorm(cout, apply_act=False)

    forward = _no_forward


class _BottleneckParams(nn.Module):
    def __init__(self, cin, cout, stride, has_proj):
        super().__init__()
        mid = cout // 4
        self.stride = stride
        self.downsample = _DownsampleParams(cin, cout, stride) if has_proj else None
        self.conv1 = _StdConv(cin, mid, 1)
        self.norm1 = _GroupNorm(mid)
        self.conv2 = _StdConv(mid, mid, 3, stride)
        self.norm2 = _GroupNorm(mid)
        self.conv3 = _StdConv(mid, cout, 1)
        self.norm3 = _GroupNorm(cout, apply_act=False)

    forward = _no_forward


class _StageParams(nn.Module):
    def __init__(self, cin, cout, stride, depth):
        super().__init__()
        self.blocks = nn.Sequential(*[
            _BottleneckParams(cin if i == 0 else cout, cout, stride if i == 0 else 1, i == 0) for i in range(depth)])

    forward = _no_forward


class _ResNetV2Params(nn.Module):
    def __init__(self, layers, channels, stem=64):
        super().__init__()
        self.stem = nn.Sequential()
        self.stem.add_module("conv", _StdConv(3, stem, 7, 2))
        self.stem.add_module("norm", _GroupNorm(stem))
        self.stem.add_module("pool", nn.Identity())        # MaxPool2dSame(3, 2): no parameters
        stages, prev = [], stem
        for i, (d, c) in enumerate(zip(layers, channels)):
            stages.append(_StageParams(prev, c, 1 if i == 0 else 2, d))
            prev = c
        self.stages = nn.Sequential(*stages)
        self.num_features = prev

    forward = _no_forward


class _HybridEmbedParams(nn.Module):
    def __init__(self, backbone, embed):
        super().__init__()
        self.backbone = backbone
        self.proj = nn.Conv2d(backbone.num_features, embed, kernel_size=1, stride=1)

    forward = _no_forward


class _ViTAttentionParams(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    forward = _no_forward


class _ViTMlpParams(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    forward = _no_forward


class _ViTBlockParams(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _ViTAttentionParams(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _ViTMlpParams(dim, dim * 4)

    forward = _no_forward


class ViTHybridParams(nn.Module):
    """``pretrained.model``: timm VisionTransformer with a HybridEmbed(ResNetV2) patch embedding (weights only)."""

    def __init__(self, backbone):
        super().__init__()
        (self.timm_name, self.img_size, self.embed_dim, self.depth, self.num_heads, layers, channels, self.hooks,
         self.tap_channels) = VIT_HYBRID_CONFIGS[backbone]
        E, g = self.embed_dim, self.img_size // 16
        self.patch_grid = (g, g)
        self.patch_size = [16, 16]             # vit.py:170
        self.start_index = 1                   # vit.py:158
        self.cls_token = nn.Parameter(torch.zeros(1, 1, E))
        self.pos_embed = nn.Parameter(torch.randn(1, g * g + 1, E) * 0.02)
        self.patch_embed = _HybridEmbedParams(_ResNetV2Params(layers, channels), E)
        self.blocks = nn.Sequential(*[_ViTBlockParams(E, self.num_heads) for _ in range(self.depth)])
        self.norm = nn.LayerNorm(E, eps=1e-6)   # computed and discarded by forward_flex (vit.py:82)
        self.head = nn.Linear(E, 1000)          # never called; present in timm's state_dict
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    forward = _no_forward


class _ProjectReadoutParams(nn.Module):
    """reference backbones/utils.py:27-40: project = Sequential(Linear(2*in, in), GELU)."""

    def __init__(self, dim):
        super().__init__()
        self.start_index = 1
        self.project = nn.Sequential(nn.Linear(2 * dim, dim), nn.GELU())

    forward = _no_forward


class HybridPretrained(nn.Module):
    """``pretrained``: ``.model`` + ``act_postprocess1..4`` as in _make_vit_b_rn50_backbone (vit.py:179-219, hybrid branch:
    taps 1 and 2 come straight from the ResNet stages, so their post-processing is three Identities)."""

    def __init__(self, backbone):
        super().__init__()
        self.model = ViTHybridParams(backbone)
        E, g = self.model.embed_dim, self.model.patch_grid
        f = self.model.tap_channels
        self.act_postprocess1 = nn.Sequential(nn.Identity(), nn.Identity(), nn.Identity())
        self.act_postprocess2 = nn.Sequential(nn.Identity(), nn.Identity(), nn.Identity())
        self.act_postprocess3 = nn.Sequential(
            _ProjectReadoutParams(E), nn.Identity(), nn.Unflatten(2, torch.Size(g)),       # readout, Transpose(1, 2), Unflatten
            nn.Conv2d(E, f[2], kernel_size=1, stride=1, padding=0))
        self.act_postprocess4 = nn.Sequential(
            _ProjectReadoutParams(E), nn.Identity(), nn.Unflatten(2, torch.Size(g)),
            nn.Conv2d(E, f[3], kernel_size=1, stride=1, padding=0),
            nn.Conv2d(f[3], f[3], kernel_size=3, stride=2, padding=1))

    forward = _no_forward
