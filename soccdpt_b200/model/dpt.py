"""Parameter containers mirroring the reference's DPT decoder (SOccDPT/model/dpt.py:30-232,
SOccDPT/model/blocks.py:139-193,348-497).  Weights only -- the arithmetic is in the CUDA kernels."""
import torch.nn as nn

from .base_model import BaseModel
from .encoder import SWIN_CONFIGS, SwinV2Params, _no_forward
from .vit_hybrid import VIT_HYBRID_CONFIGS, HybridPretrained

# backbone -> stage channels fed to scratch.layerN_rn (blocks.py:64-78)
BACKBONE_CHANNELS = {"swin2t16_256": (96, 192, 384, 768), "swin2b24_384": (128, 256, 512, 1024),
                     "vitb_rn50_384": (256, 512, 768, 768)}


class Interpolate(nn.Module):
    """Marker for the x2 bilinear (align_corners=True) stage of the heads (reference blocks.py:239-273)."""

    def __init__(self, scale_factor, mode, align_corners=False):
        super().__init__()
        self.scale_factor, self.mode, self.align_corners = scale_factor, mode, align_corners


class _RCUParams(nn.Module):
    def __init__(self, f, bn=False):
        super().__init__()
        self.bn = bn
        self.conv1 = nn.Conv2d(f, f, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(f, f, 3, 1, 1, bias=True)
        if bn:      # blocks.py:381-383 (``use_bn``: DPTSegmentationModel only); folded into the convs at pack time
            self.bn1 = nn.BatchNorm2d(f)
            self.bn2 = nn.BatchNorm2d(f)

    forward = _no_forward


class _FusionParams(nn.Module):
    def __init__(self, f, bn=False):
        super().__init__()
        self.out_conv = nn.Conv2d(f, f, 1, 1, 0, bias=True)
        self.resConfUnit1 = _RCUParams(f, bn)
        self.resConfUnit2 = _RCUParams(f, bn)

    forward = _no_forward


class _Pretrained(nn.Module):
    """``pretrained``: holds ``.model`` like _make_swin_backbone's result (swin_common.py:12-54)."""

    def __init__(self, backbone):
        super().__init__()
        self.model = SwinV2Params(backbone)

    forward = _no_forward


class DPT(BaseModel):
    """Encoder + reassemble + fusion decoder parameters around a ``head`` Sequential stored as
    ``scratch.output_conv`` (dpt.py:30-140)."""

    def __init__(self, head, features=256, backbone="swin2t16_256", use_bn=False, return_features=False, **kwargs):
        super().__init__()
        assert backbone in SWIN_CONFIGS or backbone in VIT_HYBRID_CONFIGS, f"Backbone '{backbone}' not implemented"
        self.backbone, self.features, self.return_features = backbone, features, return_features
        self.pretrained = _Pretrained(backbone) if backbone in SWIN_CONFIGS else HybridPretrained(backbone)
        ch = BACKBONE_CHANNELS[backbone]
        scratch = nn.Module()
        for i, c in enumerate(ch):
            setattr(scratch, f"layer{i + 1}_rn", nn.Conv2d(c, features, 3, 1, 1, bias=False))
        for i in range(1, 5):
            setattr(scratch, f"refinenet{i}", _FusionParams(features, use_bn))
        scratch.output_conv = head
        self.scratch = scratch

    forward = _no_forward


class DPTDepthModel(DPT):
    """DPT + monocular-depth head (dpt.py:185-232); SOccDPT_V3 runs it with ``return_features`` on."""

    def __init__(self, path=None, non_negative=True, backbone="swin2t16_256", features=256, return_features=True,
                 **kwargs):
        assert non_negative, "the fused depth-head epilogue implements non_negative=True (dpt.py:217)"
        head = nn.Sequential(
            nn.Conv2d(features, features // 2, 3, 1, 1), nn.Identity(),
            nn.Conv2d(features // 2, 32, 3, 1, 1), nn.ReLU(True),
            nn.Conv2d(32, 1, 1, 1, 0), nn.ReLU(True), nn.Identity())
        super().__init__(head, features=features, backbone=backbone, return_features=return_features, **kwargs)
        if path is not None:
            self.load_net(path)


class DPTSegmentationModel(DPT):
    """DPT + segmentation head, BatchNorm in every residual conv unit (dpt.py:235-272).  ``auxlayer`` only owns its
    weights: the reference's forward never evaluates it either."""

    def __init__(self, num_classes=3, path=None, backbone="swin2t16_256", features=256, **kwargs):
        kwargs["use_bn"] = True

        def seg_layers(tail):
            return [nn.Conv2d(features, features, 3, padding=1, bias=False), nn.BatchNorm2d(features), nn.ReLU(True),
                    nn.Dropout(0.1, False), nn.Conv2d(features, num_classes, 1)] + tail

        head = nn.Sequential(*seg_layers([Interpolate(scale_factor=2, mode="bilinear", align_corners=True), nn.Sigmoid()]))
        super().__init__(head, features=features, backbone=backbone, **kwargs)
        self.auxlayer = nn.Sequential(*seg_layers([]))
        if path is not None:
            self.load_net(path)
