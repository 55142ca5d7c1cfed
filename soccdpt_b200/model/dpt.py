"""Parameter containers mirroring the reference's DPT decoder (SOccDPT/model/dpt.py:30-232,
SOccDPT/model/blocks.py:139-193,348-497).  Weights only -- the arithmetic is in the CUDA kernels."""
import torch.nn as nn

from .base_model import BaseModel
from .encoder import SWIN_CONFIGS, SwinV2Params, _no_forward
from .vit_hybrid import VIT_HYBRID_CONFIGS, HybridPretrained

# backbone -> stage channels fed to scratch.layerN_rn (blocks.py:64-78)
BACKBONE_CHANNELS = {"swin2t16_256": (96, 192, 384, 768), "swin2b24_384": (128, 256, 512, 1024),
                     "vitb_rn50_384": (256, 512, 768, 768)}


class _RCUParams(nn.Module):
    def __init__(self, f):
        super().__init__()
        self.conv1 = nn.Conv2d(f, f, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(f, f, 3, 1, 1, bias=True)

    forward = _no_forward


class _FusionParams(nn.Module):
    def __init__(self, f):
        super().__init__()
        self.out_conv = nn.Conv2d(f, f, 1, 1, 0, bias=True)
        self.resConfUnit1 = _RCUParams(f)
        self.resConfUnit2 = _RCUParams(f)

    forward = _no_forward


class _Pretrained(nn.Module):
    """``pretrained``: holds ``.model`` like _make_swin_backbone's result (swin_common.py:12-54)."""

    def __init__(self, backbone):
        super().__init__()
        self.model = SwinV2Params(backbone)

    forward = _no_forward


class DPTDepthModel(BaseModel):
    """DPT + monocular-depth head (dpt.py:185-232); ``return_features`` is always on for SOccDPT_V3."""

    def __init__(self, path=None, non_negative=True, backbone="swin2t16_256", features=256, return_features=True,
                 **kwargs):
        super().__init__()
        assert backbone in SWIN_CONFIGS or backbone in VIT_HYBRID_CONFIGS, f"Backbone '{backbone}' not implemented"
        assert non_negative, "the fused depth-head epilogue implements non_negative=True (dpt.py:217)"
        self.backbone, self.features, self.return_features = backbone, features, return_features
        self.pretrained = _Pretrained(backbone) if backbone in SWIN_CONFIGS else HybridPretrained(backbone)
        ch = BACKBONE_CHANNELS[backbone]
        scratch = nn.Module()
        for i, c in enumerate(ch):
            setattr(scratch, f"layer{i + 1}_rn", nn.Conv2d(c, features, 3, 1, 1, bias=False))
        for i in range(1, 5):
            setattr(scratch, f"refinenet{i}", _FusionParams(features))
        scratch.output_conv = nn.Sequential(
            nn.Conv2d(features, features // 2, 3, 1, 1), nn.Identity(),
            nn.Conv2d(features // 2, 32, 3, 1, 1), nn.ReLU(True),
            nn.Conv2d(32, 1, 1, 1, 0), nn.ReLU(True), nn.Identity())
        self.scratch = scratch
        if path is not None:
            self.load_net(path)

    forward = _no_forward
