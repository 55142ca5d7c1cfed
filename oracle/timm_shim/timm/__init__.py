"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for ``timm==0.6.12``.

The reference (AdityaNG/SOccDPT) takes its encoder arithmetic from
``timm==0.6.12`` (reference ``requirements.txt:12``), which is neither vendored
under /root/reference nor installed in this image.  This package restates the
published SwinV2 algorithm of that pinned version (module tree and parameter
names as in SURVEY.md Appendix A.1) so that the reference's own call sites

    SOccDPT/model/backbones/swin2.py:7,16,25   timm.create_model("swinv2_*")
    SOccDPT/model/dpt.py:3                      timm.models.layers.get_act_layer
    SOccDPT/model/backbones/beit.py:8           timm.models.beit.gen_relative_position_index

import and run unmodified.  It is put on ``sys.path`` only by ``oracle/`` code
and by tests; the product package ``soccdpt_b200`` never imports it.

Parity status of the encoder: "parity unpinned" by the reference itself (it has
no tests, no golden vectors and ships no weights); the restatement is
cross-checked against HuggingFace ``transformers.Swinv2Model`` in
``tests/test_oracle_swinv2_vs_hf.py``.
"""
from .models.swin_transformer_v2 import SwinTransformerV2
from .models.vision_transformer_hybrid import vit_base_resnet50_384

__version__ = "0.6.12+soccdpt-oracle-shim"

_SWINV2_CFGS = {
    # name: (img_size, window, embed_dim, depths, heads, pretrained_window_sizes)
    "swinv2_tiny_window16_256": (256, 16, 96, (2, 2, 6, 2), (3, 6, 12, 24), (0, 0, 0, 0)),
    "swinv2_tiny_window8_256": (256, 8, 96, (2, 2, 6, 2), (3, 6, 12, 24), (0, 0, 0, 0)),
    "swinv2_base_window12to24_192to384_22kft1k": (
        384, 24, 128, (2, 2, 18, 2), (4, 8, 16, 32), (12, 12, 12, 6)),
    "swinv2_large_window12to24_192to384_22kft1k": (
        384, 24, 192, (2, 2, 18, 2), (6, 12, 24, 48), (12, 12, 12, 6)),
}


def create_model(model_name, pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("timm shim: no network, pretrained weights unavailable")
    if model_name in _SWINV2_CFGS:
        img, ws, dim, depths, heads, pws = _SWINV2_CFGS[model_name]
        return SwinTransformerV2(
            img_size=img, window_size=ws, embed_dim=dim, depths=depths,
            num_heads=heads, pretrained_window_sizes=pws, **kwargs)
    if model_name in ("vit_base_resnet50_384", "vit_base_r50_s16_384"):
        return vit_base_resnet50_384(**kwargs)
    raise RuntimeError(f"timm shim: model '{model_name}' is not restated")
