from . import layers, beit, swin_transformer_v2, vision_transformer_hybrid  # noqa: F401
