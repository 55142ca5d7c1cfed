from . import layers, beit, swin_transformer_v2  # noqa: F401
