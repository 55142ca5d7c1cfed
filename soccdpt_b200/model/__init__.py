"""Host-side mirror of the reference's ``SOccDPT.model`` package (same public names)."""
from .base_model import BaseModel  # noqa: F401
from .loader import load_model, load_transforms  # noqa: F401
from .SOccDPT import (  # noqa: F401
    DepthNet, SegNet, SOccDPT, SOccDPT_V1, SOccDPT_V3, SOccDPT_versions, default_depth_models, default_seg_models,
    model_types)
