"""soccdpt_b200 -- B200 (sm_100a) implementation of SOccDPT's inference hot path behind the
reference's Python model API.  See DESIGN.md / INTEGRATION.md."""
from .model import (  # noqa: F401
    BaseModel, DepthNet, SegNet, SOccDPT, SOccDPT_V1, SOccDPT_V3, SOccDPT_versions, default_depth_models,
    default_seg_models, load_model, load_transforms, model_types)

__all__ = ["BaseModel", "DepthNet", "SegNet", "SOccDPT", "SOccDPT_V1", "SOccDPT_V3", "SOccDPT_versions", "default_depth_models", "default_seg_models",
           "load_model", "load_transforms", "model_types"]
