"""Mirror of the reference's loader (SOccDPT/model/loader.py:13-138,141-272): ``model_type`` -> backbone
wiring and the network input sizes.  Only the model types on the accelerated path are constructible."""
from typing import Type

import torch

from .base_model import BaseModel

_BACKBONES = {
    "dpt_swin2_base_384": "swin2b24_384",
    "dpt_swin2_tiny_256": "swin2t16_256",
    "dpt_hybrid_384": "vitb_rn50_384",
}


def load_model(arch, model_kwargs: dict, device: torch.device, model_path: str, model_type: str = "dpt_large_384",
               optimize: bool = False) -> Type[BaseModel]:
    assert issubclass(arch, BaseModel), f"arch '{arch}' not implemented, must be a soccdpt_b200 BaseModel"
    if model_type not in _BACKBONES:
        print(f"model_type '{model_type}' not implemented")
        assert False
    model = arch(path=model_path, backbone=_BACKBONES[model_type], **model_kwargs)
    print("Model loaded, number of parameters = {:.0f}M".format(sum(p.numel() for p in model.parameters()) / 1e6))
    # `optimize` (half + channels_last in the reference, loader.py:132-134) is a no-op here: the kernels
    # already run bf16 tensor-core math on NHWC activations while the module keeps fp32 master weights.
    model.to(device)
    return model


def load_transforms(model_type: str = "dpt_large_384", height: int = 0, square: bool = False):
    """Returns (transform, net_w, net_h).  The transform is the reference's CPU pre-processing
    (cv2 bicubic Resize to a multiple of 32 -> NormalizeImage(0.5,0.5) -> PrepareForNet, transforms.py:53-251) restated on
    numpy/cv2 for callers that hold host frames; `soccdpt_b200.preprocess.load_gpu_transforms` is the same
    transform as one CUDA kernel on device-resident uint8 frames."""
    from ..preprocess import get_size, transform_config
    net_w, net_h, keep_aspect_ratio = transform_config(model_type, height, square)

    def transform(sample):
        import cv2
        import numpy as np
        w, h = get_size(sample["image"].shape[1], sample["image"].shape[0], net_w, net_h, keep_aspect_ratio)
        img = cv2.resize(sample["image"], (w, h), interpolation=cv2.INTER_CUBIC)
        img = (img - np.array([0.5, 0.5, 0.5])) / np.array([0.5, 0.5, 0.5])
        out = dict(sample)
        out["image"] = np.ascontiguousarray(np.transpose(img, (2, 0, 1))).astype(np.float32)
        return out

    return transform, net_w, net_h
