"""Input pipeline of the hot path: the reference's CPU transform
(SOccDPT/model/loader.py:141-272 -> transforms.py:53-251: cv2 bicubic Resize to a multiple of 32, NormalizeImage(0.5, 0.5),
PrepareForNet) as ONE CUDA kernel on uint8 HWC frames that are already on the device (`csrc/preprocess.cu`).

    transform, net_w, net_h = load_gpu_transforms("dpt_swin2_tiny_256")
    x = transform(frames_u8)            # frames_u8: CUDA uint8 (B, H, W, 3) or (H, W, 3)  ->  fp32 (B, 3, h, w)

The frames are NOT divided by 255 (neither are the reference's: bengaluru_driving_dataset.py:118-128), so x = 2 * resized - 1.
"""
import math

import torch

from . import _cabi

# (net_w, net_h) exactly as the reference returns them (loader.py:177-199): 256x256 even for swin2_base_384
INPUT_SIZES = {"dpt_swin2_base_384": (256, 256), "dpt_swin2_tiny_256": (256, 256), "dpt_hybrid_384": (384, 384)}
# keep_aspect_ratio as set per model type (loader.py: the Swin types force False, the others use `not square`)
_FORCE_NO_ASPECT = {"dpt_swin2_base_384", "dpt_swin2_tiny_256"}


def _constrain(x, multiple_of, min_val=0, max_val=None):
    # transforms.py:105-118 (np.round rounds half to even, like Python's round on floats)
    y = int(round(x / multiple_of) * multiple_of)
    if max_val is not None and y > max_val:
        y = int(math.floor(x / multiple_of) * multiple_of)
    if y < min_val:
        y = int(math.ceil(x / multiple_of) * multiple_of)
    return y


def get_size(width, height, net_w, net_h, keep_aspect_ratio, multiple_of=32):
    """Resize.get_size for resize_method='minimal' (transforms.py:120-177): output (width, height) for a frame."""
    scale_h, scale_w = net_h / height, net_w / width
    if keep_aspect_ratio:
        if abs(1 - scale_w) < abs(1 - scale_h):
            scale_h = scale_w
        else:
            scale_w = scale_h
    return _constrain(scale_w * width, multiple_of), _constrain(scale_h * height, multiple_of)


def transform_config(model_type, height=0, square=False):
    if model_type not in INPUT_SIZES:
        print(f"model_type '{model_type}' not implemented")
        assert False
    net_w, net_h = INPUT_SIZES[model_type]
    keep = False if model_type in _FORCE_NO_ASPECT else (not square)
    if height != 0:
        net_w, net_h = height, height
    return net_w, net_h, keep


class GpuTransform:
    """uint8 HWC CUDA frames -> the network's fp32 NCHW input.  No CPU fallback: CPU tensors raise."""

    def __init__(self, net_w, net_h, keep_aspect_ratio):
        self.net_w, self.net_h, self.keep_aspect_ratio = net_w, net_h, keep_aspect_ratio
        self._ws = {}

    def output_size(self, height, width):
        w, h = get_size(width, height, self.net_w, self.net_h, self.keep_aspect_ratio)
        return h, w

    def __call__(self, frames, out=None):
        if isinstance(frames, dict):           # the reference's sample-dict protocol
            res = dict(frames)
            res["image"] = self(frames["image"])[0]
            return res
        if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8):
            raise RuntimeError("GpuTransform needs a CUDA uint8 tensor of HWC frames (there is no CPU fallback; "
                               "use load_transforms for the reference's CPU transform)")
        if frames.dim() == 3:
            frames = frames.unsqueeze(0)
        assert frames.dim() == 4 and frames.shape[-1] == 3, "frames must be (B, H, W, 3)"
        frames = frames.contiguous()
        B, H, W, _ = frames.shape
        dh, dw = self.output_size(H, W)
        if out is None:
            out = torch.empty((B, 3, dh, dw), dtype=torch.float32, device=frames.device)
        assert out.shape == (B, 3, dh, dw) and out.dtype == torch.float32 and out.is_contiguous()
        lib = _cabi.load()
        key = (dh, dw, frames.device)
        if key not in self._ws:
            self._ws[key] = torch.empty(int(lib.soccdpt_preprocess_workspace_bytes(dh, dw)), dtype=torch.uint8,
                                        device=frames.device)
        ws = self._ws[key]
        with torch.cuda.device(frames.device):
            _cabi.check(lib.soccdpt_preprocess_fwd(frames.data_ptr(), B, H, W, 3, out.data_ptr(), dh, dw, ws.data_ptr(),
                                                   ws.numel(), _cabi.current_stream()), "preprocess")
        return out


def load_gpu_transforms(model_type="dpt_swin2_tiny_256", height=0, square=False):
    """GPU counterpart of load_transforms: (transform, net_w, net_h), same arguments and sizes."""
    net_w, net_h, keep = transform_config(model_type, height, square)
    return GpuTransform(net_w, net_h, keep), net_w, net_h
