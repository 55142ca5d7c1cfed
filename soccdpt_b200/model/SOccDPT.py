"""Drop-in mirror of the reference's model API (SOccDPT/model/SOccDPT.py): same class names, constructor
keywords, output tuple, attribute names and state_dict keys -- with the arithmetic on B200 kernels.

    SOccDPT        base class: geometry constants + get_semantic_occupancy      (SOccDPT.py:133-463)
    SOccDPT_V1     two complete DPT networks (depth, segmentation) on one voxeliser  (SOccDPT.py:470-523)
    SOccDPT_V3     DPT depth net with return_features + segmentation head        (SOccDPT.py:626-685)
    DepthNet/SegNet adaptors picking tuple element 0 / 1                         (SOccDPT.py:697-724)

Deliberately preserved quirks of the reference: the returned inverse depth is clamped at
1e-8; points #0,#1,#2 of every frame carry pc_scale/pc_shift; the occupancy grid is binary and, in the
default ``occupancy_mode="reference_union"``, is the OR over the batch written to every batch element;
planes i=0/j=0/k=0 are never filled; at batch 1 the segmentation output loses its batch dim.
There is no CPU path: calling the model on CPU tensors raises.
"""
import os
from typing import Type

import numpy as np
import torch
import torch.nn as nn

from .. import _cabi
from ..engine import NetworkEngine
from ..geometry import load_calib, make_geometry
from .base_model import BaseModel
from .dpt import DPTDepthModel, DPTSegmentationModel, Interpolate

cpu_device = torch.device("cpu")

# reference: SOccDPT/datasets/bdd_helper.py:53-56 (not shipped with either repo)
DEFAULT_CALIB = os.path.join("~", "Datasets", "Depth_Dataset_Bengaluru", "calibration", "pocoX3", "calib.yaml")

default_depth_models = {
    "dpt_swin2_base_384": "weights/dpt_swin2_base_384.pt",
    "dpt_swin2_tiny_256": "weights/dpt_swin2_tiny_256.pt",
    "dpt_hybrid_384": "weights/dpt_hybrid_384.pt",
}
model_types = default_depth_models.keys()

# SOccDPT.py:43-55: no segmentation checkpoints are published for any model type
default_seg_models = {k: None for k in default_depth_models}

DEPTH_l39icv3q = "checkpoints_pretrained/depth_dpt_hybrid/l39icv3q/checkpoint_epoch15.pth"  # reference defaults
SEG_wrlnq5jb = "checkpoints_pretrained/seg_dpt_hybrid/wrlnq5jb/checkpoint_epoch15.pth"


class ScaledTanh(nn.Module):
    """0.5*tanh(x)+0.5 (reference scaled_tanh.py:8-10); marker module, evaluated in seg_finish_kernel."""


class SOccDPT(BaseModel):
    def __init__(self, model_type="dpt_swin2_tiny_256", backbone="swin2t16_256", path=None, num_classes: int = 3,
                 camera_intrinsics_yaml=DEFAULT_CALIB, point_compute_method="torch",
                 grid_size=(256, 256, 32), scale=(2.0, 2.0, 0.666), shift=(0.0, 0.0, 0.0),
                 pc_scale=(10000.0, 50000.0, 800.0), pc_shift=(55.0, -20.0, 15.0), correction_angle=(7.0, 0, 0),
                 compute_occ=False, occupancy_mode="reference_union", occupancy_output="dense", device_convention="cpu",
                 **kwargs):
        super(SOccDPT, self).__init__(**kwargs)
        self.compute_occ = compute_occ
        self.grid_size, self.scale, self.shift = grid_size, scale, shift
        self.pc_scale, self.pc_shift, self.correction_angle = pc_scale, pc_shift, correction_angle
        self.backbone, self.model_type, self.path = backbone, model_type, path
        self.num_classes = num_classes
        # SOccDPT.py:175-181: occupancy grid size in metres, fp32
        self.occupancy_shape = np.array([float(grid_size[i] / scale[i]) for i in range(len(grid_size))], dtype=np.float32)
        self.features = 256
        assert point_compute_method in ("torch", "numpy")
        self.point_compute_method = point_compute_method
        assert occupancy_mode in ("reference_union", "per_frame")
        self.occupancy_mode = occupancy_mode
        # "dense": the reference's (B,G0,G1,G2,C) fp32 grid.  "packed" (extension): the 4th output is the
        # voxeliser's bit-packed mask instead -- int32 words, (mask_words,) or (B, mask_words) in per_frame mode; see
        # soccdpt_b200.occupancy.packed_to_points and SOCCDPT_OCC_PACKED in include/soccdpt_b200.h
        assert occupancy_output in ("dense", "packed")
        self.occupancy_output = occupancy_output
        # The reference's eager post-processing is device dependent in the last bits of X / Y (ATen divides a tensor by a host
        # scalar on the CPU and multiplies by the scalar's fp32 reciprocal on CUDA; cos / sin of the rotation come from libm or
        # from CUDA): "cpu" (default) is bit-exact against the reference run on CPU tensors -- the convention of its fixtures
        # --, "cuda" against the reference run on a CUDA device (tests/test_gpu_device_convention.py).
        assert device_convention in ("cpu", "cuda")
        self.device_convention = device_convention
        self._geom_cache = {}

        self.camera_intrinsics_yaml = os.path.expanduser(camera_intrinsics_yaml)
        self.cam_settings = load_calib(self.camera_intrinsics_yaml)
        cs = self.cam_settings
        self.DistCoef = np.array([cs["Camera.k1"], cs["Camera.k2"], cs["Camera.p1"], cs["Camera.p2"], cs.get("Camera.k3", 0)])
        self.intrinsic_matrix = np.array([[cs["Camera.fx"], 0.0, cs["Camera.cx"]], [0.0, cs["Camera.fy"], cs["Camera.cy"]],
                                          [0.0, 0.0, 1.0]])
        self.fx, self.fy = self.intrinsic_matrix[0, 0], self.intrinsic_matrix[1, 1]
        self.cx, self.cy = self.intrinsic_matrix[0, 2], self.intrinsic_matrix[1, 2]
        self.width, self.height = cs["Camera.width"], cs["Camera.height"]
        self.occupancy_conv = nn.Identity()
        self._workspaces = {}

    def forward(self, x: torch.Tensor):
        assert False, "Not implemented, take input batch and produce inv_depth, segmentation and call " \
                      "self.get_semantic_occupancy(inv_depth, segmentation)"

    # ------------------------------------------------------------------ A8 / A9
    def _geometry(self, device=None):
        key = (self.device_convention, _cabi.normalize_device(device).index if (device is not None and self.device_convention == "cuda") else -1,
               tuple(self.grid_size), tuple(float(v) for v in self.occupancy_shape), tuple(self.pc_scale), tuple(self.pc_shift),
               tuple(self.correction_angle), int(self.num_classes))
        g = self._geom_cache.get(key)
        if g is None:       # cached: the "cuda" convention evaluates cos / sin on the device (a host sync)
            g = make_geometry(self.fx, self.fy, self.cx, self.cy, self.height, self.width, self.num_classes,
                              self.grid_size, self.occupancy_shape, self.pc_scale, self.pc_shift, self.correction_angle,
                              self.device_convention, device)
            self._geom_cache = {key: g}
        return g

    def _workspace(self, geom, B, mode, device):
        lib = _cabi.load()
        need = int(lib.soccdpt_voxel_workspace_bytes(ctypes_byref(geom), B, mode))
        key = (_cabi.normalize_device(device).index, need)
        if key not in self._workspaces:
            self._workspaces[key] = torch.empty(max(need, 16), dtype=torch.uint8, device=device)
        return self._workspaces[key], need

    def get_semantic_occupancy(self, inv_depth, segmentation):
        """(B,h,w)|(B,1,h,w) inverse depth + (B,C,h,w) class scores -> the reference's 4-tuple
        (inv_depth_up, segmentation_up, points_batched, occupancy_grid | None); SOccDPT.py:264-372."""
        if not (inv_depth.is_cuda and segmentation.is_cuda):
            raise _cabi.SoccdptError("soccdpt_b200 runs on CUDA devices only (no CPU fallback)")
        lib = _cabi.load()
        if inv_depth.dim() == 4:
            inv_depth = inv_depth[:, 0]
        inv_depth = inv_depth.to(torch.float32).contiguous()
        segmentation = segmentation.to(torch.float32).contiguous()
        B, h, w = inv_depth.shape
        C = segmentation.shape[1]
        assert C == self.num_classes and segmentation.shape[0] == B and tuple(segmentation.shape[2:]) == (h, w)
        dev = inv_depth.device
        H, Wd = int(self.height), int(self.width)
        geom = self._geometry(dev)
        mode = _cabi.OCC_PER_FRAME if self.occupancy_mode == "per_frame" else _cabi.OCC_REFERENCE_UNION
        inv_up = torch.empty((B, H, Wd), dtype=torch.float32, device=dev)
        seg_up = torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev)
        points = torch.empty((B, H, Wd, 3), dtype=torch.float32, device=dev)
        grid = None
        packed = self.compute_occ and self.occupancy_output == "packed"
        if self.compute_occ and not packed:
            G = self.grid_size
            grid = torch.empty((B, G[0], G[1], G[2], C), dtype=torch.float32, device=dev)
        if packed:
            mode |= _cabi.OCC_PACKED
        ws, need = self._workspace(geom, B, mode, dev)      # voxel mask + resize tables
        with torch.cuda.device(dev):                        # the C ABI launches on the current device
            rc = lib.soccdpt_postprocess_fwd(
                _cabi.ptr(inv_depth), _cabi.ptr(segmentation), B, h, w, ctypes_byref(geom), _cabi.ptr(inv_up), _cabi.ptr(seg_up),
                _cabi.ptr(points), _cabi.ptr(grid), mode, _cabi.ptr(ws), need, _cabi.current_stream(dev))
        _cabi.check(rc, "soccdpt_postprocess_fwd")
        if packed:
            grid = self._packed_mask(ws, B)
        # the reference's .squeeze() calls (SOccDPT.py:276,282-285)
        seg_out = seg_up.squeeze()
        inv_out = inv_up.squeeze()
        if inv_out.dim() == 2:
            inv_out = inv_out.unsqueeze(0)
        return inv_out, seg_out, points, grid

    def voxelize(self, inv_depth_up, segmentation_up):
        """Standalone voxeliser on maps already at camera resolution (BASELINE config 5): clamps
        ``inv_depth_up`` IN PLACE like the reference and returns (points, occupancy_grid | None)."""
        if not (inv_depth_up.is_cuda and segmentation_up.is_cuda):
            raise _cabi.SoccdptError("soccdpt_b200 runs on CUDA devices only (no CPU fallback)")
        lib = _cabi.load()
        assert inv_depth_up.dtype == torch.float32 and inv_depth_up.is_contiguous()
        seg = segmentation_up.to(torch.float32).contiguous()
        B, H, Wd = inv_depth_up.shape
        assert (H, Wd) == (int(self.height), int(self.width)) and seg.shape == (B, self.num_classes, H, Wd)
        dev = inv_depth_up.device
        geom = self._geometry(dev)
        mode = _cabi.OCC_PER_FRAME if self.occupancy_mode == "per_frame" else _cabi.OCC_REFERENCE_UNION
        points = torch.empty((B, H, Wd, 3), dtype=torch.float32, device=dev)
        grid, ws, need = None, None, 0
        packed = self.compute_occ and self.occupancy_output == "packed"
        if packed:
            mode |= _cabi.OCC_PACKED
        if self.compute_occ:
            G = self.grid_size
            if not packed:
                grid = torch.empty((B, G[0], G[1], G[2], self.num_classes), dtype=torch.float32, device=dev)
            ws, need = self._workspace(geom, B, mode, dev)
        with torch.cuda.device(dev):
            rc = lib.soccdpt_voxelize_fwd(_cabi.ptr(inv_depth_up), _cabi.ptr(seg), B, ctypes_byref(geom), _cabi.ptr(points),
                                          _cabi.ptr(grid), mode, _cabi.ptr(ws), need, _cabi.current_stream(dev))
        _cabi.check(rc, "soccdpt_voxelize_fwd")
        if packed:
            grid = self._packed_mask(ws, B)
        return points, grid

    def _packed_mask(self, ws, B):
        """copy of the bit-packed voxel mask the call left at the start of its workspace (int32 words)."""
        from ..occupancy import mask_words
        n = mask_words(self.grid_size)
        if self.occupancy_mode == "per_frame":
            return ws[: B * n * 4].view(torch.int32).view(B, n).clone()
        return ws[: n * 4].view(torch.int32).clone()


def ctypes_byref(s):
    import ctypes
    return ctypes.byref(s)


class _EngineOwner:
    """Weights changed (load_state_dict / .to()) -> the engines repack on the next forward."""

    def _engines(self):
        return [e for e in (getattr(self, "_engine", None), getattr(self, "_seg_engine", None)) if e is not None]

    def _invalidate(self):
        for e in self._engines():
            e.invalidate()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._invalidate()
        return out

    def _conv_impl(self, conv_impl):
        return conv_impl or os.environ.get("SOCCDPT_CONV_IMPL", "tcgen05")


class SOccDPT_V1(_EngineOwner, SOccDPT):
    """Two independent DPT networks -- ``depth_net`` (DPTDepthModel) and ``seg_net`` (DPTSegmentationModel, BatchNorm in its
    residual conv units) -- feeding the shared resize + unproject + voxelise stage (SOccDPT.py:470-523).  Each network is its
    own launch plan over the same kernels as SOccDPT_V3; the segmentation plan runs on a second CUDA stream so the two
    encoders / decoders overlap on the GPU."""

    def __init__(self, load_depth: str = DEPTH_l39icv3q, load_seg: str = SEG_wrlnq5jb, **kwargs):
        super(SOccDPT_V1, self).__init__(**kwargs)
        from .loader import load_model

        depth_model_weights = load_depth
        if depth_model_weights is None:
            depth_model_weights = default_depth_models[self.model_type]
        self.depth_net = load_model(DPTDepthModel, dict(non_negative=True, return_features=False), cpu_device,
                                    depth_model_weights, self.model_type)
        self.pretrained = self.depth_net.pretrained     # alias, as in the reference (SOccDPT.py:496)

        seg_model_weights = load_seg
        if seg_model_weights is None:
            seg_model_weights = default_seg_models[self.model_type]
        self.seg_net = load_model(DPTSegmentationModel, dict(num_classes=self.num_classes), cpu_device,
                                  seg_model_weights, self.model_type)
        self._engine = None
        self._seg_engine = None
        self._side_streams = {}
        self.load_net(self.path)

    def engine(self, conv_impl=None):
        """the depth network's engine (``seg_engine()`` is the other one)."""
        impl = self._conv_impl(conv_impl)
        if self._engine is None or (conv_impl is not None and self._engine.conv_impl != impl):
            self._engine = NetworkEngine(self, impl, dpt=self.depth_net, heads=("depth",))
        return self._engine

    def seg_engine(self, conv_impl=None):
        impl = self._conv_impl(conv_impl)
        if self._seg_engine is None or (conv_impl is not None and self._seg_engine.conv_impl != impl):
            self._seg_engine = NetworkEngine(self, impl, dpt=self.seg_net, seg_head=self.seg_net.scratch.output_conv,
                                             heads=("seg",))
        return self._seg_engine

    def network(self, x, conv_impl=None):
        """image -> (depth_net(x) (B,h,w) f32, seg_net(x) (B,C,h,w) f32), SOccDPT.py:519-520.  Static engine buffers."""
        if self.training:
            raise _cabi.SoccdptError("soccdpt_b200 implements the inference path only: call net.eval() first")
        if not x.is_cuda:
            raise _cabi.SoccdptError("soccdpt_b200 runs on CUDA devices only (no CPU fallback)")
        main = torch.cuda.current_stream(x.device)
        side = self._side_streams.get(x.device)
        if side is None:
            side = self._side_streams[x.device] = torch.cuda.Stream(device=x.device)
        side.wait_stream(main)                          # x is ready on the caller's stream
        with torch.cuda.stream(side):
            _, segmentation = self.seg_engine(conv_impl).run(x)
        x.record_stream(side)
        inv_depth, _ = self.engine(conv_impl).run(x)
        main.wait_stream(side)
        return inv_depth, segmentation

    def network_outputs(self, batch, device):
        """the static (inverse depth, segmentation) buffers `network` fills for this batch size (soccdpt_b200.pipeline)."""
        return (self.engine().plan_for(batch, device)["depth"], self.seg_engine().plan_for(batch, device)["seg"])

    def forward(self, x: torch.Tensor):
        inv_depth, segmentation = self.network(x)
        return self.get_semantic_occupancy(inv_depth, segmentation)


class SOccDPT_V3(_EngineOwner, SOccDPT):
    def __init__(self, sigmoid=True, load_depth: str = DEPTH_l39icv3q, **kwargs):
        super(SOccDPT_V3, self).__init__(**kwargs)
        from .loader import load_model

        depth_model_weights = load_depth
        if depth_model_weights is None:
            depth_model_weights = default_depth_models[self.model_type]
        print("Loading depth net")
        self.depth_net = load_model(DPTDepthModel, dict(non_negative=True, return_features=True), cpu_device,
                                    depth_model_weights, self.model_type)
        self.depth_net.return_features = True
        self.pretrained = self.depth_net.pretrained     # alias: duplicated state_dict prefix, as in the reference

        activation = nn.Sigmoid() if sigmoid else ScaledTanh()
        self.seg_head = nn.Sequential(
            nn.Conv2d(self.features, self.features, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(self.features),
            nn.ReLU(True),
            nn.Dropout(0.1, False),
            nn.Conv2d(self.features, self.num_classes, kernel_size=1),
            Interpolate(scale_factor=2, mode="bilinear", align_corners=True),
            activation,
        )
        self._engine = None
        self.load_net(self.path)

    def engine(self, conv_impl=None):
        if self._engine is None or (conv_impl is not None and self._engine.conv_impl != conv_impl):
            self._engine = NetworkEngine(self, self._conv_impl(conv_impl))
        return self._engine

    def network(self, x):
        """image -> (inverse depth (B,h,w) f32, segmentation (B,C,h,w) f32), i.e. depth_net + seg_head
        (SOccDPT.py:682-683).  Returned tensors are static engine buffers."""
        if self.training:
            raise _cabi.SoccdptError("soccdpt_b200 implements the inference path only: call net.eval() first")
        return self.engine().run(x)

    def network_outputs(self, batch, device):
        """the static (inverse depth, segmentation) buffers `network` fills for this batch size (soccdpt_b200.pipeline)."""
        plan = self.engine().plan_for(batch, device)
        return plan["depth"], plan["seg"]

    def forward(self, x: torch.Tensor):
        inv_depth, segmentation = self.network(x)
        return self.get_semantic_occupancy(inv_depth, segmentation)


# SOccDPT_V2 is not constructible in the reference either (``seg_ead`` / ``seg_head`` typo, SOccDPT.py:580-621)
SOccDPT_versions = {1: SOccDPT_V1, 3: SOccDPT_V3}


class DepthNet:
    def __init__(self, net: Type[SOccDPT]) -> None:
        self.net = net

    def __call__(self, x: torch.Tensor):
        y_disp_pred, _, _, _ = self.net(x)
        return y_disp_pred

    def eval(self):
        self.net.eval()

    def train(self):
        self.net.train()


class SegNet:
    def __init__(self, net: Type[SOccDPT]) -> None:
        self.net = net

    def __call__(self, x: torch.Tensor):
        _, y_seg_pred, _, _ = self.net(x)
        return y_seg_pred

    def eval(self):
        self.net.eval()

    def train(self):
        self.net.train()
