"""ctypes binding of include/soccdpt_b200.h (the drop-in boundary).

There is NO fallback: if the shared library is missing or a call fails, this module raises.
PyTorch is used by callers only to own device memory and streams; nothing here takes torch types
across the boundary -- just raw pointers, sizes and a cudaStream_t.
"""
import ctypes
import os

# SOCCDPT_LIB: an alternative build of the SAME library (A/B experiments with compile-time switches, tools/ only)
_LIB_PATH = os.environ.get("SOCCDPT_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libsoccdpt_b200.so")
_lib = None

c_float_p = ctypes.POINTER(ctypes.c_float)
c_void_p = ctypes.c_void_p

OCC_REFERENCE_UNION = 0
OCC_PER_FRAME = 1
OCC_PACKED = 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


class SoccdptError(RuntimeError):
    pass


class Geometry(ctypes.Structure):
    """soccdpt_geometry_t"""
    _fields_ = [
        ("fx", ctypes.c_float), ("fy", ctypes.c_float), ("cx", ctypes.c_float), ("cy", ctypes.c_float),
        ("height", ctypes.c_int), ("width", ctypes.c_int), ("num_classes", ctypes.c_int),
        ("grid", ctypes.c_int * 3), ("occ_shape", ctypes.c_float * 3),
        ("pc_scale", ctypes.c_float * 3), ("pc_shift", ctypes.c_float * 3), ("rot", ctypes.c_float * 27),
        ("scalar_div_by_reciprocal", ctypes.c_int), ("rcp_fx", ctypes.c_float), ("rcp_fy", ctypes.c_float),
    ]


class Conv(ctypes.Structure):
    """soccdpt_conv_t"""
    _fields_ = [
        ("x", c_void_p), ("wgt", c_void_p), ("bias", c_void_p), ("res1", c_void_p), ("res2", c_void_p),
        ("y", c_void_p), ("y_relu", c_void_p),
        ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("Cin", ctypes.c_int),
        ("Cout", ctypes.c_int), ("KH", ctypes.c_int), ("KW", ctypes.c_int), ("act", ctypes.c_int),
        ("proj_w", c_void_p), ("proj_b", c_void_p), ("proj_out", c_void_p),
        ("proj_n", ctypes.c_int), ("proj_relu", ctypes.c_int), ("stride", ctypes.c_int), ("pad_trim", ctypes.c_int),
        ("qk_scale", c_void_p), ("qk_heads", ctypes.c_int),
        ("up_src", c_void_p), ("up_h", ctypes.c_int), ("up_w", ctypes.c_int),
    ]


class BlockTail(ctypes.Structure):
    """soccdpt_block_tail_t"""
    _fields_ = [
        ("x", c_void_p), ("w1", c_void_p), ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p),
        ("gamma", c_void_p), ("beta", c_void_p), ("master", c_void_p), ("y", c_void_p),
        ("M", ctypes.c_longlong), ("K1", ctypes.c_int), ("HID", ctypes.c_int), ("C", ctypes.c_int), ("eps", ctypes.c_float),
    ]


# name -> (restype, argtypes); every symbol include/soccdpt_b200.h declares
_I, _LL, _F, _SZ = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_size_t
SYMBOLS = {
    "soccdpt_abi_version": (_I, []),
    "soccdpt_last_error": (ctypes.c_char_p, []),
    "soccdpt_launch_count": (_LL, []),
    "soccdpt_set_pdl": (_I, [_I]),
    "soccdpt_device_info": (_I, [ctypes.POINTER(_I)] * 3),
    "soccdpt_preprocess_workspace_bytes": (_SZ, [_I, _I]),
    "soccdpt_preprocess_fwd": (_I, [c_void_p, _I, _I, _I, _I, c_void_p, _I, _I, c_void_p, _SZ, c_void_p]),
    "soccdpt_occupancy_mask_bytes": (_SZ, [ctypes.POINTER(_I)]),
    "soccdpt_occupancy_points_workspace_bytes": (_SZ, [ctypes.POINTER(_I), _I]),
    "soccdpt_grid_pack_fwd": (_I, [c_void_p, ctypes.POINTER(_I), _I, c_void_p, c_void_p]),
    "soccdpt_occupancy_points_fwd": (_I, [c_void_p, ctypes.POINTER(_I), ctypes.POINTER(_F), _I, c_void_p, _LL, c_void_p,
                                         c_void_p, _SZ, c_void_p]),
    "soccdpt_voxel_count_fwd": (_I, [c_void_p, _I, c_void_p, _I, _LL, ctypes.POINTER(_I), ctypes.POINTER(_F), _I, c_void_p, c_void_p,
                                     c_void_p]),
    "soccdpt_voxel_count_finish_fwd": (_I, [c_void_p, ctypes.POINTER(_I), _I, _F, c_void_p, c_void_p, c_void_p, c_void_p]),
    "soccdpt_voxel_workspace_bytes": (_SZ, [ctypes.POINTER(Geometry), _I, _I]),
    "soccdpt_voxelize_fwd": (_I, [c_void_p, c_void_p, _I, ctypes.POINTER(Geometry), c_void_p, c_void_p, _I,
                                  c_void_p, _SZ, c_void_p]),
    "soccdpt_postprocess_fwd": (_I, [c_void_p, c_void_p, _I, _I, _I, ctypes.POINTER(Geometry), c_void_p, c_void_p,
                                     c_void_p, c_void_p, _I, c_void_p, _SZ, c_void_p]),
    "soccdpt_selftest_exact_math": (_I, [ctypes.POINTER(Geometry), ctypes.POINTER(ctypes.c_ulonglong * 6), c_void_p]),
    "soccdpt_conv_fwd": (_I, [ctypes.POINTER(Conv), c_void_p]),
    "soccdpt_conv_ref_fwd": (_I, [ctypes.POINTER(Conv), c_void_p]),
    "soccdpt_conv_f32_fwd": (_I, [c_void_p] * 3 + [_I] * 7 + [c_void_p]),
    "soccdpt_groupnorm_f32_fwd": (_I, [c_void_p] * 5 + [_I] * 3 + [ctypes.c_float, _I, c_void_p]),
    "soccdpt_maxpool3s2_f32_fwd": (_I, [c_void_p] * 2 + [_I] * 4 + [c_void_p]),
    "soccdpt_nchw_to_nhwc_f32": (_I, [c_void_p] * 2 + [_I] * 3 + [c_void_p]),
    "soccdpt_patch_embed_fwd": (_I, [c_void_p] * 7 + [_I] * 4 + [c_void_p]),
    "soccdpt_window_attention_fwd": (_I, [c_void_p] * 4 + [_I] * 7 + [c_void_p]),
    "soccdpt_window_attention_normed_fwd": (_I, [c_void_p] * 4 + [_I] * 6 + [c_void_p]),
    "soccdpt_layernorm_fwd": (_I, [c_void_p] * 5 + [_LL, _I, _F, c_void_p]),
    "soccdpt_layernorm_master_fwd": (_I, [c_void_p, c_void_p, _I, c_void_p, c_void_p, c_void_p, _LL, _I, _F, c_void_p]),
    "soccdpt_swin_block_tail_fwd": (_I, [ctypes.POINTER(BlockTail), c_void_p]),
    "soccdpt_patch_merge_gather_fwd": (_I, [c_void_p, c_void_p, _I, _I, _I, _I, c_void_p]),
    "soccdpt_upsample_bilinear_fwd": (_I, [c_void_p, c_void_p] + [_I] * 6 + [c_void_p]),
    "soccdpt_seg_finish_fwd": (_I, [c_void_p, c_void_p] + [_I] * 5 + [c_void_p]),
    "soccdpt_depth_tail_fwd": (_I, [c_void_p] * 5 + [_I] * 3 + [c_void_p]),
    "soccdpt_f32_to_bf16": (_I, [c_void_p, c_void_p, _LL, c_void_p]),
    "soccdpt_bf16_to_f32": (_I, [c_void_p, c_void_p, _LL, c_void_p]),
    "soccdpt_stem_conv7_fwd": (_I, [c_void_p] * 3 + [_I] * 3 + [c_void_p]),
    "soccdpt_groupnorm_fwd": (_I, [c_void_p] * 5 + [_I] * 3 + [_F, _I, c_void_p, c_void_p]),
    "soccdpt_maxpool3s2_fwd": (_I, [c_void_p] * 2 + [_I] * 4 + [c_void_p]),
    "soccdpt_vit_tokens_fwd": (_I, [c_void_p] * 5 + [_I] * 3 + [c_void_p]),
    "soccdpt_prenorm_fwd": (_I, [c_void_p] * 6 + [_LL, _I, _F, c_void_p]),
    "soccdpt_readout_concat_fwd": (_I, [c_void_p] * 2 + [_I] * 3 + [c_void_p]),
    "soccdpt_global_attention_fwd": (_I, [c_void_p] * 2 + [_I] * 4 + [c_void_p]),
}


def lib_path():
    return _LIB_PATH


def load():
    """Loads the shared library (once) and types every exported symbol. Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise SoccdptError(
                f"{_LIB_PATH} not found: build it with `python -m soccdpt_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.soccdpt_abi_version() != 1:
            raise SoccdptError("ABI version mismatch between _cabi.py and the shared library")
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().soccdpt_last_error().decode("utf-8", "replace")
        raise SoccdptError(f"{what or 'soccdpt call'} failed (rc={rc}): {msg}")


def ptr(t):
    """Raw device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream(device=None):
    """cudaStream_t of torch's current stream on ``device`` (default: the current device).  Launch sites pass the device of
    the tensors they hand over and run inside ``torch.cuda.device(device)``: the library launches on the CURRENT device."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def normalize_device(device):
    """torch.device('cuda') -> torch.device('cuda', current index): ONE spelling per device for caches keyed by device."""
    import torch
    device = torch.device(device)
    if device.type != "cuda":
        raise SoccdptError("soccdpt_b200 runs on CUDA devices only (no CPU fallback)")
    return torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())


def launch_count():
    return int(load().soccdpt_launch_count())
