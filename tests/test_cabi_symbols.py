"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every symbol
include/soccdpt_b200.h declares; argument validation returns error codes (no GPU compute here)."""
import ctypes
import os
import re

import pytest

from soccdpt_b200 import _cabi
from soccdpt_b200.build import build


@pytest.fixture(scope="module")
def lib():
    build()
    return _cabi.load()


def test_header_symbols_are_exported(lib, repo_root):
    hdr = open(os.path.join(repo_root, "include", "soccdpt_b200.h")).read()
    declared = set(re.findall(r"\b(soccdpt_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"soccdpt_stream_t", "soccdpt_geometry_t", "soccdpt_conv_t"}
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.SYMBOLS), (declared ^ set(_cabi.SYMBOLS))
    raw = ctypes.CDLL(_cabi.lib_path())
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"


def test_struct_layouts_match_header():
    # soccdpt_geometry_t: 4 f32, 3 i32, 3 i32, 3+3+3+27 f32, 1 i32, 2 f32 ; soccdpt_conv_t: 7 ptr, 8 i32, 3 ptr, 4 i32, 1 ptr, 1 i32 (+ 4 bytes padding), 1 ptr, 2 i32
    assert ctypes.sizeof(_cabi.Geometry) == 4 * (4 + 3 + 3 + 3 + 3 + 3 + 27 + 1 + 2)
    # soccdpt_block_tail_t: 9 ptr, i64, 3 i32, f32
    assert ctypes.sizeof(_cabi.BlockTail) == 9 * 8 + 8 + 4 * 4
    assert ctypes.sizeof(_cabi.Conv) == 7 * 8 + 8 * 4 + 3 * 8 + 4 * 4 + 8 + 8 + 8 + 8
    assert _cabi.Conv.qk_scale.offset == 128 and _cabi.Conv.qk_heads.offset == 136
    assert _cabi.Conv.up_src.offset == 144 and _cabi.Conv.up_h.offset == 152 and _cabi.Conv.up_w.offset == 156


def test_version_and_error_paths(lib):
    assert lib.soccdpt_abi_version() == 1
    g = _cabi.Geometry()
    g.num_classes = 9  # unsupported
    rc = lib.soccdpt_voxelize_fwd(None, None, 1, ctypes.byref(g), None, None, 0, None, 0, None)
    assert rc == -1 and b"num_classes" in lib.soccdpt_last_error()
    rc = lib.soccdpt_conv_fwd(None, None)
    assert rc == -1
    assert lib.soccdpt_voxel_workspace_bytes(None, 1, 0) == 0
    g.num_classes = 3
    g.grid[0], g.grid[1], g.grid[2] = 256, 256, 32
    g.height, g.width = 1080, 1920
    tables = 1920 * 32 + 1080 * 48
    assert lib.soccdpt_voxel_workspace_bytes(ctypes.byref(g), 4, 0) == 256 * 256 * 32 // 8 * 4 + tables
    assert lib.soccdpt_voxel_workspace_bytes(ctypes.byref(g), 4, 1) == 4 * 256 * 256 * 32 // 8 * 4 + tables


def test_product_never_imports_oracle(repo_root):
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(repo_root, "soccdpt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "soccdpt_oracle" not in src and "voxel_oracle" not in src and "ref_env" not in src, f
                assert "timm_shim" not in src, f


def test_cpu_tensors_fail_loudly(tmp_path):
    import torch
    from soccdpt_b200 import SOccDPT
    from soccdpt_b200.synthetic import write_calib_yaml
    net = SOccDPT(camera_intrinsics_yaml=write_calib_yaml(str(tmp_path / "c.yaml")), compute_occ=True)
    with pytest.raises(_cabi.SoccdptError):
        net.get_semantic_occupancy(torch.rand(1, 8, 8), torch.rand(1, 3, 8, 8))
