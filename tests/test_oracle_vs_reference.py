"""Live pin: oracle vs the UNMODIFIED reference imported from /root/reference (build container
only -- skipped on the GPU box where the reference tree does not exist)."""
import numpy as np
import pytest
import torch

import golden_util as GU
import ref_env
import soccdpt_oracle as O
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml

pytestmark = pytest.mark.skipif(not ref_env.reference_available(), reason="reference tree not present")


def _eq(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return a.shape == b.shape and torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))


def test_oracle_equals_reference_end_to_end(tmp_path):
    ref_loader, ref_model = ref_env.import_reference()
    yml = write_calib_yaml(str(tmp_path / "calib.yaml"))
    mt = "dpt_swin2_tiny_256"
    net = ref_loader.load_model(
        arch=ref_model.SOccDPT_versions[3],
        model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=False, compute_occ=True,
                          camera_intrinsics_yaml=yml, model_type=mt),
        device=torch.device("cpu"), model_path=None, model_type=mt).eval()
    sd = seeded_state_dict(net.state_dict(), 3)
    net.load_state_dict(sd, strict=True)
    orc = O.OracleV3(sd, sigmoid=False)
    for B in (1, 2):
        x = synthetic_frames(B, 256, 5)
        with torch.no_grad():
            ref = net(x)
        out = orc(x)
        for r, o in zip(ref, out):
            assert _eq(r, o)      # includes the B=1 squeeze quirk on the segmentation output


@pytest.mark.parametrize("name", ["small_b2", "small_b1_tanh"])
def test_voxel_oracle_equals_reference_live(name, tmp_path):
    _, ref_model = ref_env.import_reference()
    z, calib, geom, inv, seg = GU.load_voxel_case(name)
    yml = write_calib_yaml(str(tmp_path / "calib.yaml"), calib)
    net = ref_model.SOccDPT(camera_intrinsics_yaml=yml, compute_occ=True, grid_size=geom.grid_size,
                            scale=tuple(float(s) for s in z["scale"]))
    with torch.no_grad():
        r = net.get_semantic_occupancy(inv.clone(), seg.clone())
    o = O.get_semantic_occupancy(inv.clone(), seg.clone(), geom)
    for a, b in zip(r, o):
        assert _eq(a, b)
    assert np.array_equal(net.occupancy_shape, geom.occupancy_shape)


def test_hybrid_oracle_equals_repaired_reference(tmp_path):
    """the reference's dpt_hybrid_384 constructor raises NameError as shipped; with the one-token
    repair applied in memory (oracle/ref_env.py) the oracle restatement is bit-equal to it on all four outputs."""
    ref_loader, ref_model = ref_env.import_reference()
    yml = write_calib_yaml(str(tmp_path / "calib.yaml"))
    mt = "dpt_hybrid_384"
    net = ref_loader.load_model(
        arch=ref_model.SOccDPT_versions[3],
        model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                          camera_intrinsics_yaml=yml, model_type=mt),
        device=torch.device("cpu"), model_path=None, model_type=mt).eval()
    sd = seeded_state_dict(net.state_dict(), 2)
    net.load_state_dict(sd, strict=True)
    orc = O.OracleV3(sd, mt)
    x = synthetic_frames(1, 384, 3)
    with torch.no_grad():
        ref = net(x)
    out = orc(x)
    for r, o in zip(ref, out):
        assert _eq(r, o)
