"""Host-side mirror of the reference's model API: names, constructor behaviour, state_dict keys."""
import ast
import os

import pytest
import torch

import golden_util as GU
from soccdpt_b200 import DepthNet, SegNet, SOccDPT, SOccDPT_V3, SOccDPT_versions, load_model, load_transforms, model_types
from soccdpt_b200.synthetic import write_calib_yaml


@pytest.fixture(scope="module")
def net(tmp_path_factory):
    yml = write_calib_yaml(str(tmp_path_factory.mktemp("calib") / "c.yaml"))
    return load_model(arch=SOccDPT_versions[3],
                      model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                                        camera_intrinsics_yaml=yml, model_type="dpt_swin2_tiny_256"),
                      device=torch.device("cpu"), model_path=None, model_type="dpt_swin2_tiny_256")


def test_state_dict_keys_equal_reference(net):
    ref = {}
    with open(os.path.join(GU.GOLD, "state_keys_tiny.txt")) as f:
        for line in f:
            k, shp = line.rstrip("\n").split(" ", 1)
            ref[k] = ast.literal_eval(shp)
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert list(mine.keys()) == list(ref.keys())
    assert mine == ref


def test_seeded_weights_load_strict(net):
    sd = GU.tiny_state_dict(0)
    assert net.load_state_dict(sd, strict=True).missing_keys == []


def test_reference_attributes(net):
    assert isinstance(net, SOccDPT) and isinstance(net, SOccDPT_V3)
    assert net.pretrained is net.depth_net.pretrained
    assert isinstance(net.occupancy_conv, torch.nn.Identity)
    assert (net.width, net.height) == (1920, 1080)
    assert net.occupancy_shape.dtype.name == "float32" and abs(float(net.occupancy_shape[2]) - 48.048048) < 1e-4
    assert net.compute_occ is True and net.grid_size == (256, 256, 32)
    assert "dpt_swin2_tiny_256" in model_types
    DepthNet(net).eval(), SegNet(net).eval()


def test_constructor_errors_match_reference(tmp_path):
    with pytest.raises(FileNotFoundError):
        SOccDPT(camera_intrinsics_yaml=str(tmp_path / "missing.yaml"))
    bad = tmp_path / "bad.yaml"
    bad.write_text("Camera.fx: 1.0\n")
    with pytest.raises(KeyError):
        SOccDPT(camera_intrinsics_yaml=str(bad))
    yml = write_calib_yaml(str(tmp_path / "c.yaml"))
    with pytest.raises(AssertionError):
        SOccDPT(camera_intrinsics_yaml=yml, point_compute_method="cupy")
    with pytest.raises(AssertionError):
        load_model(SOccDPT_V3, dict(load_depth=False, camera_intrinsics_yaml=yml), torch.device("cpu"), None, "dpt_nope")


def test_load_transforms_sizes():
    for mt, wh in (("dpt_swin2_tiny_256", (256, 256)), ("dpt_swin2_base_384", (256, 256)), ("dpt_hybrid_384", (384, 384))):
        t, w, h = load_transforms(mt)
        assert (w, h) == wh
    t, w, h = load_transforms("dpt_swin2_tiny_256", height=320)
    assert (w, h) == (320, 320)


def test_base_384_state_dict_keys_equal_reference_live(tmp_path):
    """dpt_swin2_base_384 builds with the reference's exact key set (live, build container only)."""
    import ref_env
    if not ref_env.reference_available():
        pytest.skip("reference tree not present")
    ref_loader, ref_model = ref_env.import_reference()
    yml = write_calib_yaml(str(tmp_path / "c.yaml"))
    mt = "dpt_swin2_base_384"
    kw = dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True, camera_intrinsics_yaml=yml, model_type=mt)
    ref = ref_loader.load_model(arch=ref_model.SOccDPT_versions[3], model_kwargs=dict(kw), device=torch.device("cpu"),
                                model_path=None, model_type=mt)
    mine = load_model(arch=SOccDPT_versions[3], model_kwargs=dict(kw), device=torch.device("cpu"), model_path=None,
                      model_type=mt)
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert list(a.keys()) == list(b.keys()) and a == b
    for k, v in ref.state_dict().items():
        if k.endswith("attn_mask"):
            assert torch.equal(v, mine.state_dict()[k])


def test_hybrid_state_dict_keys_equal_reference(tmp_path):
    """dpt_hybrid_384 builds with the key set / order / shapes of the (repaired) reference constructor,
    recorded by oracle/make_golden.py in tests/golden/state_keys_hybrid.txt."""
    ref = {}
    with open(os.path.join(GU.GOLD, "state_keys_hybrid.txt")) as f:
        for line in f:
            k, shp = line.rstrip("\n").split(" ", 1)
            ref[k] = ast.literal_eval(shp)
    yml = write_calib_yaml(str(tmp_path / "c.yaml"))
    mt = "dpt_hybrid_384"
    mine = load_model(arch=SOccDPT_versions[3],
                      model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                                        camera_intrinsics_yaml=yml, model_type=mt),
                      device=torch.device("cpu"), model_path=None, model_type=mt)
    got = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert list(got.keys()) == list(ref.keys())
    assert got == ref
    assert mine.pretrained is mine.depth_net.pretrained


def test_checkpoint_round_trip_like_reference(net, tmp_path):
    """both on-disk layouts the reference writes / reads load through the same path
    (raw state_dict, train_SOccDPT.py:437-449; {"optimizer","model"} wrapper, base_model.py:15-17)."""
    sd = GU.tiny_state_dict(1)
    raw, wrapped = tmp_path / "checkpoint_epoch_1.pth", tmp_path / "wrapped.pth"
    torch.save(sd, raw)
    torch.save({"optimizer": {"state": {}}, "model": sd}, wrapped)
    yml = net.camera_intrinsics_yaml
    for path in (raw, wrapped):
        m = load_model(arch=SOccDPT_versions[3],
                       model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                                         camera_intrinsics_yaml=yml, model_type="dpt_swin2_tiny_256"),
                       device=torch.device("cpu"), model_path=str(path), model_type="dpt_swin2_tiny_256")
        got = m.state_dict()
        for k in ("depth_net.pretrained.model.layers.2.blocks.3.attn.qkv.weight", "seg_head.1.running_var",
                  "depth_net.scratch.refinenet2.resConfUnit1.conv2.bias", "pretrained.model.patch_embed.proj.weight"):
            assert torch.equal(got[k], sd[k]), k
    # a checkpoint with foreign / missing keys loads with strict=False semantics (prints, never raises)
    partial = {k: v for k, v in sd.items() if "seg_head" not in k}
    partial["unknown.key"] = torch.zeros(1)
    torch.save(partial, tmp_path / "partial.pth")
    net.load_net(str(tmp_path / "partial.pth"))
