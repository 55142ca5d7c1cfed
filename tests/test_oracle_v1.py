"""SOccDPT_V1 (reference SOccDPT.py:470-523): the oracle restatement pinned against the
fixture the UNMODIFIED reference produced (oracle/make_golden_v1.py) and, in the build container, against the
reference itself; the drop-in's state_dict keys against the reference's recorded key list.  CPU only."""
import numpy as np
import pytest
import torch

import golden_util as GU
import ref_env
import soccdpt_oracle as O
from soccdpt_b200 import SOccDPT_versions, load_model
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml

MT = "dpt_swin2_tiny_256"


def _kwargs(yml):
    return dict(load_depth=False, load_seg=False, num_classes=3, compute_occ=True, camera_intrinsics_yaml=yml, model_type=MT)


def test_v1_oracle_matches_reference_fixture():
    z = np.load(GU.GOLD + "/net_v1_tiny_b2.npz")
    sd = GU.v1_tiny_state_dict(0)
    assert len(sd) == int(z["n_state_keys"])
    orc = O.OracleV1(sd)
    d, g, _, _ = orc.network(synthetic_frames(2, 256, 0))
    # same torch build -> bit-equal; other builds/CPUs -> fp32 reassociation noise only
    assert np.allclose(d.numpy(), z["depth"], rtol=1e-4, atol=1e-5)
    assert np.allclose(g.numpy(), z["seg"], rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(not ref_env.reference_available(), reason="reference tree not present")
def test_v1_oracle_equals_reference_live(tmp_path):
    ref_loader, ref_model = ref_env.import_reference()
    yml = write_calib_yaml(str(tmp_path / "calib.yaml"))
    net = ref_loader.load_model(arch=ref_model.SOccDPT_versions[1], model_kwargs=_kwargs(yml), device=torch.device("cpu"),
                                model_path=None, model_type=MT).eval()
    sd = seeded_state_dict(net.state_dict(), 4)
    net.load_state_dict(sd, strict=True)
    orc = O.OracleV1(sd)
    for B in (1, 2):
        x = synthetic_frames(B, 256, 6)
        with torch.no_grad():
            ref = net(x)
        out = orc(x)
        for r, o in zip(ref, out):
            r, o = torch.as_tensor(r), torch.as_tensor(o)
            assert r.shape == o.shape and torch.equal(torch.nan_to_num(r, nan=-7.0), torch.nan_to_num(o, nan=-7.0))


def test_v1_drop_in_keys_equal_reference(tmp_path):
    yml = write_calib_yaml(str(tmp_path / "calib.yaml"))
    net = load_model(arch=SOccDPT_versions[1], model_kwargs=_kwargs(yml), device=torch.device("cpu"), model_path=None,
                     model_type=MT)
    with open(GU.GOLD + "/state_keys_v1_tiny.txt") as f:
        ref_lines = [line.rstrip("\n") for line in f]
    ours = [f"{k} {tuple(v.shape)}" for k, v in net.state_dict().items()]
    assert ours == ref_lines                      # same keys, same shapes, same order
    net.load_state_dict(GU.v1_tiny_state_dict(0), strict=True)
    assert net.pretrained is net.depth_net.pretrained and hasattr(net.seg_net, "auxlayer")
    with pytest.raises(Exception):                # no CPU path
        net.eval()(torch.zeros(1, 3, 256, 256))
