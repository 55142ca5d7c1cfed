"""Debug aid: walk the hybrid encoder step by step on the GPU next to the oracle's modules (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import torch
import cuda_ops as K
import soccdpt_oracle as O
from soccdpt_b200 import SOccDPT_versions, load_model
from soccdpt_b200.engine import NetworkEngine
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml

MT = "dpt_hybrid_384"
yml = write_calib_yaml("/tmp/c.yaml")
net = load_model(arch=SOccDPT_versions[3], model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                 camera_intrinsics_yaml=yml, model_type=MT), device=torch.device("cpu"), model_path=None, model_type=MT)
sd = seeded_state_dict(net.state_dict(), 0)
net.load_state_dict(sd, strict=True)
net.to("cuda").eval()
orc = O.OracleV3(sd, MT)
bb = orc.encoder.patch_embed.backbone
x = synthetic_frames(1, 384, 0)


def cmp(name, mine, ref):
    got = mine.float().permute(0, 3, 1, 2).cpu()
    e = (got - ref).abs()
    print(f"{name:34s} max {e.max().item():.3e} mean {e.mean().item():.3e} | ref max {ref.abs().max().item():.3e} mean {ref.abs().mean().item():.3e}")


def pk(w):
    return w.reshape(w.shape[0], -1, w.shape[-1])


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()


with torch.no_grad():
    eng = net.engine()
    W = eng._pack(torch.device("cuda"))["hy"]
    r = bb.stem.conv(x)
    m = K.stem_conv7(x.cuda(), W["stem_w"])
    cmp("stem conv", m, r)
    r2 = bb.stem.norm(r)
    m2 = K.groupnorm(m, *W["stem_n"], relu=True)
    cmp("stem gn", m2, r2)
    cmp("stem gn (ref input)", K.groupnorm(nhwc(r), *W["stem_n"], relu=True), r2)
    r3 = bb.stem.pool(r2)
    m3 = K.maxpool3s2(m2)
    cmp("pool", m3, r3)
    blk = bb.stages[0].blocks[0]
    wb = W["stages"][0][0]
    rin = r3
    ra = blk.conv1(rin)
    ma, _, _ = K.conv(nhwc(rin), pk(wb["w1"]))
    cmp("b0 conv1 (ref input)", ma, ra)
    ra2 = blk.norm1(ra)
    cmp("b0 gn1 (ref input)", K.groupnorm(nhwc(ra), *wb["n1"], relu=True), ra2)
    rb = blk.conv2(ra2)
    mb, _, _ = K.conv(nhwc(ra2), pk(wb["w2"]))
    cmp("b0 conv2 (ref input)", mb, rb)
    rb2 = blk.norm2(rb)
    rc = blk.conv3(rb2)
    mc, _, _ = K.conv(nhwc(rb2), pk(wb["w3"]))
    cmp("b0 conv3 (ref input)", mc, rc)
    rsc = blk.downsample(rin)
    msc, _, _ = K.conv(nhwc(rin), pk(wb["down"][0]))
    cmp("b0 down conv (ref input)", msc, blk.downsample.conv(rin))
    rout = blk(rin)
    mout = K.groupnorm(nhwc(rc), *wb["n3"], shortcut=nhwc(rsc), relu=True)
    cmp("b0 gn3+sc (ref input)", mout, rout)
    # whole stage 0 through the engine plan taps
    net.network(x.cuda())
    torch.cuda.synchronize()
    plan = eng.plan_for(1, torch.device("cuda", 0))
    s0 = bb.stages[0](r3)
    t, H, Wd, C = plan["taps"][0]
    cmp("stage0 (engine)", t.reshape(1, H, Wd, C), s0)
    s1 = bb.stages[1](s0)
    t, H, Wd, C = plan["taps"][1]
    cmp("stage1 (engine)", t.reshape(1, H, Wd, C), s1)
    # stage-2 strided block with ref input
    blk = bb.stages[1].blocks[0]
    wb = W["stages"][1][0]
    ra2 = blk.norm1(blk.conv1(s0))
    rb = blk.conv2(ra2)
    mb, _, _ = K.conv(nhwc(ra2), pk(wb["w2"]), stride=2, pad_trim=1)
    cmp("s1b0 conv2 s2 (ref input)", mb, rb)
    msc, _, _ = K.conv(nhwc(s0), pk(wb["down"][0]), stride=2)
    cmp("s1b0 down s2 (ref input)", msc, blk.downsample.conv(s0))
    ref = orc.encoder_taps(x)
    for i in (2, 3):
        t, H, Wd, C = plan["taps"][i]
        cmp(f"tap{i+1} (engine)", t.reshape(1, H, Wd, C), ref[i])
