"""Experiment: one B-frame forward against the same frames as k concurrent B/k-frame forwards on k CUDA streams (independent
launch plans fill each other's kernel tails and launch gaps, the way SOccDPT_V1's two networks do).  Network only.

    PYTHONPATH=. python tools/bench_split.py [--batch 64] [--ways 2]
"""
import argparse

import torch

from soccdpt_b200 import SOccDPT_versions, load_model
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml


def build(model, yml, sd=None):
    net = load_model(arch=SOccDPT_versions[3],
                     model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                                       camera_intrinsics_yaml=yml, model_type=model),
                     device=torch.device("cpu"), model_path=None, model_type=model)
    if sd is None:
        sd = seeded_state_dict(net.state_dict(), 0, residual_gain=0.1)
    net.load_state_dict(sd, strict=True)
    return net.to("cuda").eval(), sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dpt_swin2_tiny_256")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ways", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    yml = write_calib_yaml("/tmp/bench_split_calib.yaml")
    nets, sd = [], None
    for _ in range(a.ways):
        net, sd = build(a.model, yml, sd)
        nets.append(net)
    x = synthetic_frames(a.batch, nets[0].depth_net.pretrained.model.img_size, 0).cuda()
    parts = list(x.chunk(a.ways))
    streams = [torch.cuda.Stream() for _ in range(a.ways)]
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def whole():
        nets[0].network(x)

    def split():
        main_s = torch.cuda.current_stream()
        for st, net, p in zip(streams, nets, parts):
            st.wait_stream(main_s)
            with torch.cuda.stream(st):
                net.network(p)
        for st in streams:
            main_s.wait_stream(st)

    with torch.no_grad():
        for name, fn in (("one forward", whole), (f"{a.ways} concurrent forwards", split), ("one forward", whole),
                         (f"{a.ways} concurrent forwards", split)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            s.record()
            for _ in range(a.iters):
                fn()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / a.iters
            print(f"{a.model} B={a.batch} {name:24s}: {ms:.3f} ms/step -> {a.batch / ms * 1e3:.0f} frames/s (network only)")


if __name__ == "__main__":
    main()
