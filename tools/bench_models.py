"""Device-resident throughput of the image -> (inverse depth, segmentation) network for every model type the package
builds (rows A10-A13), with a per-kernel-family breakdown.  Not the headline bench (bench.py); a profiling aid.

    PYTHONPATH=. python tools/bench_models.py --model dpt_hybrid_384 --batch 16
"""
import argparse
import collections

import torch

from soccdpt_b200 import SOccDPT_versions, load_model
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dpt_hybrid_384")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--version", type=int, default=3, choices=(1, 3), help="SOccDPT_V1 (two DPTs, two streams) or SOccDPT_V3")
    a = ap.parse_args()
    yml = write_calib_yaml("/tmp/bench_models_calib.yaml")
    kw = dict(load_depth=False, num_classes=3, compute_occ=True, camera_intrinsics_yaml=yml, model_type=a.model)
    kw.update(dict(sigmoid=True) if a.version == 3 else dict(load_seg=False))
    net = load_model(arch=SOccDPT_versions[a.version], model_kwargs=kw,
                     device=torch.device("cpu"), model_path=None, model_type=a.model)
    net.load_state_dict(seeded_state_dict(net.state_dict(), 0, residual_gain=0.1), strict=True)
    net.to("cuda").eval()
    size = net.depth_net.pretrained.model.img_size
    x = synthetic_frames(a.batch, size, 0).cuda()
    with torch.no_grad():
        for _ in range(3):
            net.network(x)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.iters):
            net.network(x)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / a.iters
        print(f"V{a.version} {a.model} B={a.batch}: {ms:.3f} ms/step network only -> {a.batch / ms * 1e3:.1f} frames/s")
        for _ in range(2):
            net(x)
        torch.cuda.synchronize()
        s.record()
        for _ in range(a.iters):
            net(x)
        e.record()
        torch.cuda.synchronize()
        ms2 = s.elapsed_time(e) / a.iters
        print(f"V{a.version} {a.model} B={a.batch}: {ms2:.3f} ms/step image -> occupancy -> {a.batch / ms2 * 1e3:.1f} frames/s")
        if a.version == 1:
            for name, eng in (("depth DPT", net.engine()), ("segmentation DPT", net.seg_engine())):
                s.record()
                for _ in range(a.iters):
                    eng.run(x)
                e.record()
                torch.cuda.synchronize()
                print(f"  {name} alone on one stream: {s.elapsed_time(e) / a.iters:.3f} ms/step")
        # per-family breakdown: every op bracketed by events (serialising, so the sum exceeds the step time a little)
        from soccdpt_b200 import _cabi
        plan = net.engine().plan_for(a.batch, x.device)
        stream = _cabi.current_stream()
        tot = collections.defaultdict(float)
        cnt = collections.Counter()
        for rep in range(3):
            evs = []
            for op in plan["ops"]:
                s1, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s1.record()
                op(stream)
                e1.record()
                evs.append((op, s1, e1))
            torch.cuda.synchronize()
            if rep == 2:
                for op, s1, e1 in evs:
                    name = op.name
                    if name == "conv":
                        c = op.args[0]._obj
                        name = f"conv{c.KH}x{c.KH}" + (f"/s{c.stride}" if c.stride > 1 else "") + (" gemm" if c.H == 1 and c.N == 1 else "")
                    tot[name] += s1.elapsed_time(e1)
                    cnt[name] += 1
        total = sum(tot.values())
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            print(f"  {k:22s} {cnt[k]:4d} launches {v:8.3f} ms {100 * v / total:5.1f}%")


if __name__ == "__main__":
    main()
