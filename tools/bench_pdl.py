"""In-process A/B of programmatic dependent launch per kernel family (csrc/common.cuh; soccdpt_set_pdl mask bits:
1 conv / linear, 2 attention, 4 normalisation / element-wise, 8 post-processing) on the headline workload: the masks are
visited round-robin several times on the same box, CUDA events around `--steps` forwards each, median per mask.

    PYTHONPATH=. python tools/bench_pdl.py [--batch 64] [--masks 0,1,3,5,9,15]
"""
import argparse
import statistics

import torch

from soccdpt_b200 import SOccDPT_versions, _cabi, load_model
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dpt_swin2_tiny_256")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--masks", default="0,1,3,5,9,15")
    ap.add_argument("--network-only", action="store_true")
    a = ap.parse_args()
    masks = [int(m) for m in a.masks.split(",")]
    yml = write_calib_yaml("/tmp/bench_pdl_calib.yaml")
    net = load_model(arch=SOccDPT_versions[3],
                     model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                                       camera_intrinsics_yaml=yml, model_type=a.model),
                     device=torch.device("cpu"), model_path=None, model_type=a.model)
    net.load_state_dict(seeded_state_dict(net.state_dict(), 0, residual_gain=0.1), strict=True)
    net.to("cuda").eval()
    x = synthetic_frames(a.batch, net.depth_net.pretrained.model.img_size, 0).cuda()
    fwd = net.network if a.network_only else net
    lib = _cabi.load()
    times = {m: [] for m in masks}
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(5):
            fwd(x)
        for r in range(a.rounds):
            for m in masks:
                lib.soccdpt_set_pdl(m)
                fwd(x)
                torch.cuda.synchronize()
                s.record()
                for _ in range(a.steps):
                    fwd(x)
                e.record()
                torch.cuda.synchronize()
                times[m].append(s.elapsed_time(e) / a.steps)
    lib.soccdpt_set_pdl(-1)
    base = statistics.median(times[masks[0]])
    for m in masks:
        med = statistics.median(times[m])
        print(f"{a.model} B={a.batch} pdl mask {m:2d}: median {med:.3f} ms/step (min {min(times[m]):.3f} max {max(times[m]):.3f})  "
              f"{a.batch / med * 1e3:.0f} frames/s  {100 * (base / med - 1):+.2f}% vs mask {masks[0]}")


if __name__ == "__main__":
    main()
