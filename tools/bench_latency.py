"""Small-batch latency of the network (image -> inverse depth + segmentation at network resolution): eager launch list vs
CUDA-graph replay (`net.engine().enable_graphs()`), wall clock around a synchronised forward, median of many calls.

    PYTHONPATH=. python tools/bench_latency.py [--model dpt_swin2_tiny_256]
"""
import argparse
import statistics
import time

import torch

from soccdpt_b200 import SOccDPT_versions, load_model
from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames, write_calib_yaml


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dpt_swin2_tiny_256")
    ap.add_argument("--iters", type=int, default=200)
    a = ap.parse_args()
    yml = write_calib_yaml("/tmp/bench_latency_calib.yaml")
    net = load_model(arch=SOccDPT_versions[3],
                     model_kwargs=dict(load_depth=False, num_classes=3, sigmoid=True, compute_occ=True,
                                       camera_intrinsics_yaml=yml, model_type=a.model),
                     device=torch.device("cpu"), model_path=None, model_type=a.model)
    net.load_state_dict(seeded_state_dict(net.state_dict(), 0), strict=True)
    net.to("cuda").eval()
    size = net.depth_net.pretrained.model.img_size
    for B in (1, 2, 4, 8):
        x = synthetic_frames(B, size, 0).cuda()
        res = {}
        for mode in ("eager", "graph"):
            net.engine().enable_graphs(mode == "graph")
            with torch.no_grad():
                for _ in range(5):
                    net.network(x)
                torch.cuda.synchronize()
                ts = []
                for _ in range(a.iters):
                    t0 = time.perf_counter()
                    net.network(x)
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
            res[mode] = statistics.median(ts) * 1e3
        print(f"{a.model} B={B}: eager {res['eager']:.3f} ms  graph {res['graph']:.3f} ms  "
              f"({B / res['graph'] * 1e3:.0f} frames/s at this batch)")
    net.engine().enable_graphs(False)


if __name__ == "__main__":
    main()
