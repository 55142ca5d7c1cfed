"""Input-pipeline kernel (csrc/preprocess.cu) on 64 synthetic 1080p uint8 frames: CUDA events, effective bandwidth over the
bytes it has to touch (4 source rows per output row -- at 1080 -> 256 that is 1024 of the 1080 rows -- plus the fp32 output),
beside cv2.resize + normalise on the host cores (one frame per core is the reference's own arrangement: a DataLoader worker each).

    PYTHONPATH=. python tools/bench_preprocess.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccdpt_b200.preprocess import load_gpu_transforms  # noqa: E402
from soccdpt_b200 import load_transforms  # noqa: E402


def main():
    peak = 6539.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbps"]
    except Exception:
        pass
    B, H, W = 64, 1080, 1920
    rng = np.random.default_rng(0)
    frames = torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8))
    for mt in ("dpt_swin2_tiny_256", "dpt_hybrid_384"):
        t, _, _ = load_gpu_transforms(mt)
        x = frames.cuda()
        dh, dw = t.output_size(H, W)
        out = torch.empty((B, 3, dh, dw), dtype=torch.float32, device="cuda")
        for _ in range(3):
            t(x, out)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            t(x, out)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        nbytes = B * (min(4 * dh, H) * W * 3 + 4 * 3 * dh * dw)
        tc, _, _ = load_transforms(mt)
        f0 = frames[0].numpy()
        t0 = time.time()
        n = 0
        while time.time() - t0 < 2.0:
            tc({"image": f0})
            n += 1
        cpu_ms = (time.time() - t0) / n * 1e3
        print(f"{mt}: {B} x {H}x{W} -> {dh}x{dw}: {ms:.3f} ms ({B / ms * 1e3:.0f} frames/s, {nbytes / ms / 1e6:.0f} GB/s = "
              f"{nbytes / ms / 1e6 / peak:.2f} of HBM peak); cv2 on one host core: {cpu_ms:.2f} ms/frame ({1e3 / cpu_ms:.0f} frames/s)")


if __name__ == "__main__":
    main()
