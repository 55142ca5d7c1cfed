#!/usr/bin/env python
"""Benchmark of the SOccDPT inference hot path (BASELINE.json metric):

    frames/s, SOccDPT-V3 dpt_swin2_tiny_256, image -> (inverse depth, segmentation, points, occupancy grid)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--impl ours|reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); frames are batch-sharded, every rank
runs the same per-GPU batch (weak scaling), there is no collective on the data path.  A "step" is one
``net(x)`` call on one batch of synthetic frames (random seeded weights, randn images, the synthetic
1920x1080 pinhole camera of SURVEY.md 8d).  Rank 0 prints ONE JSON line.

  value     frames/s with the input batch already resident in HBM (CUDA events, max over ranks)
  e2e       frames/s through the public host-frame API (soccdpt_b200.pipeline.FrameStream) with the batch in
            PINNED HOST memory: every step copies the images host->device and reads the step's result back
            device->host (network-resolution inverse depth + class maps, and the occupancy grid of the call --
            in the reference's semantics all B grid copies are identical, one copy is read); copies overlap
            with compute on separate streams, the timed region spans first H2D to last D2H
  roofline  tensor-pipe roofline of the dominant kernel (conv_tcgen05_kernel): algorithmic FLOPs of all its
            launches in a step / their CUDA-event time, vs MEASURED_PEAKS.json bf16 sustained
  roofline_voxeliser   HBM roofline of the post-processing kernels (compulsory bytes / event time)
  cpu_baseline         the oracle port of the reference (oracle/soccdpt_oracle.py) on the host cores, bounded sample

``--impl reference`` times the reference's algorithm on the host CPU (the oracle port; the unmodified
reference is a Python package that imports timm==0.6.12, which is absent from the image and the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "frames/s SOccDPT-V3 swin2_tiny_256 image->occupancy"
MODEL_TYPE = "dpt_swin2_tiny_256"
FLOPS_PER_FRAME = 78.8e9          # SURVEY.md 8d / BASELINE.md section 3 (2*MAC, analytic)
CAM_H, CAM_W, GRID, NCLS = 1080, 1920, (256, 256, 32), 3


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region: NVML polled every 5 ms from a thread
    (nvidia-smi -lms cannot sample faster than ~100 ms; falls back to it if pynvml is missing)."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40), ("hw_power_brake", 0x80))

    def __init__(self, index):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[0].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None

    def stop(self):
        self._stop.set()
        if self._thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self._thread.join(timeout=1.0)
        reasons = sorted(n for n, bit in self.REASONS if self.mask & bit)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm)}


def build_net(device, batch_hint=None):
    from soccdpt_b200 import SOccDPT_versions, load_model
    from soccdpt_b200.synthetic import seeded_state_dict, write_calib_yaml
    yml = write_calib_yaml(os.path.join("/tmp", f"soccdpt_bench_calib_{os.getpid()}.yaml"))
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        net = load_model(arch=SOccDPT_versions[3],
                         model_kwargs=dict(load_depth=False, num_classes=NCLS, sigmoid=True, compute_occ=True,
                                           camera_intrinsics_yaml=yml, model_type=MODEL_TYPE),
                         device=torch.device("cpu"), model_path=None, model_type=MODEL_TYPE)
    sd = seeded_state_dict(net.state_dict(), 0)
    net.load_state_dict(sd, strict=True)
    net.to(device).eval()
    return net, sd


def cpu_reference_fps(sd, frames, reps, threads=None):
    """The reference's algorithm on the host cores (oracle port), frames/s over `reps` batches of `frames`."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import soccdpt_oracle as O
    from soccdpt_b200.synthetic import synthetic_frames
    if threads:
        torch.set_num_threads(threads)
    orc = O.OracleV3(sd, MODEL_TYPE, sigmoid=True, geom=O.Geometry(), compute_occ=True)
    x = synthetic_frames(frames, 256, 0)
    orc(x[:1])                                    # warm-up (thread pools, allocator)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc(x)
        times.append(time.perf_counter() - t0)
    return frames / statistics.median(times), times


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from soccdpt_b200.synthetic import seeded_state_dict
    sd = _tiny_state_dict()
    frames = args.ref_frames
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import soccdpt_oracle as O
    from soccdpt_b200.synthetic import synthetic_frames
    orc = O.OracleV3(sd, MODEL_TYPE, sigmoid=True, geom=O.Geometry(), compute_occ=True)
    x = synthetic_frames(frames, 256, 0)
    for _ in range(args.warmup):
        orc(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc(x)
    dt = time.perf_counter() - t0
    fps = frames * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SOccDPT V3 {MODEL_TYPE} image->depth+seg+points+occupancy, camera {CAM_W}x{CAM_H}, "
                               f"grid {GRID}, {frames} frames per step on the host CPU (oracle port of the reference)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps x {frames} frames, fp32, torch {torch.__version__} CPU + C voxeliser"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def _tiny_state_dict():
    """Seeded weights without building the CUDA model (reference arm)."""
    from soccdpt_b200 import SOccDPT_versions, load_model
    from soccdpt_b200.synthetic import seeded_state_dict, write_calib_yaml
    import contextlib
    import io
    yml = write_calib_yaml(os.path.join("/tmp", f"soccdpt_bench_calib_{os.getpid()}.yaml"))
    with contextlib.redirect_stdout(io.StringIO()):
        net = load_model(arch=SOccDPT_versions[3],
                         model_kwargs=dict(load_depth=False, num_classes=NCLS, sigmoid=True, compute_occ=True,
                                           camera_intrinsics_yaml=yml, model_type=MODEL_TYPE),
                         device=torch.device("cpu"), model_path=None, model_type=MODEL_TYPE)
    return seeded_state_dict(net.state_dict(), 0)


def conv_flops(c):
    return 2.0 * c.N * c.H * c.W * c.Cout * c.KH * c.KW * c.Cin + 2.0 * c.N * c.H * c.W * c.Cout * c.proj_n


def instrumented_pass(net, x, reps):
    """Per-kernel CUDA-event times of one step (events bracket every launch on the launching stream);
    one untimed pass first, then the per-kernel MEDIAN over `reps` passes."""
    from soccdpt_b200 import _cabi
    eng = net.engine()
    plan = eng.plan_for(x.shape[0], x.device)
    stream = _cabi.current_stream()
    samples = {}
    meta = {}
    PP = "postprocess(unproject_scatter+grid_expand)"
    for rep in range(reps + 1):
        plan["x_in"].copy_(x)
        evs = []
        for op in plan["ops"]:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            op(stream)
            e.record()
            evs.append((op, s, e))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = net.get_semantic_occupancy(plan["depth"], plan["seg"])
        e.record()
        torch.cuda.synchronize()
        del out
        if rep == 0:
            continue
        tot = {}
        for op, a, b in evs:
            name = "conv_tcgen05_kernel" if op.name == "conv" else op.name
            tot[name] = tot.get(name, 0.0) + a.elapsed_time(b)
            m = meta.setdefault(name, {"launches": 0, "flops": 0.0})
            if rep == 1:
                m["launches"] += 1
                if op.name == "conv":
                    m["flops"] += conv_flops(op.args[0]._obj)
        tot[PP] = s.elapsed_time(e)
        meta.setdefault(PP, {"launches": 2, "flops": 0.0})
        for k, v in tot.items():
            samples.setdefault(k, []).append(v)
    return {k: {"ms": statistics.median(v), "launches": meta[k]["launches"], "flops": meta[k]["flops"]}
            for k, v in samples.items()}


_emit = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step (BASELINE config 2: 64)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-frames", type=int, default=2, help="frames per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly ONE line (the JSON): everything libraries print while the run is set up (NCCL's version banner,
    # model-loading messages) goes to stderr -- at the file-descriptor level, so native code is covered too
    global _emit
    sys.stdout.flush()
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def _emit(line):
        sys.stdout.flush()
        os.dup2(_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from soccdpt_b200 import _cabi
    from soccdpt_b200.synthetic import synthetic_frames
    net, sd = build_net(dev)
    B = args.batch
    x_host = synthetic_frames(B, 256, seed=rank).pin_memory()
    x = x_host.to(dev)

    def step_resident():
        return net(x)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = step_resident()
        barrier()
        launches0 = _cabi.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            out = step_resident()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        clocks = sampler.stop() if rank == 0 else None
        launches = _cabi.launch_count() - launches0

        # ---- end to end through the public host-frame API (soccdpt_b200.pipeline.FrameStream): pinned host
        # images in, results back in pinned host buffers, every step; upload / compute / download overlap on
        # three streams with double buffering.  Timed from the first H2D to the last D2H with CUDA events.
        from soccdpt_b200.pipeline import FrameStream
        fs = FrameStream(net, B, dev)

        def host_batches(n):
            for _ in range(n):
                yield x_host

        for _ in fs.run(host_batches(3)):
            pass
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record(fs.up)
        for _ in fs.run(host_batches(args.steps)):
            pass
        e2.record(fs.down)
        barrier()
        ms_e2e = s2.elapsed_time(e2)
        h2d, d2h = fs.h2d_bytes, fs.d2h_bytes

        agg = instrumented_pass(net, x, 5) if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    hbm_peak, tf_sustained, tf_burst, peak_kind = _peaks()
    frames_total = B * world * args.steps
    value = frames_total / (ms * 1e-3)
    e2e_value = frames_total / (ms_e2e * 1e-3)

    conv = agg["conv_tcgen05_kernel"]
    traffic = None          # DRAM bytes of all conv launches of one step, from the committed ncu capture of this command
    tp = os.path.join(ROOT, "profiles", "r1_conv_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            t = json.load(f)
        if t.get("batch") == B:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    conv_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
    step_kernel_ms = sum(d["ms"] for d in agg.values())
    pp = agg["postprocess(unproject_scatter+grid_expand)"]
    pp_bytes = B * (4 * 256 * 256 * (1 + NCLS) + 4 * CAM_H * CAM_W * (1 + NCLS) + 12 * CAM_H * CAM_W
                    + 4 * GRID[0] * GRID[1] * GRID[2] * NCLS)
    pp_gbs = pp_bytes / (pp["ms"] * 1e-3) / 1e9

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # timed on rank 0 at N = 1 only (torchrun pins every rank to one thread)
        fps, times = cpu_reference_fps(sd, frames=2, reps=3)
        cpu = {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "3 x 2 frames of the same workload, fp32 oracle port (torch CPU + C voxeliser), median"}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"SOccDPT V3 {MODEL_TYPE} inference, batch {B} synthetic 256x256 frames per GPU, "
                               f"image->depth+seg+points+occupancy (camera {CAM_W}x{CAM_H}, grid {GRID}, fp32 outputs 83 MB/frame)",
                   "per_gpu_batch": B, "parallelism": f"frame-sharded dp{world}, no data-path collective",
                   "l2": "per-step working set (>5 GB of outputs + activations) exceeds the 126 MB L2; no explicit flush",
                   "weights": "seeded random init (soccdpt_b200.synthetic.seeded_state_dict, seed 0)"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tcgen05_kernel", "achieved": conv_tflops, "peak": tf_sustained,
                     "unit": "TFLOP/s", "frac": conv_tflops / tf_sustained, "traffic": traffic,
                     "traffic_note": "sum of dram__bytes_read+write over the 76 conv launches of one step (profiles/r1_conv_traffic.json)",
                     "launches_per_step": conv["launches"], "ms_per_step": conv["ms"],
                     "share_of_step_kernel_time": conv["ms"] / step_kernel_ms, "peak_source": peak_kind + " (sustained bf16)"},
        "roofline_voxeliser": {"bound": "hbm", "kernel": "unproject_scatter_kernel+grid_expand_kernel", "achieved": pp_gbs,
                               "peak": hbm_peak, "unit": "GB/s", "frac": pp_gbs / hbm_peak, "ms_per_step": pp["ms"],
                               "bytes_per_step": pp_bytes, "peak_source": peak_kind},
        "model_tflops": value / world * FLOPS_PER_FRAME / 1e12,
        "model_frac_of_peak": value / world * FLOPS_PER_FRAME / 1e12 / tf_sustained,
        "kernels_ms_per_step": {k: round(v["ms"], 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
        "cpu_baseline": cpu,
    }
    _emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
