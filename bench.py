#!/usr/bin/env python
"""Benchmark of the SOccDPT inference hot path (BASELINE.json metric):

    frames/s, SOccDPT-V3 dpt_swin2_tiny_256, image -> (inverse depth, segmentation, points, occupancy grid)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--model tiny|base_384|hybrid_384]
                    [--impl ours|reference] [--stream-frames F] [--voxeliser-sweep] [--gather]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); frames are batch-sharded, every rank runs the same per-GPU
batch (weak scaling), there is no collective on the data path.  A "step" is one ``net(x)`` call on one batch of synthetic
frames (random seeded weights, randn images, the synthetic 1920x1080 pinhole camera of SURVEY.md 8d).  Rank 0 prints ONE
JSON line.

  value     frames/s with the input batch already resident in HBM (CUDA events, max over ranks)
  e2e       frames/s through the public host-frame API (soccdpt_b200.pipeline.FrameStream) with the batch in PINNED HOST
            memory: every step copies the frames host->device (uint8 at network resolution, normalised on the device by the
            input-pipeline kernel) and reads the step's result back device->host (network-resolution inverse depth + class
            maps as bf16, and the bit-packed occupancy mask of the call); copies overlap with compute on separate streams,
            the timed region spans first H2D to last D2H.  ``net(x)`` still produces the reference's full fp32 4-tuple
            (83 MB per frame) on the device every step: what is REDUCED is only what crosses PCIe (stated in the e2e dict).
  roofline  tensor-pipe roofline of the dominant kernel (conv_tcgen05_kernel): algorithmic FLOPs of all its launches in a
            step / their CUDA-event time, vs MEASURED_PEAKS.json bf16 sustained
  roofline_voxeliser   HBM roofline of the post-processing kernels (compulsory bytes / event time)
  roofline_block_tail  HBM roofline of the fused Swin block tails (x in, master in/out, y out: 12 bytes per element)
  cpu_baseline         the reference's algorithm on the host cores, bounded sample: oracle port of the network (torch fp32 CPU)
                       + the op-for-op ATen restatement of its post-processing (oracle/torch_postprocess.py: masked_select /
                       nonzero / index_put_, what the reference really executes); cpu_baseline_c_voxeliser is the same network
                       with the C voxeliser (round 1's baseline, ~3x faster than the reference's own post-processing)

``--model`` runs BASELINE configs 3 / 4 (dpt_swin2_base_384, dpt_hybrid_384); ``--stream-frames F`` is config 3's frame
stream: F frames split over the ranks (strong scaling), each rank feeding its shard through FrameStream; ``--voxeliser-sweep``
is config 5 (grid resolution x batch, camera-resolution and network-resolution maps, parity asserted against the C oracle).
``--impl reference`` times the reference's algorithm on the host CPU (the oracle port; the unmodified reference is a Python
package that imports timm==0.6.12, which is absent from the image and the GPU box).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "frames/s SOccDPT-V3 swin2_tiny_256 image->occupancy"
# model key -> (model_type, analytic FLOPs per frame (BASELINE.md section 3, 2*MAC), network input size, default per-GPU batch)
MODELS = {
    "tiny": ("dpt_swin2_tiny_256", 78.8e9, 256, 64),
    "base_384": ("dpt_swin2_base_384", 259.7e9, 384, 32),
    "hybrid_384": ("dpt_hybrid_384", 298.8e9, 384, 32),
}
CAM_H, CAM_W, GRID, NCLS = 1080, 1920, (256, 256, 32), 3


def metric_name(model):
    return METRIC if model == "tiny" else f"frames/s SOccDPT-V3 {MODELS[model][0][4:]} image->occupancy"


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


def _quiet_load(model_type, device):
    from soccdpt_b200 import SOccDPT_versions, load_model
    from soccdpt_b200.synthetic import write_calib_yaml
    import contextlib
    import io
    yml = write_calib_yaml(os.path.join("/tmp", f"soccdpt_bench_calib_{os.getpid()}.yaml"))
    with contextlib.redirect_stdout(io.StringIO()):
        return load_model(arch=SOccDPT_versions[3],
                          model_kwargs=dict(load_depth=False, num_classes=NCLS, sigmoid=True, compute_occ=True,
                                            camera_intrinsics_yaml=yml, model_type=model_type),
                          device=device, model_path=None, model_type=model_type)


def build_net(device, model="tiny"):
    from soccdpt_b200.synthetic import seeded_state_dict
    model_type = MODELS[model][0]
    net = _quiet_load(model_type, torch.device("cpu"))
    # the hybrid's random-init ResNetV2 trunk is numerically chaotic under bf16 storage: damped residual
    # branches, like trained weights; throughput does not depend on the values
    sd = seeded_state_dict(net.state_dict(), 0, residual_gain=0.1 if model == "hybrid_384" else 1.0)
    net.load_state_dict(sd, strict=True)
    net.to(device).eval()
    return net, sd


# ------------------------------------------------------------------------------------------------ CPU legs (the checker, timed)
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import soccdpt_oracle as O
    import torch_postprocess as TP
    return O, TP


class CpuReference:
    """The reference's algorithm on the host: oracle port of the network + (faithful) the ATen restatement of its
    post-processing, or (c_voxeliser) the C voxeliser."""

    def __init__(self, sd, model_type):
        O, TP = _oracle()
        self.O, self.TP = O, TP
        self.orc = O.OracleV3(sd, model_type, sigmoid=True, geom=O.Geometry(), compute_occ=True)

    def __call__(self, x, faithful=True):
        if not faithful:
            return self.orc(x)
        d, g, _, _ = self.orc.network(x)
        return self.TP.get_semantic_occupancy(d, g, self.orc.geom, True)


def cpu_reference_fps(sd, model_type, img, frames, reps, faithful):
    from soccdpt_b200.synthetic import synthetic_frames
    ref = CpuReference(sd, model_type)
    x = synthetic_frames(frames, img, 0)
    ref(x[:1], faithful)                          # warm-up (thread pools, allocator)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ref(x, faithful)
        times.append(time.perf_counter() - t0)
    return frames / statistics.median(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from soccdpt_b200.synthetic import seeded_state_dict, synthetic_frames
    model_type, _, img, _ = MODELS[args.model]
    net = _quiet_load(model_type, torch.device("cpu"))
    sd = seeded_state_dict(net.state_dict(), 0, residual_gain=0.1 if args.model == "hybrid_384" else 1.0)
    del net
    frames = args.ref_frames
    ref = CpuReference(sd, model_type)
    x = synthetic_frames(frames, img, 0)
    for _ in range(args.warmup):
        ref(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref(x)
    dt = time.perf_counter() - t0
    fps = frames * args.steps / dt
    line = {
        "impl": "reference", "metric": metric_name(args.model), "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SOccDPT V3 {model_type} image->depth+seg+points+occupancy, camera {CAM_W}x{CAM_H}, "
                               f"grid {GRID}, {frames} frames per step on the host CPU (oracle port of the reference)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps x {frames} frames, fp32, torch {torch.__version__} CPU: network oracle + "
                                   "op-for-op ATen post-processing (masked_select / nonzero / index_put_)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------ per-kernel accounting
def conv_flops(c):
    """2*MAC of one soccdpt_conv_fwd launch; outputs live on the strided grid (Ho = ceil(H / stride))."""
    s = c.stride if c.stride > 1 else 1
    Ho, Wo = (c.H + s - 1) // s, (c.W + s - 1) // s
    return 2.0 * c.N * Ho * Wo * c.Cout * c.KH * c.KW * c.Cin + 2.0 * c.N * Ho * Wo * c.Cout * c.proj_n


def tail_flops_bytes(a):
    """(2*MAC, algorithmic HBM bytes) of one fused Swin block tail."""
    fl = 2.0 * a.M * (a.K1 * a.HID + a.HID * a.C) if a.HID else 2.0 * a.M * a.K1 * a.C
    return fl, a.M * (2.0 * a.K1 + a.C * (4 + 4 + 2))


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
        eng.bind_input(plan, x)
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
            name = "conv_tcgen05_kernel" if op.name == "conv" else (
                "swin_block_tail_kernel" if op.name.startswith("block_tail") else op.name)
            tot[name] = tot.get(name, 0.0) + a.elapsed_time(b)
            m = meta.setdefault(name, {"launches": 0, "flops": 0.0, "bytes": 0.0})
            if rep == 1:
                m["launches"] += 1
                if op.name == "conv":
                    m["flops"] += conv_flops(op.args[0]._obj)
                elif op.name.startswith("block_tail"):
                    fl, by = tail_flops_bytes(op.args[0]._obj)
                    m["flops"] += fl
                    m["bytes"] += by
        tot[PP] = s.elapsed_time(e)
        meta.setdefault(PP, {"launches": 2, "flops": 0.0, "bytes": 0.0})
        for k, v in tot.items():
            samples.setdefault(k, []).append(v)
    return {k: dict(meta[k], ms=statistics.median(v)) for k, v in samples.items()}


# ------------------------------------------------------------------------------------------------ config 5: voxeliser sweep
def voxeliser_sweep(dev):
    """BASELINE config 5: grid resolution x batch; camera-resolution maps (the stand-alone voxeliser) and
    256x256 network-resolution maps (resize fused in); outputs asserted bit-equal to the C oracle on the smallest batch of
    every grid; the reference's own (ATen) voxelisation is timed on the host for the 256x256x32 grid."""
    O, TP = _oracle()
    import numpy as np
    from soccdpt_b200 import SOccDPT
    from soccdpt_b200.synthetic import write_calib_yaml
    hbm_peak, _, _, peak_kind = _peaks()
    yml = write_calib_yaml(f"/tmp/soccdpt_sweep_calib_{os.getpid()}.yaml")
    grids = [((64, 64, 8), (0.5, 0.5, 0.1665)), ((128, 128, 16), (1.0, 1.0, 0.333)), ((256, 256, 32), (2.0, 2.0, 0.666)),
             ((512, 512, 64), (4.0, 4.0, 1.332))]
    rows = []
    for grid, scale in grids:
        net = SOccDPT(camera_intrinsics_yaml=yml, compute_occ=True, grid_size=grid, scale=scale)
        geom = O.Geometry(grid_size=grid, scale=scale)
        cells = grid[0] * grid[1] * grid[2] * NCLS
        for maps in ("camera", "network"):
            H, W = (CAM_H, CAM_W) if maps == "camera" else (256, 256)
            batches = (1, 8, 64) if cells <= 256 * 256 * 32 * NCLS else (1, 8, 16)      # 512x512x64: 201 MB per dense grid
            for B in batches:
                inv, seg = O.config5_maps(B, H, W, NCLS, seed=B)
                inv_d, seg_d = inv.to(dev), seg.to(dev)
                work = inv_d.clone()

                def call():
                    if maps == "camera":                  # clamps in place: idempotent, same work every repetition
                        return net.voxelize(work, seg_d)
                    out = net.get_semantic_occupancy(inv_d, seg_d)
                    return out[2], out[3]
                for _ in range(3):
                    pts, g = call()
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10 if B * cells < 2 ** 31 else 4
                s.record()
                for _ in range(reps):
                    pts, g = call()
                e.record()
                torch.cuda.synchronize()
                ms = s.elapsed_time(e) / reps
                # algorithmic bytes: maps in, clamped inverse depth (+ class maps) + points + dense grid out
                nbytes = B * (4 * H * W * (1 + NCLS) + 4 * CAM_H * CAM_W * (1 + (NCLS if maps == "network" else 0))
                              + 12 * CAM_H * CAM_W + 4 * cells)
                row = {"grid": list(grid), "maps": maps, "batch": B, "ms": ms, "frames_per_s": B / ms * 1e3,
                       "GBps": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / hbm_peak}
                if B == 1 and maps == "camera":           # the bit-exact stage (the fused resize is a tolerance stage)
                    t0 = time.perf_counter()
                    _, pts_o, grid_o = O.voxelize(inv.numpy(), seg.numpy(), geom)
                    row["c_oracle_frames_per_s"] = B / (time.perf_counter() - t0)
                    assert np.array_equal(pts.cpu().numpy().view(np.uint32), pts_o.view(np.uint32)), (grid, maps)
                    assert np.array_equal(g.cpu().numpy(), grid_o), (grid, maps)
                    row["parity"] = "bit-exact vs C oracle (points, grid)"
                    if grid == (256, 256, 32):
                        t0 = time.perf_counter()
                        _, _, grid_t = TP.voxelize(inv, seg, geom, device="cpu")
                        row["reference_aten_cpu_frames_per_s"] = B / (time.perf_counter() - t0)
                        assert torch.equal(grid_t, g.cpu())
                rows.append(row)
                del pts, g
        del net
        torch.cuda.empty_cache()
    best = max((r for r in rows if r["grid"] == [256, 256, 32] and r["maps"] == "camera"), key=lambda r: r["frac_of_hbm"])
    return {"metric": "frames/s stand-alone depth->occupancy voxeliser (BASELINE config 5)", "value": best["frames_per_s"],
            "unit": "frames/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "voxeliser sweep: grids 64x64x8 .. 512x512x64, batch 1/8/64, camera-resolution (1080x1920) and "
                                   "network-resolution (256x256, resize fused) maps, 3 classes; value = 256x256x32 grid, camera maps"},
            "roofline": {"bound": "hbm", "kernel": "unproject_scatter_kernel+grid_expand_kernel", "achieved": best["GBps"],
                         "peak": hbm_peak, "unit": "GB/s", "frac": best["frac_of_hbm"], "traffic": None, "peak_source": peak_kind},
            "cpu_cores": os.cpu_count(), "sweep": rows}


_emit = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--model", default="tiny", choices=sorted(MODELS))
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (default: 64 tiny / 32 base_384, hybrid_384)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-frames", type=int, default=2, help="frames per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stream-frames", type=int, default=0,
                    help="BASELINE config 3: a stream of this many frames split over the ranks (strong scaling), end to end")
    ap.add_argument("--voxeliser-sweep", action="store_true", help="BASELINE config 5 (one GPU)")
    ap.add_argument("--gather", action="store_true", help="also time the NCCL OR-gather of the occupancy masks (optional exchange)")
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
    if args.voxeliser_sweep:
        if rank == 0:
            _emit(voxeliser_sweep(dev))
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from soccdpt_b200 import _cabi
    from soccdpt_b200.pipeline import FrameStream, gather_masks, shard_range
    from soccdpt_b200.synthetic import synthetic_frames
    model_type, flops_per_frame, img, default_batch = MODELS[args.model]
    net, sd = build_net(dev, args.model)
    B = args.batch or default_batch
    x = synthetic_frames(B, img, seed=rank).to(dev)
    # the end-to-end input: uint8 frames at network resolution in pinned host memory (the reference's loaders hand over uint8
    # frames that are NOT divided by 255: bengaluru_driving_dataset.py:118-128), normalised on the device
    gen = torch.Generator().manual_seed(1000 + rank)
    u8_host = torch.randint(0, 256, (B, img, img, 3), dtype=torch.uint8, generator=gen).pin_memory()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def host_batches(n):
        for _ in range(n):
            yield u8_host

    fs = FrameStream(net, B, dev, frames="u8", frame_shape=(img, img), result="packed")
    e2e_note = {"input": f"uint8 frames at network resolution ({img}x{img}x3) from pinned host memory, normalised on the device",
                "result": "REDUCED: bf16 network-resolution inverse depth + class maps and the bit-packed occupancy mask of the "
                          "call; the reference's fp32 4-tuple at camera resolution (83 MB per frame) is produced on the device "
                          "every step but not copied to the host",
                "overlap": ("post-processing of batch i on its own stream under the network of batch i+1 (FrameStream overlap_post)"
                            if fs.overlap_post else "none: network and post-processing of a batch run back to back on one stream")}

    with torch.no_grad():
        if args.stream_frames:
            # ---------------- config 3: a frame stream split over the ranks, end to end through FrameStream (strong scaling)
            b0, b1 = shard_range(args.stream_frames // B, rank, world)          # whole batches per rank
            for _ in fs.run(host_batches(max(args.warmup, 3))):
                pass
            barrier()
            sampler = ClockSampler(local)
            if rank == 0:
                sampler.start()
            launches0 = _cabi.launch_count()
            s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s2.record(fs.up)
            for _ in fs.run(host_batches(b1 - b0)):
                pass
            e2.record(fs.down)
            barrier()
            (ms,) = reduce_max([s2.elapsed_time(e2)])
            clocks = sampler.stop() if rank == 0 else None
            if rank == 0:
                frames = (args.stream_frames // B) * B
                fps = frames / (ms * 1e-3)
                _emit({"metric": metric_name(args.model), "value": fps, "unit": "frames/s", "n_gpus": world,
                       "steps": (b1 - b0), "warmup": max(args.warmup, 3), "ms_per_step": ms / max(b1 - b0, 1),
                       "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                       "gpu_launches": int(_cabi.launch_count() - launches0), "clocks": clocks,
                       "config": {"workload": f"SOccDPT V3 {model_type} inference, stream of {frames} synthetic {img}x{img} frames "
                                              f"split over {world} GPUs in batches of {B} (BASELINE config 3), end to end through "
                                              "FrameStream", "per_gpu_batch": B, "frames_total": frames,
                                  "parallelism": f"frame-sharded dp{world}, no data-path collective",
                                  "l2": "per-step working set exceeds the 126 MB L2; no explicit flush"},
                       "e2e": dict({"value": fps, "unit": "frames/s", "h2d_bytes_per_step": fs.h2d_bytes,
                                    "d2h_bytes_per_step": fs.d2h_bytes}, **e2e_note),
                       "model_tflops_per_gpu": fps / world * flops_per_frame / 1e12})
            if world > 1:
                import torch.distributed as dist
                dist.destroy_process_group()
            return

        for _ in range(max(args.warmup, 3)):
            out = net(x)
        barrier()
        launches0 = _cabi.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            out = net(x)
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        clocks = sampler.stop() if rank == 0 else None
        launches = _cabi.launch_count() - launches0
        del out

        # ---- end to end through the public host-frame API (soccdpt_b200.pipeline.FrameStream): pinned host frames in,
        # results back in pinned host buffers, every step; upload / compute / download overlap on three streams with double
        # buffering.  Timed from the first H2D to the last D2H with CUDA events.
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

        ms_gather = None
        if args.gather and world > 1:
            # optional exchange step: the reference's union-over-batch occupancy ACROSS shards = an OR over the ranks' masks
            mask = torch.zeros(GRID + (NCLS,), dtype=torch.float32, device=dev)
            for _ in range(3):
                gather_masks(mask)
            barrier()
            s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s3.record()
            for _ in range(20):
                gather_masks(mask)
            e3.record()
            barrier()
            ms_gather = s3.elapsed_time(e3) / 20

        agg = instrumented_pass(net, x, 5) if rank == 0 else None

    vals = reduce_max([ms, ms_e2e] + ([ms_gather] if ms_gather is not None else []))
    ms, ms_e2e = vals[0], vals[1]
    if ms_gather is not None:
        ms_gather = vals[2]

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
    tp = os.path.join(ROOT, "profiles", "r2_conv_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            t = json.load(f)
        if t.get("batch") == B and t.get("model", "tiny") == args.model:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    conv_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
    step_kernel_ms = sum(d["ms"] for d in agg.values())
    pp = agg["postprocess(unproject_scatter+grid_expand)"]
    pp_bytes = B * (4 * img * img * (1 + NCLS) + 4 * CAM_H * CAM_W * (1 + NCLS) + 12 * CAM_H * CAM_W
                    + 4 * GRID[0] * GRID[1] * GRID[2] * NCLS)
    pp_gbs = pp_bytes / (pp["ms"] * 1e-3) / 1e9
    tail = agg.get("swin_block_tail_kernel")

    cpu = cpu_c = None
    if not args.no_cpu_baseline and world == 1:      # timed on rank 0 at N = 1 only (torchrun pins every rank to one thread)
        torch.set_num_threads(os.cpu_count() or 1)
        fps = cpu_reference_fps(sd, model_type, img, frames=2, reps=10, faithful=True)
        cpu = {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "10 x 2 frames of the same workload, fp32: oracle port of the network (torch CPU) + op-for-op ATen "
                         "restatement of the reference's post-processing (masked_select / nonzero / index_put_), median"}
        fps_c = cpu_reference_fps(sd, model_type, img, frames=2, reps=5, faithful=False)
        cpu_c = {"value": fps_c, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                 "sample": "5 x 2 frames, same network oracle with the C voxeliser (oracle/voxel_oracle.c) instead of the ATen ops"}

    line = {
        "metric": metric_name(args.model), "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"SOccDPT V3 {model_type} inference, batch {B} synthetic {img}x{img} frames per GPU, "
                               f"image->depth+seg+points+occupancy (camera {CAM_W}x{CAM_H}, grid {GRID}, fp32 outputs 83 MB/frame)",
                   "per_gpu_batch": B, "parallelism": f"frame-sharded dp{world}, no data-path collective",
                   "l2": "per-step working set (>5 GB of outputs + activations) exceeds the 126 MB L2; no explicit flush",
                   "weights": "seeded random init (soccdpt_b200.synthetic.seeded_state_dict, seed 0)"},
        "e2e": dict({"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": fs.h2d_bytes, "d2h_bytes_per_step": fs.d2h_bytes,
                     "ms_per_step": ms_e2e / args.steps}, **e2e_note),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tcgen05_kernel", "achieved": conv_tflops, "peak": tf_sustained,
                     "unit": "TFLOP/s", "frac": conv_tflops / tf_sustained, "traffic": traffic,
                     "traffic_note": "sum of dram__bytes_read+write over the conv launches of one step (profiles/r2_conv_traffic.json)",
                     "launches_per_step": conv["launches"], "ms_per_step": conv["ms"],
                     "share_of_step_kernel_time": conv["ms"] / step_kernel_ms, "peak_source": peak_kind + " (sustained bf16)"},
        "roofline_voxeliser": {"bound": "hbm", "kernel": "unproject_scatter_kernel+grid_expand_kernel", "achieved": pp_gbs,
                               "peak": hbm_peak, "unit": "GB/s", "frac": pp_gbs / hbm_peak, "ms_per_step": pp["ms"],
                               "bytes_per_step": pp_bytes, "peak_source": peak_kind},
        "model_tflops": value / world * flops_per_frame / 1e12,
        "model_frac_of_peak": value / world * flops_per_frame / 1e12 / tf_sustained,
        "kernels_ms_per_step": {k: round(v["ms"], 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
        "cpu_baseline": cpu,
        "cpu_baseline_c_voxeliser": cpu_c,
    }
    if tail is not None:
        gbs = tail["bytes"] / (tail["ms"] * 1e-3) / 1e9
        line["roofline_block_tail"] = {"bound": "hbm", "kernel": "swin_block_tail_kernel", "achieved": gbs, "peak": hbm_peak,
                                       "unit": "GB/s", "frac": gbs / hbm_peak, "launches_per_step": tail["launches"],
                                       "ms_per_step": tail["ms"], "tflops": tail["flops"] / (tail["ms"] * 1e-3) / 1e12,
                                       "peak_source": peak_kind}
    if ms_gather is not None:
        line["nccl_gather_masks_ms"] = ms_gather
    _emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
