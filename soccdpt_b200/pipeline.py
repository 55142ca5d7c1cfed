"""Frame-stream driver around a SOccDPT model: the public end-to-end path for host-resident frames.

    stream = FrameStream(net, batch=64)
    for result in stream.run(host_batches):      # host_batches: iterable of pinned (B,3,S,S) fp32 tensors
        result.inv_depth, result.segmentation    # pinned host tensors at network resolution
        result.occupancy                          # pinned host (G0,G1,G2,C) grid of the call (union over the batch)
    stream = FrameStream(net, batch=64, frames="u8", frame_shape=(256, 256), result="packed")   # 4x / 2.7x fewer bytes each way

Three CUDA streams (upload / compute / download) and double-buffered device + pinned host buffers overlap
the host->device copy of batch i+1 and the device->host copy of batch i-1 with the kernels of batch i.
``overlap_post=True`` (or SOCCDPT_STREAM_OVERLAP=1) adds a fourth stream that runs the HBM-bound post-processing of batch i
under the network of batch i+1 (the network's depth / class maps are copied out of the engine's static buffers first).  It is
OFF by default: measured on B200 it changes nothing (9237 vs 9310 frames/s end to end, profiles/r2b_stream_overlap.log) -- the
persistent network kernels hold all 148 SMs with ~226 KB of shared memory each, so the voxeliser's CTAs find no SM to share.
Multi-GPU: one process per GPU, each rank feeds its own shard of the frame stream (`shard_range`); there is no
collective on the data path.  `gather_masks` is the optional exchange step for callers that want the
reference's union-over-batch occupancy ACROSS shards (an OR over the ranks' grids).
"""
from collections import namedtuple

import torch

FrameResult = namedtuple("FrameResult", "inv_depth segmentation occupancy index")


def shard_range(total, rank, world):
    """Contiguous, balanced split of `total` frames (or batches) over `world` ranks: [begin, end)."""
    assert 0 <= rank < world
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_masks(grid, group=None):
    """OR-reduce a binary occupancy grid over the process group (reference_union across shards).
    Works with NCCL on CUDA tensors and with gloo on CPU tensors (values are {0,1} -> MAX == OR)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return grid
    out = grid.clone()
    dist.all_reduce(out, op=dist.ReduceOp.MAX, group=group)
    return out


class FrameStream:
    """Double-buffered host <-> device pipeline around ``net``.

    input   ``frames="fp32"``: pinned (B,3,S,S) fp32 tensors at network resolution (what the reference's CPU transform yields);
            ``frames="u8"`` with ``frame_shape=(H, W)``: pinned uint8 (B,H,W,3) frames (camera frames, or frames already at
            network resolution: 4x fewer bytes than fp32), resized / normalised on the device by the input-pipeline kernel
            (soccdpt_b200.preprocess.GpuTransform = the reference's load_transforms).
    result  ``result="dense"``: network-resolution inverse depth + class maps (fp32) and the dense fp32 occupancy grid of the
            call (reference_union: the B grid copies are identical, one is read back; per_frame: all B);
            ``result="packed"``: the same maps as bf16 and the voxeliser's bit-packed occupancy mask (1 MB instead of 25 MB per
            grid; soccdpt_b200.occupancy.packed_to_points reads it) -- 2.7x fewer device -> host bytes per step.
    The reference's full 4-tuple (camera-resolution maps + points, 83 MB per frame) is produced on the device by ``net(x)``
    either way; it is not copied to the host by this class.
    """

    def __init__(self, net, batch, device=None, camera_frames=None, frames=None, frame_shape=None, result="dense",
                 overlap_post=None):
        import os
        from . import _cabi
        if overlap_post is None:
            overlap_post = os.environ.get("SOCCDPT_STREAM_OVERLAP", "0") == "1"
        self.overlap_post = bool(overlap_post)
        self.net = net
        self.batch = batch
        self.device = _cabi.normalize_device(device if device is not None else next(net.parameters()).device)
        if camera_frames is not None:                     # round-1 spelling: uint8 camera frames of this (H, W)
            frames, frame_shape = "u8", camera_frames
        frames = frames or "fp32"
        assert frames in ("fp32", "u8") and result in ("dense", "packed")
        if not getattr(net, "compute_occ", False):
            raise ValueError("FrameStream returns the occupancy of every call: construct the model with compute_occ=True")
        self.result = result
        self.per_frame = net.occupancy_mode == "per_frame"
        img = net.depth_net.pretrained.model.img_size
        C, G = net.num_classes, net.grid_size
        dev = self.device
        self.up, self.comp, self.down = (torch.cuda.Stream(dev) for _ in range(3))
        self.post = torch.cuda.Stream(dev) if self.overlap_post else self.comp
        self.x_dev = [torch.empty((batch, 3, img, img), dtype=torch.float32, device=dev) for _ in range(2)]
        self.transform, self.u8_dev = None, None
        if frames == "u8":
            from .preprocess import GpuTransform
            assert frame_shape is not None, "frames='u8' needs frame_shape=(H, W)"
            fh, fw = frame_shape
            self.transform = GpuTransform(img, img, keep_aspect_ratio=False)
            self.u8_dev = [torch.empty((batch, fh, fw, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        map_dtype = torch.float32 if result == "dense" else torch.bfloat16
        ng = batch if self.per_frame else 1
        if result == "dense":
            gshape, gdtype = (ng, G[0], G[1], G[2], C), torch.float32
        else:
            from .occupancy import mask_words
            gshape, gdtype = (ng, mask_words(G)), torch.int32
        # the voxeliser's output form follows the stream's result form (the model's own setting is restored after every call)
        self._occ_output = "dense" if result == "dense" else "packed"
        self.d_dev = [torch.empty((batch, img, img), dtype=map_dtype, device=dev) for _ in range(2)]
        self.s_dev = [torch.empty((batch, C, img, img), dtype=map_dtype, device=dev) for _ in range(2)]
        self.g_dev = [torch.empty(gshape, dtype=gdtype, device=dev) for _ in range(2)]
        self.d_host = [torch.empty((batch, img, img), dtype=map_dtype).pin_memory() for _ in range(2)]
        self.s_host = [torch.empty((batch, C, img, img), dtype=map_dtype).pin_memory() for _ in range(2)]
        self.g_host = [torch.empty(gshape, dtype=gdtype).pin_memory() for _ in range(2)]
        if self.overlap_post:      # the network outputs of batch i outlive the launch plan's static buffers
            self.nd_dev = [torch.empty((batch, img, img), dtype=torch.float32, device=dev) for _ in range(2)]
            self.ns_dev = [torch.empty((batch, C, img, img), dtype=torch.float32, device=dev) for _ in range(2)]
        stage = self.u8_dev[0] if self.u8_dev is not None else self.x_dev[0]
        self.h2d_bytes = stage.numel() * stage.element_size()
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in (self.d_host[0], self.s_host[0], self.g_host[0]))

    def _to_maps(self, src, dst):
        """fp32 network output -> the stream's map dtype (bf16 through the library's conversion kernel)."""
        if dst.dtype == torch.float32:
            dst.copy_(src, non_blocking=True)
        else:
            from . import _cabi
            _cabi.check(_cabi.load().soccdpt_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(),
                                                         _cabi.current_stream(self.device)), "f32_to_bf16")

    @torch.no_grad()
    def run(self, host_batches):
        """Yields one FrameResult per input batch, in order.  A result's host buffers are reused two batches later."""
        ev_up = [torch.cuda.Event() for _ in range(2)]
        ev_comp = [torch.cuda.Event() for _ in range(2)]
        ev_net = [torch.cuda.Event() for _ in range(2)]
        ev_down = [torch.cuda.Event() for _ in range(2)]
        ev_free_x = [None, None]       # compute finished reading x_dev[slot]
        pending = []
        net = self.net
        with torch.cuda.device(self.device):
            for i, xh in enumerate(host_batches):
                slot = i & 1
                stage_in = self.u8_dev if self.u8_dev is not None else self.x_dev
                assert tuple(xh.shape) == tuple(stage_in[0].shape) and xh.dtype == stage_in[0].dtype, \
                    "every batch must have the FrameStream's shape and dtype"
                with torch.cuda.stream(self.up):
                    if ev_free_x[slot] is not None:
                        self.up.wait_event(ev_free_x[slot])
                    stage_in[slot].copy_(xh, non_blocking=True)
                    ev_up[slot].record(self.up)
                with torch.cuda.stream(self.comp):
                    self.comp.wait_event(ev_up[slot])
                    if self.transform is not None:
                        self.transform(self.u8_dev[slot], out=self.x_dev[slot])
                    depth, seg = net.network(self.x_dev[slot])      # the buffers this call filled (static engine buffers)
                    if self.overlap_post:
                        if i >= 2:
                            self.comp.wait_event(ev_comp[slot])     # post-processing of batch i-2 has read nd_dev / ns_dev[slot]
                        self.nd_dev[slot].copy_(depth.reshape(self.nd_dev[slot].shape), non_blocking=True)
                        self.ns_dev[slot].copy_(seg.reshape(self.ns_dev[slot].shape), non_blocking=True)
                        depth, seg = self.nd_dev[slot], self.ns_dev[slot]
                    ev_net[slot].record(self.comp)
                    ev_free_x[slot] = ev_net[slot]
                with torch.cuda.stream(self.post):
                    self.post.wait_event(ev_net[slot])
                    if i >= 2:
                        self.post.wait_event(ev_down[slot])      # download of batch i-2 has drained this slot
                    saved = net.occupancy_output
                    net.occupancy_output = self._occ_output
                    try:
                        occ = net.get_semantic_occupancy(depth, seg)[3]
                    finally:
                        net.occupancy_output = saved
                    self._to_maps(depth, self.d_dev[slot])
                    self._to_maps(seg, self.s_dev[slot])
                    self.g_dev[slot].copy_(occ.reshape(self.g_dev[slot].shape) if self.per_frame or self.result == "packed"
                                           else occ[:1], non_blocking=True)
                    ev_comp[slot].record(self.post)
                with torch.cuda.stream(self.down):
                    self.down.wait_event(ev_comp[slot])
                    self.d_host[slot].copy_(self.d_dev[slot], non_blocking=True)
                    self.s_host[slot].copy_(self.s_dev[slot], non_blocking=True)
                    self.g_host[slot].copy_(self.g_dev[slot], non_blocking=True)
                    ev_down[slot].record(self.down)
                pending.append((i, slot))
                if len(pending) == 2:
                    j, sj = pending.pop(0)
                    ev_down[sj].synchronize()
                    yield self._result(sj, j)
            for j, sj in pending:
                ev_down[sj].synchronize()
                yield self._result(sj, j)

    def _result(self, slot, index):
        g = self.g_host[slot]
        return FrameResult(self.d_host[slot], self.s_host[slot], g if self.per_frame else g[0], index)
