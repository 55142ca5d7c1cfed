"""Frame-stream driver around a SOccDPT model: the public end-to-end path for host-resident frames.

    stream = FrameStream(net, batch=64)
    for result in stream.run(host_batches):      # host_batches: iterable of pinned (B,3,S,S) fp32 tensors
        result.inv_depth, result.segmentation    # pinned host tensors at network resolution
        result.occupancy                          # pinned host (G0,G1,G2,C) grid of the call (union over the batch)

Three CUDA streams (upload / compute / download) and double-buffered device + pinned host buffers overlap
the host->device copy of batch i+1 and the device->host copy of batch i-1 with the kernels of batch i.
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
    """`camera_frames=(H, W)`: the batches are pinned uint8 (B, H, W, 3) camera frames instead of network-resolution fp32
    tensors; they are uploaded as they are and resized / normalised on the device by the input-pipeline kernel
    (soccdpt_b200.preprocess.GpuTransform, the reference's load_transforms)."""

    def __init__(self, net, batch, device=None, camera_frames=None):
        self.net = net
        self.batch = batch
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("FrameStream needs a CUDA device (the hot path has no CPU fallback)")
        img = net.depth_net.pretrained.model.img_size
        C, G = net.num_classes, net.grid_size
        self.up, self.comp, self.down = (torch.cuda.Stream(self.device) for _ in range(3))
        self.x_dev = [torch.empty((batch, 3, img, img), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.transform, self.u8_dev = None, None
        if camera_frames is not None:
            from .preprocess import GpuTransform
            fh, fw = camera_frames
            self.transform = GpuTransform(img, img, keep_aspect_ratio=False)
            self.u8_dev = [torch.empty((batch, fh, fw, 3), dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.d_dev = [torch.empty((batch, img, img), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.s_dev = [torch.empty((batch, C, img, img), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.g_dev = [torch.empty((G[0], G[1], G[2], C), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.d_host = [torch.empty((batch, img, img), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.s_host = [torch.empty((batch, C, img, img), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.g_host = [torch.empty((G[0], G[1], G[2], C), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.h2d_bytes = self.u8_dev[0].numel() if self.u8_dev is not None else self.x_dev[0].numel() * 4
        self.d2h_bytes = (self.d_host[0].numel() + self.s_host[0].numel() + self.g_host[0].numel()) * 4

    @torch.no_grad()
    def run(self, host_batches):
        """Yields one FrameResult per input batch, in order.  A result's host buffers are reused two batches later."""
        ev_up = [torch.cuda.Event() for _ in range(2)]
        ev_comp = [torch.cuda.Event() for _ in range(2)]
        ev_down = [torch.cuda.Event() for _ in range(2)]
        ev_free_x = [None, None]       # compute finished reading x_dev[slot]
        pending = []
        i = -1
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
                if i >= 2:
                    self.comp.wait_event(ev_down[slot])      # download of batch i-2 has drained this slot
                if self.transform is not None:
                    self.transform(self.u8_dev[slot], out=self.x_dev[slot])
                out = self.net(self.x_dev[slot])
                depth, seg = self.net.network_outputs(self.batch, self.device)
                self.d_dev[slot].copy_(depth, non_blocking=True)
                self.s_dev[slot].copy_(seg, non_blocking=True)
                self.g_dev[slot].copy_(out[3][0], non_blocking=True)
                ev_comp[slot].record(self.comp)
                ev_free_x[slot] = ev_comp[slot]
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
                yield FrameResult(self.d_host[sj], self.s_host[sj], self.g_host[sj], j)
        for j, sj in pending:
            ev_down[sj].synchronize()
            yield FrameResult(self.d_host[sj], self.s_host[sj], self.g_host[sj], j)
