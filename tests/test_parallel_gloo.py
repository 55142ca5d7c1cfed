"""Multi-GPU host logic on CPU: frame sharding and the optional cross-shard occupancy union, exercised with
world_size=2 over gloo (the data path itself has no collective)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as GU
import soccdpt_oracle as O
from soccdpt_b200.pipeline import gather_masks, shard_range


def test_shard_range_is_a_partition():
    for total in (0, 1, 7, 64, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z, calib, geom, inv, seg = GU.load_voxel_case("ragged_b3_grid64")     # 3 frames
    b, e = shard_range(inv.shape[0], rank, world)
    _, _, grid = O.voxelize(inv[b:e].numpy(), seg[b:e].numpy(), geom)     # this rank's shard (the oracle stands in
    local = torch.from_numpy(grid[0])                                      # for the GPU call in this CPU-only test)
    merged = gather_masks(local)
    q.put((rank, (b, e), merged.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_union_equals_single_call_union():
    world, port = 2, 29500 + (os.getpid() % 500)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z, calib, geom, inv, seg = GU.load_voxel_case("ragged_b3_grid64")
    _, _, full = O.voxelize(inv.numpy(), seg.numpy(), geom)
    spans = sorted(r[1] for r in results)
    assert spans == [(0, 2), (2, 3)]
    for _, _, merged in results:
        assert np.array_equal(merged, full[0])      # OR over shards == the reference's union over the whole batch
