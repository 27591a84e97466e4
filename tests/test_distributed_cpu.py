"""CPU-only (gloo, world_size 2): the host-side multi-rank logic — tile ownership is a partition of the frame,
and the per-frame combine (one sum-reduce of buffers that are zero outside a rank's pixels) reassembles it."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from path_tracer_ai_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, tile, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        owner = D.owner_map(W, H, world, tile)
        # what b2pt_render writes for this rank: a value for owned pixels, 0 elsewhere
        truth = np.arange(W * H * 3, dtype=np.float32).reshape(H, W, 3) * np.float32(0.25) + np.float32(1.0)
        mine = np.where((owner == rank)[..., None], truth, np.float32(0.0)).astype(np.float32)
        buf = torch.from_numpy(mine.copy())
        D.combine_frames(buf, dst=0)
        # sample-range partition: partial means add up to the full mean
        part = D.sample_partition(rank, world, 5)
        assert part == dict(sample_begin=5 * rank, sample_count=5)
        partial = torch.full((4,), float(rank + 1) / world)
        D.combine_frames(partial, dst=0)
        if rank == 0:
            ok = np.array_equal(buf.numpy(), truth) and torch.allclose(partial, torch.full((4,), sum(range(1, world + 1)) / world))
            out.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W,H,tile", [(64, 36, 4), (33, 17, 8)])
def test_tile_partition_and_combine_gloo_world2(W, H, tile):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port() if r < 0 else PORT, W, H, tile, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


PORT = _free_port()


def test_owner_map_is_a_partition():
    for world in (1, 2, 3, 8):
        om = D.owner_map(50, 30, world, 4)
        counts = [D.owned_pixel_count(50, 30, r, world, 4) for r in range(world)]
        assert sum(counts) == 50 * 30 and om.min() == 0 and om.max() == world - 1
        assert max(counts) - min(counts) <= 16      # interleaving keeps the load balanced to within one run
    assert D.tile_partition(0, 1) is None and D.sample_partition(0, 1, 7) is None
    assert D.tile_partition(3, 8, 16) == dict(tile_rank=3, tile_world=8, tile_size=16)
