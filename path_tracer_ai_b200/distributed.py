"""Multi-GPU plumbing: one process per GPU (torch.distributed), scene replicated, frame partitioned, one
sum-reduce per frame.  The path shards without any data-path collective: the only exchange is the final
frame combine.

Partitions mirror `b2pt_partition` (include/b2pt.h):
* tiles   — runs of tile_size^2 consecutive pixels (row-major); run k belongs to rank k % world.  Every pixel is
            computed wholly by one rank with Philox keyed by (pixel, sample): the combined image is bit-identical
            for any world size.
* samples — rank r traces samples [r*spp, (r+1)*spp) of a frame with spp*world samples per pixel (weak scaling).
Pixels a rank does not own are written as 0 and every rank divides by the frame's total spp, so the combine is a
plain SUM.
"""
from __future__ import annotations

import numpy as np


def tile_partition(rank: int, world: int, tile_size: int = 32) -> dict | None:
    return None if world <= 1 else dict(tile_rank=rank, tile_world=world, tile_size=tile_size)


def sample_partition(rank: int, world: int, spp_per_rank: int) -> dict | None:
    return None if world <= 1 else dict(sample_begin=rank * spp_per_rank, sample_count=spp_per_rank)


def owner_map(width: int, height: int, world: int, tile_size: int = 32) -> np.ndarray:
    """owner[y, x] = rank that renders pixel (x, y) under the tile partition (same rule as render.cu pixel_of)."""
    idx = np.arange(width * height, dtype=np.int64)
    if world <= 1:
        return np.zeros((height, width), np.int32)
    return ((idx // (tile_size * tile_size)) % world).astype(np.int32).reshape(height, width)


def owned_pixel_count(width: int, height: int, rank: int, world: int, tile_size: int = 32) -> int:
    return int((owner_map(width, height, world, tile_size) == rank).sum())


def combine_frames(buf, dst: int = 0, group=None):
    """One collective per frame: SUM-reduce the per-rank W*H*3 buffers onto `dst` (NCCL for CUDA tensors, gloo for
    CPU tensors).  No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return buf


def render_distributed(engine, cam, width, height, spp_total, bounces, d_rgb, part, seed=1234, dst=0, group=None):
    """Renders this rank's share into the CUDA tensor `d_rgb` (W*H*3 float32) and combines on `dst`.

    Stream contract (include/b2pt.h): the engine works on its own non-blocking stream and returns when the frame is
    complete, but it does not order itself after work queued on other streams — so whatever torch still has in flight on
    `d_rgb` (the previous frame's reduce, a fill) is waited for first."""
    if d_rgb.is_cuda:
        import torch
        torch.cuda.current_stream(d_rgb.device).synchronize()
    engine.render_device(cam, width, height, spp_total, bounces, d_rgb.data_ptr(), seed=seed, part=part)
    stats = engine.stats()
    combine_frames(d_rgb, dst, group)
    return stats
