"""Multi-GPU plumbing: one process per GPU (torchrun), scene replicated, SAMPLES sharded, one reduce.

The reference is single-GPU (no collective anywhere).  The path shards over
independent (pixel, sample) pairs: rank r traces sample indices [base_r, base_r + count_r) of every
pixel -- disjoint Philox sample indices, so the union over ranks is exactly the sample set a single
GPU would trace -- and the per-rank radiance sums (W*H*3 float32, 24.9 MB at 1080p) are added with
ONE reduce to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).  Nothing else crosses ranks.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple


def shard_samples(spp: int, rank: int, world: int, base: int = 0) -> Tuple[int, int]:
    """Contiguous split of `spp` sample indices; the first spp % world ranks get one more."""
    if world < 1 or not (0 <= rank < world) or spp < 0:
        raise ValueError("bad shard request spp=%d rank=%d world=%d" % (spp, rank, world))
    q, r = divmod(spp, world)
    count = q + (1 if rank < r else 0)
    start = rank * q + min(rank, r)
    return base + start, count


def shard_tiles(rank: int, world: int) -> Tuple[int, int]:
    """Interleaved tile sharding: rank r traces the 8x4-pixel tiles t with t % world == r (drb_opts.tile_rank /
    tile_count).  Every pixel's samples stay on one GPU in sample order, so the reduced image is bit-identical to
    the single-GPU image; neighbouring tiles go to different ranks, which balances the load."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard request rank=%d world=%d" % (rank, world))
    return rank, world


def tile_owner_mask(width: int, height: int, rank: int, world: int):
    """(H, W) bool array: the pixels `rank` owns under interleaved tile sharding -- tile t = (y // 4) * ceil(W / 8) +
    x // 8 belongs to rank t % world (slot_to_pixel / k_resolve in csrc/render.cu, drb_render_multi's merge)."""
    import numpy as np
    tile_rank, tile_count = shard_tiles(rank, world)
    tiles_x = (width + 7) // 8
    y, x = np.mgrid[0:height, 0:width]
    return ((y // 4) * tiles_x + x // 8) % tile_count == tile_rank


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend: str):
    """torch.distributed rendezvous from torchrun's environment (127.0.0.1 unless MASTER_ADDR is set)."""
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def render_sharded(render_fn: Callable[[int, int], "object"], spp: int, rank: int, world: int, sample_base: int = 0,
                   reduce_dst: Optional[int] = 0):
    """Trace this rank's share with render_fn(base, count) -> tensor of radiance SUMS, then reduce.

    A rank's share may be empty (spp < world): render_fn is then called with count == 0 and must return zeros --
    Scene.render_device(sample_count=0) does exactly that (an integer sample_count is taken literally,
    DRB_FLAG_EXACT_SAMPLES; only sample_count=None means "the settings' spp").

    Returns (tensor, count): on `reduce_dst` (or on every rank when reduce_dst is None -> all-reduce)
    the tensor holds the sum over all spp samples; elsewhere its contents are unspecified.
    """
    import torch.distributed as dist
    base, count = shard_samples(spp, rank, world, sample_base)
    acc = render_fn(base, count)
    if world > 1:
        if reduce_dst is None:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        else:
            dist.reduce(acc, dst=reduce_dst, op=dist.ReduceOp.SUM)
    return acc, count


def render_progressive(render_fn: Callable[[int, int], "object"], spp: int, chunk: int, rank: int, world: int, sample_base: int = 0,
                       reduce_dst: Optional[int] = 0):
    """Progressive rendering with a periodic reduce: the headless, multi-GPU stand-in for the reference's
    accumulate-and-redraw loop (kernel.cu:2154-2224).

    The frame's `spp` sample indices are consumed `chunk` at a time; within a chunk every rank traces its contiguous
    share (shard_samples) with render_fn(base, count) -> tensor of radiance SUMS, the chunk is reduced to `reduce_dst`
    (all ranks if None) and added to the running total.  Yields (total, samples_done) after every chunk: on the
    destination rank `total` is the sum over all samples so far (divide by samples_done for the mean image); other
    ranks get their local partial and should only use the count.  One reduce of W*H*3 floats per chunk is the only
    traffic between ranks.  With chunk < world some ranks have an empty share of a chunk: render_fn(base, 0) must
    return zeros (see render_sharded).
    """
    import torch
    import torch.distributed as dist
    if chunk < 1:
        raise ValueError("chunk must be >= 1")
    total = None
    done = 0
    while done < spp:
        n = min(chunk, spp - done)
        base, count = shard_samples(n, rank, world, sample_base + done)
        part = render_fn(base, count)
        if world > 1:
            if reduce_dst is None:
                dist.all_reduce(part, op=dist.ReduceOp.SUM)
            else:
                dist.reduce(part, dst=reduce_dst, op=dist.ReduceOp.SUM)
        total = part.clone() if total is None else total.add_(part)
        done += n
        yield total, done
