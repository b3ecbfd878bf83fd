"""Multi-GPU plumbing: one process per GPU, frames sharded with no data-path collective, one
broadcast of the streak database at init (torch.distributed / NCCL).

The reference scales out with OS processes over disjoint frame ranges
(main_threaded.py:109-170, at most 10 children); frames are independent (per-frame reseed,
common/generator.py:318), so the same partition maps onto GPUs.
"""
from __future__ import annotations

import os

import numpy as np


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: str = "nccl"):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [start, stop) of ``n_items`` frames for ``rank`` (the reference's
    threaded launcher also hands out contiguous frame ranges, main_threaded.py:114-137)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class _DevicePtr:
    """Expose a raw device allocation through __cuda_array_interface__ so torch can wrap it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=3)


def broadcast_streak_db(ctx, textures=None, src: int = 0):
    """The single collective of the path.  Rank ``src`` uploads the normalised textures
    (``ctx.set_streak_db``); every other rank allocates the same layout and receives the bytes with
    one ``dist.broadcast`` straight into the library's device buffer (NCCL over NVLink)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == src:
        assert textures is not None
        ctx.set_streak_db(textures)
        meta = [ctx.db_heights.tolist(), int(ctx.db_width), ctx.db_ratios.tolist()]
    else:
        meta = [None, None, None]
    if world == 1:
        return
    dist.broadcast_object_list(meta, src=src)
    if rank != src:
        ctx.alloc_streak_db(np.array(meta[0], np.int32), meta[1])
        ctx.db_ratios = np.array(meta[2])
    ptr, nbytes = ctx.streak_db_device_ptr()
    t = torch.as_tensor(_DevicePtr(ptr, nbytes), device=torch.device("cuda", ctx.device))
    dist.broadcast(t, src=src)
    torch.cuda.synchronize()


def broadcast_streak_db_host(textures=None, src: int = 0):
    """Host variant (gloo / CPU tests): returns the texture list on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    if rank == src:
        heights = [int(t.shape[0]) for t in textures]
        width = int(textures[0].shape[1])
        flat = np.concatenate([np.ascontiguousarray(t, np.uint8).reshape(-1) for t in textures])
        meta = [heights, width]
    else:
        meta = [None, None]
    dist.broadcast_object_list(meta, src=src)
    heights, width = meta
    buf = torch.from_numpy(flat.copy()) if rank == src else torch.empty(sum(heights) * width, dtype=torch.uint8)
    dist.broadcast(buf, src=src)
    flat = buf.numpy()
    out, o = [], 0
    for h in heights:
        out.append(flat[o:o + h * width].reshape(h, width).copy())
        o += h * width
    return out
