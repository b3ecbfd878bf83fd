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


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(local_rank: int):
    """Pin the calling thread (and the threads it creates afterwards) to the CPUs of the NUMA node the GPU hangs off,
    BEFORE page-locked staging buffers are allocated: pages are placed on the node of the thread that first touches
    them, and a staging buffer on the other socket makes every host<->device copy cross the inter-socket link.
    The reference's launcher leaves placement to the OS (main_threaded.py:176); with one process per GPU it decides
    how far end-to-end throughput scales past one GPU.  Returns a dict describing what was done (for the bench line);
    never raises -- on hosts without NUMA information nothing changes.  RAIN_B200_NUMA_BIND=0 disables it."""
    info = {"bound": False}
    if os.environ.get("RAIN_B200_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return info
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            parts = vis.split(",")
            if local_rank < len(parts) and parts[local_rank].strip().isdigit():
                idx = int(parts[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read())
        cpus = _parse_cpulist(open(base + "/local_cpulist").read())
        info.update(node=node, gpu_bus=bus)
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if node < 0 or not want or want == allowed:
            info["reason"] = "no NUMA locality to exploit (node %d, %d local of %d allowed CPUs)" % (node, len(want), len(allowed))
            return info
        os.sched_setaffinity(0, want)
        info.update(bound=True, cpus=len(want), allowed=len(allowed))
    except Exception as e:          # no NVML / no sysfs: leave placement to the OS
        info["reason"] = "%s: %s" % (type(e).__name__, e)
    return info


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
