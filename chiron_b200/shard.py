"""Read sharding across the GPUs of one box.

The unit of work is a read (assembly needs all windows of a read in order); windows never depend on other windows, so
there is no collective on the decode path.  The only optional exchange is one broadcast of the packed weight blob."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple


def rank_world() -> Tuple[int, int]:
    """(rank, world_size) from the torchrun environment; (0, 1) when not launched distributed."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def assign_reads(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time partition: reads sorted by size (largest first) go to the least loaded rank.
    Deterministic (ties by index), returns per-rank lists of read indices in ascending order."""
    loads = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(sizes)), key=lambda j: (-int(sizes[j]), j)):
        r = min(range(world), key=lambda q: (loads[q], q))
        parts[r].append(i)
        loads[r] += int(sizes[i])
    return [sorted(p) for p in parts]


def broadcast_blob(blob: bytes = None, src: int = 0, device=None) -> bytes:
    """One broadcast of the weight blob from rank `src` over the initialised torch.distributed group (NCCL on the GPU
    box, gloo in CPU tests).  Ranks other than `src` pass blob=None."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        if blob is None:
            raise ValueError("no process group and no blob")
        return blob
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([len(blob) if dist.get_rank() == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if dist.get_rank() == src:
        buf = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    else:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def gather_to_rank0(obj, timeout_s: float = 1800.0):
    """Every rank's ``obj`` as a list on rank 0 (None elsewhere); [obj] when not launched distributed.  The one
    control-plane exchange of a sharded `chiron call`: the per-rank timings that rank 0 merges into meta/all.meta
    (SURVEY.md section 8e).  CPU tensors over a gloo group built from the torchrun environment (MASTER_ADDR/MASTER_PORT);
    an already initialised process group is used as it is and left alone."""
    rank, world = rank_world()
    if world == 1:
        return [obj]
    import datetime
    import torch.distributed as dist
    own = not dist.is_initialized()
    if own:
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=timeout_s))
    try:
        group = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
        out = [None] * world if rank == 0 else None
        dist.gather_object(obj, out, dst=0, group=group)
        return out
    finally:
        if own:
            dist.destroy_process_group()
