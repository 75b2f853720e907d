"""Multi-GPU sharding of the batch axis (SURVEY.md §8 e): contiguous index ranges per rank,
element table replicated, no exchange step on the data path.  The optional final gather of
result slabs to rank 0 is the only collective (NCCL over NVLink on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous slice; same formula as the C library uses per device."""
    return n_units * rank // world, n_units * (rank + 1) // world


def sharded_sweep(solve_slice: Callable[[int, int], np.ndarray], n_units: int, gather: bool = True,
                  group=None) -> Optional[np.ndarray]:
    """Each rank solves its slice [lo, hi) of the batch axis (axis 0 of the returned array).
    With gather=True rank 0 returns the concatenated result (others return None); otherwise
    every rank returns its own slab."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return solve_slice(0, n_units)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n_units, rank, world)
    local = np.ascontiguousarray(solve_slice(lo, hi))
    if not gather:
        return local
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    is_complex = np.iscomplexobj(local)
    flat = torch.from_numpy(local.view(np.float64) if is_complex else local).to(dev)
    sizes = [shard_range(n_units, r, world) for r in range(world)]
    # Only rank 0 needs the slabs: a gather (point-to-point sends on NCCL) moves each slab once, where an all-gather
    # would deliver every slab to every rank.  Slabs differ in length by at most one unit, so they are padded to the
    # longest one for the collective and trimmed afterwards.
    longest = max(h - l for l, h in sizes)
    tail = tuple(flat.shape[1:])
    send = flat
    if flat.shape[0] < longest:
        send = torch.zeros((longest,) + tail, dtype=flat.dtype, device=dev)
        send[: flat.shape[0]] = flat
    dst = dist.get_global_rank(group, 0) if group is not None else 0
    bufs = [torch.empty((longest,) + tail, dtype=flat.dtype, device=dev) for _ in sizes] if rank == 0 else None
    dist.gather(send.contiguous(), bufs, dst=dst, group=group)
    if rank != 0:
        return None
    out = torch.cat([b[: h - l] for b, (l, h) in zip(bufs, sizes)], dim=0).cpu().numpy()
    return out.view(np.complex128) if is_complex else out


