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
    row = int(np.prod(flat.shape[1:])) if flat.dim() > 1 else 1
    bufs = [torch.empty((h - l,) + tuple(flat.shape[1:]), dtype=flat.dtype, device=dev) for l, h in sizes]
    dist.all_gather(bufs, flat, group=group) if all(b.shape == bufs[0].shape for b in bufs) else \
        _uneven_all_gather(bufs, flat, group)
    if rank != 0:
        return None
    out = torch.cat(bufs, dim=0).cpu().numpy()
    del row
    return out.view(np.complex128) if is_complex else out


def _uneven_all_gather(bufs, flat, group):
    import torch.distributed as dist
    for r, b in enumerate(bufs):
        if dist.get_rank(group) == r:
            b.copy_(flat)
        dist.broadcast(b, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
