"""Multi-GPU plumbing: static task sharding (one process per GPU) and the single result gather of the path.

Planning tasks are independent (`RRTStar` instances never interact), so there is no data-path collective: each rank
plans its own shard and rank 0 receives the fixed-size result records with one `torch.distributed.gather` (NCCL on
GPUs, gloo in the CPU tests).  SURVEY.md §8e.
"""
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_tasks: int, rank: int, world: int, interleaved: bool = True) -> np.ndarray:
    """Task ids of `rank`.  Interleaved (t mod world == rank) spreads the 10 tasks of a map and the expensive tasks
    evenly; contiguous keeps a map's tasks on one rank (one encoder pass per map)."""
    if interleaved:
        return np.arange(rank, n_tasks, world, dtype=np.int64)
    per = (n_tasks + world - 1) // world
    return np.arange(rank * per, min(n_tasks, (rank + 1) * per), dtype=np.int64)


def gather_records(records: torch.Tensor, index: torch.Tensor, n_tasks: int, dst: int = 0) -> Optional[torch.Tensor]:
    """records [n_local, F] (float64) of the tasks `index` [n_local] -> [n_tasks, F] on rank `dst` (None elsewhere).
    Shards may be ragged: they are padded to the largest shard for the collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = torch.empty((n_tasks, records.shape[1]), dtype=records.dtype, device=records.device)
        out[index.to(records.device)] = records
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    n_max = (n_tasks + world - 1) // world
    pad = torch.full((n_max, records.shape[1] + 1), -1.0, dtype=records.dtype, device=records.device)
    pad[:records.shape[0], 0] = index.to(records.device, records.dtype)
    pad[:records.shape[0], 1:] = records
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_tasks, records.shape[1]), dtype=records.dtype, device=records.device)
    for b in bufs:
        keep = b[:, 0] >= 0
        out[b[keep, 0].long()] = b[keep, 1:]
    return out
