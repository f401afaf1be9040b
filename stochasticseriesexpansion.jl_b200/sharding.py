"""Multi-GPU plumbing: walkers shard across ranks (one process per GPU); the only collective is the
reduction of binned observables.  `torch.distributed` is plumbing here (NCCL over
NVLink on GPUs, gloo in the CPU tests); nothing is exchanged inside a sweep."""
from __future__ import annotations

import numpy as np


def shard_walkers(n_total: int, rank: int, world: int):
    """Contiguous block partition -> (global id of the first local walker, local count)."""
    base, rem = divmod(n_total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def reduce_bins(sums, counts, group_index=None, n_groups: int = 1, device=None):
    """All-reduce one bin: per-walker accumulator sums [W, n_obs] and counts [W, 2] (torch tensors on the
    rank's device, or numpy arrays) -> (group sums [n_groups, n_obs], group counts [n_groups, 2]) summed
    over all walkers of all ranks.  `group_index[w]` maps a local walker to its temperature group."""
    import torch
    import torch.distributed as dist

    s = torch.as_tensor(sums, dtype=torch.float64, device=device)
    c = torch.as_tensor(counts, device=device).to(torch.float64)
    W = s.shape[0]
    if group_index is None:
        group_index = np.zeros(W, dtype=np.int64)
    gi = torch.as_tensor(np.asarray(group_index), dtype=torch.int64, device=s.device)
    buf = torch.zeros((n_groups, s.shape[1] + c.shape[1]), dtype=torch.float64, device=s.device)
    buf.index_add_(0, gi, torch.cat([s, c], dim=1))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf)
    return buf[:, : s.shape[1]], buf[:, s.shape[1]:]
