"""Cell sharding across the GPUs of one box.

Per-cell chains are independent, so the data path needs no collective: every rank owns a contiguous
block of cells (columns) with all genes.  The only exchanges are the two small ones the driver
needs per recursion level — integer pooled counts (all-reduce, exact and invariant to the GPU
count) and per-region posteriors (all-gather).  One process per GPU, `torch.distributed` with the
NCCL backend on the box and gloo in the CPU tests; device work stays in libhb_b200.so.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_cells: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the contiguous block of cells owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_cells, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_cells: int, world: int) -> list[int]:
    return [shard_range(n_cells, r, world)[1] - shard_range(n_cells, r, world)[0] for r in range(world)]


def allreduce_pooled_counts(mafl_local, sizel_local):
    """Sum the per-rank partial counts of R/HoneyBADGER.R:1404-1405 (`rowSums(. > 0)` over the
    rank's cells): int32 all-reduce — exact, so pooled allele chains are identical at any GPU count."""
    import torch
    import torch.distributed as dist
    t = torch.stack([torch.as_tensor(mafl_local), torch.as_tensor(sizel_local)]).to(torch.int32)
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t[0], t[1]


def allgather_region_posteriors(local, n_cells: int):
    """local: [regions, cells_of_this_rank] -> [regions, n_cells] on every rank (cells in global order).

    Shards may differ by one cell, so blocks are padded to the largest shard for the fixed-size
    all-gather and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    local = torch.as_tensor(local)
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local
    world = dist.get_world_size()
    sizes = shard_sizes(n_cells, world)
    mx = max(sizes)
    pad = torch.zeros((local.shape[0], mx), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:, :s] for o, s in zip(out, sizes)], dim=1)


def allgather_states(y_local, n_cells: int):
    """uint8 Viterbi states [positions, cells_of_this_rank] -> [positions, n_cells] (debug / tests;
    production keeps per-cell outputs sharded)."""
    import torch
    t = allgather_region_posteriors(torch.as_tensor(y_local).t().contiguous().t(), n_cells)
    return t


def merge_votes(vote_local):
    """Votes are integer counts per position: all-reduce(sum)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(vote_local)).to(torch.int64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
