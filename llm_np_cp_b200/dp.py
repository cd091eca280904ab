"""Data-parallel plumbing for prefill / decode batches.

The hot path shards embarrassingly: sequences never interact (the batch axis is untouched by every op,
llama3.2_model_numpy.py:418-429), so each rank runs its own engine on its own slice with replicated
weights and NO data-path collective.  torch.distributed is used only for the barrier around a timed
region, the max-over-ranks time, and (optionally) gathering last-token logits to rank 0.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of n_items for `rank` (first n_items % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(ids: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Rows of a [B,S] id matrix owned by `rank`."""
    b, e = shard_bounds(ids.shape[0], rank, world)
    return ids[b:e]


def max_over_ranks(values: List[float], device=None) -> List[float]:
    """Element-wise MAX of per-rank timings (every multi-GPU number is the max over ranks)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def gather_rows(local, world_rows: int, rank: int, world: int):
    """Gather per-rank [b_r, V] row blocks (e.g. last-token logits) into [world_rows, V] on every rank.
    Blocks may be ragged (shard_bounds); shorter ones are padded for the collective and trimmed after."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return local
    sizes = [shard_bounds(world_rows, r, world) for r in range(world)]
    mx = max(e - b for b, e in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: e - b] for o, (b, e) in zip(out, sizes)], dim=0)
