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


class DataParallelLM:
    """The multi-GPU product surface for the path that shards: one process per GPU (torchrun /
    torch.distributed already initialised, or a single process), every rank holds a full replica of the weights and runs
    its own engine on its contiguous slice of the batch; NO collective touches the data path.  The only communication is
    the optional gather of the last-token logits, which is what generate()-style callers consume
    (llama3.2_model_numpy.py:899).

        dpm = DataParallelLM(model)                  # model: this rank's LlamaForCausalLM / Gemma2ForCausalLM
        logits = dpm.prefill_last_logits(ids)        # ids: the GLOBAL [B,S] batch (same array on every rank)
                                                     # -> [B,V] on every rank, rows in the global order
    Every row equals what a single GPU computes for that sequence bit for bit (a sequence's result does not depend on its
    batch neighbours; asserted by tests/test_gpu_model.py and tests/test_gpu_dp.py).
    """

    def __init__(self, model, rank: int = None, world: int = None):
        import torch.distributed as dist
        on = dist.is_available() and dist.is_initialized()
        self.model = model
        self.rank = rank if rank is not None else (dist.get_rank() if on else 0)
        self.world = world if world is not None else (dist.get_world_size() if on else 1)

    def local_rows(self, n_rows: int) -> Tuple[int, int]:
        return shard_bounds(n_rows, self.rank, self.world)

    def prefill_local(self, ids: np.ndarray, **kw):
        """This rank's slice through model(...): returns the reference 5-tuple for the local rows (may be empty)."""
        mine = shard_batch(np.asarray(ids), self.rank, self.world)
        if mine.shape[0] == 0:
            return None
        return self.model(mine, **kw)

    def prefill_last_logits(self, ids: np.ndarray, gather: bool = True):
        """Last-position logits [B,V] of the global batch (gathered to every rank), or this rank's rows if not gather."""
        import torch
        from .model import LOGITS_LAST
        ids = np.asarray(ids)
        mine = shard_batch(ids, self.rank, self.world)
        if mine.shape[0] > 0:
            _, local, _, _, _ = self.model(mine, use_cache=False, output_hidden_states=False, output_attentions=False,
                                           logits_mode=LOGITS_LAST)
        else:
            local = torch.empty((0, self.model.vocab_size), dtype=torch.float32, device=self.model.device)
        if not gather:
            return local
        return gather_rows(local, ids.shape[0], self.rank, self.world)
