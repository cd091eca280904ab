"""CPU: the N>1 plumbing (sharding, max-over-ranks timing, ragged gather) with world_size-2 gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llm_np_cp_b200 import dp


def test_shard_bounds_cover_exactly_once():
    for n in [0, 1, 7, 8, 64, 65]:
        for world in [1, 2, 3, 4, 8]:
            seen = []
            for r in range(world):
                b, e = dp.shard_bounds(n, r, world)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
            sizes = [dp.shard_bounds(n, r, world)[1] - dp.shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_bounds(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = np.arange(5 * 3).reshape(5, 3)                      # 5 sequences over 2 ranks -> 3 + 2
        mine = dp.shard_batch(ids, rank, world)
        tmax = dp.max_over_ranks([10.0 + rank, 5.0 - rank])
        local = torch.from_numpy(mine.astype(np.float32)) * 2.0   # stand-in for per-rank last-token logits
        full = dp.gather_rows(local, ids.shape[0], rank, world)
        q.put((rank, mine.tolist(), tmax, full.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.arange(15).reshape(5, 3)
    assert res[0][1] == ids[:3].tolist() and res[1][1] == ids[3:].tolist()      # disjoint, ordered shards
    for r in res:
        assert r[2] == [11.0, 5.0]                                               # max over ranks
        assert r[3] == (ids * 2.0).tolist()                                      # ragged gather restores order


class _FakeModel:
    """Stands in for an engine on CPU: 'logits' of a row = f(ids of that row) only, like the real path."""
    vocab_size = 7
    device = torch.device("cpu")

    def __call__(self, ids, **kw):
        ids = np.asarray(ids)
        lg = torch.from_numpy((ids.sum(axis=1, keepdims=True) + np.arange(self.vocab_size)[None]).astype(np.float32))
        return None, lg, None, None, None


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = np.arange(5 * 4).reshape(5, 4)
        dpm = dp.DataParallelLM(_FakeModel())
        full = dpm.prefill_last_logits(ids)
        local = dpm.prefill_last_logits(ids, gather=False)
        one_row = dpm.prefill_last_logits(ids[:1])                # fewer rows than ranks: rank 1 has an empty shard
        q.put((rank, dpm.local_rows(5), full.tolist(), local.shape[0], one_row.tolist()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_lm_two_ranks_gloo():
    """DataParallelLM (the multi-GPU product surface): global batch in, gathered last-token logits out, rows in global
    order and equal to the single-process result; ragged and empty shards included."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.arange(20).reshape(5, 4)
    want = _FakeModel()(ids)[1].tolist()
    assert dp.DataParallelLM(_FakeModel(), rank=0, world=1).prefill_last_logits(ids).tolist() == want
    assert res[0][1] == (0, 3) and res[1][1] == (3, 5)
    for r in res:
        assert r[2] == want and r[4] == want[:1]
    assert res[0][3] == 3 and res[1][3] == 2
