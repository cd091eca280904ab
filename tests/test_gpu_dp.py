"""GPU: the data-parallel surface (llm_np_cp_b200.dp.DataParallelLM).  DP correctness = every shard equals the 1-GPU
result BIT FOR BIT: a sequence's logits do not depend on which rows share its batch.  With one
GPU the shards are run one after the other on the same device; with >= 2 GPUs a two-process NCCL run is added."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import ROOT
from oracle import np_oracle as O
import llm_np_cp_b200 as M
from llm_np_cp_b200 import dp


def _model(device=0):
    cfg = O.llama_config("tiny")
    return cfg, M.LlamaForCausalLM(M.AttributeDict(cfg), O.synth_weights(cfg, 1, 0.08, 0.1), device=device)


def test_shards_are_bit_identical_to_the_full_batch():
    cfg, model = _model()
    ids = np.random.default_rng(2).integers(0, cfg.vocab_size, (5, 300))
    full = dp.DataParallelLM(model, rank=0, world=1).prefill_last_logits(ids)
    for world in (2, 3):
        parts = [dp.DataParallelLM(model, rank=r, world=world).prefill_last_logits(ids, gather=False) for r in range(world)]
        assert torch.equal(torch.cat(parts, 0), full)


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import np_oracle as O
import llm_np_cp_b200 as M
from llm_np_cp_b200 import dp
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
cfg = O.llama_config("tiny")
model = M.LlamaForCausalLM(M.AttributeDict(cfg), O.synth_weights(cfg, 1, 0.08, 0.1), device=rank)
ids = np.random.default_rng(2).integers(0, cfg.vocab_size, (5, 300))
full = dp.DataParallelLM(model).prefill_last_logits(ids)
if rank == 0:
    np.save(sys.argv[2], full.cpu().numpy())
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_process_nccl_run_matches_one_gpu(tmp_path):
    cfg, model = _model()
    ids = np.random.default_rng(2).integers(0, cfg.vocab_size, (5, 300))
    want = dp.DataParallelLM(model, rank=0, world=1).prefill_last_logits(ids).cpu().numpy()
    script, out = tmp_path / "w.py", tmp_path / "full.npy"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script), ROOT, str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert np.array_equal(np.load(out), want)          # gathered rows == the single-GPU result, bit for bit
