#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: prefill tokens/s, Llama-3.2-1B shapes, bf16, seq=2048.

  python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU module (oracle/_ref build) on host cores
  python bench.py --workload decode [--prompt P]           # config #5 as the headline line
  torchrun --nproc-per-node N bench.py --gpus N --tp N ... # optional tensor parallelism: one batch over N GPUs

A "step" is one forward of B sequences x 2048 synthetic tokens per GPU producing the full [B,S,V] fp32
logits (what the reference's model(ids) returns).  Batches shard across GPUs with no collective (weak
scaling).  One JSON line on stdout (rank 0); besides the contract keys it carries "decode" (a config-#5 rollout:
BASELINE's metric ends "...; decode tok/s"), "precise" (the cost of the parity mode, labelled), "roofline",
"cpu_baseline" and "clocks".
"""
import os

os.environ.setdefault("OPENBLAS_NUM_THREADS", str(os.cpu_count() or 1))   # before NumPy loads its BLAS
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))

import argparse
import json
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.lines, self.p = gpu, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.p.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_reference_sample(cfg, sample_tokens, steps, warmup):
    """Times the reference's CPU path on a bounded sample (one sequence of `sample_tokens` tokens per step, default
    arguments of model(ids): attentions and hidden states retained, llama3.2_model_numpy.py:785-786).
    kind "reference": the reference's OWN LlamaForCausalLM_np, loaded from /root/reference when mounted or from the
    byte-compiled build oracle/_ref (oracle/build_ref.py) on the GPU box, weights injected through its `weights` seam.
    kind "port": the oracle port issuing the reference's own N-D np.dot calls (faithful=True) when neither exists (or the
    model is Gemma-2, whose reference file is CuPy-only)."""
    from oracle import np_oracle as O
    from oracle import ref_harness as R
    w = O.synth_weights(cfg, seed=0, std=0.02)
    ids = np.random.default_rng(0).integers(0, cfg.vocab_size, (1, sample_tokens))
    try:        # every host core for the BLAS calls, also when a launcher (torchrun) exported OMP_NUM_THREADS=1
        from threadpoolctl import threadpool_limits
        limit = threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        limit = None
    kind = "port"
    if cfg.family == "llama" and R.available():
        try:
            mod = R.load_llama_numpy()
            model, _ = R.build_reference_model(mod, cfg, w)
            step = lambda: model(ids, use_cache=False)           # noqa: E731  the reference's own __call__
            kind = "reference"
        except Exception as e:                                    # a broken build of oracle/_ref must not kill the bench
            sys.stderr.write(f"[bench] reference module unavailable ({e!r}); timing the NumPy port instead\n")
    if kind == "port":
        W = O.to_numpy_f32(w)
        step = lambda: O.forward(cfg, W, ids, faithful=True, keep_attn=True)   # noqa: E731
    for _ in range(warmup):
        step()
    t = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t) / max(steps, 1)
    if limit is not None:
        limit.restore_original_limits()
    return sample_tokens / dt, dt, kind


def measure_decode(cfg, model, dev, dist, rank, world, local, pk, B, P, NT, steps, warmup, sample_clocks=True, tp=1):
    """BASELINE.json config #5: autoregressive decode with the KV cache (the loop of generate(),
    llama3.2_model_numpy.py:881-909): B sequences per GPU (replicas only across GPUs), greedy feedback on the device, every
    one-token step replayed as one CUDA graph.  One rollout = `NT` one-token forwards for the batch; the prompt prefill is
    outside the timed region.  Returns the dict that goes into the JSON line (rank 0) or None."""
    import torch
    import llm_np_cp_b200 as M
    from llm_np_cp_b200.model import LOGITS_LAST
    from llm_np_cp_b200 import dp
    V = cfg.vocab_size
    rng = np.random.default_rng(99 + (rank if tp <= 1 else 0))     # tensor parallel: every rank feeds the same batch
    prompt = torch.from_numpy(rng.integers(0, V, (B, P)).astype(np.int64)).to(dev)
    cache = M.KVCache(max_batch=B, max_tokens=P + NT)
    replicas = world if tp <= 1 else 1
    model.set_decode_graph(not os.environ.get("B2_NO_GRAPH"))
    tok = torch.empty((B, 1), dtype=torch.int64, device=dev)
    tok_host = torch.empty((B, 1), dtype=torch.int64).pin_memory()
    lg = torch.empty((B, V), dtype=torch.float32, device=dev)

    def rollout(host_loop):
        cache.reset()
        _, l0, _, _, _ = model(prompt, use_cache=True, kv_cache=cache, output_hidden_states=False, logits_mode=LOGITS_LAST)
        model.argmax(l0, out=tok.view(-1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(NT):
            if host_loop:      # e2e: the sampled ids go to the host and come back every token, like generate()
                tok_host.copy_(tok, non_blocking=False)
                tok.copy_(tok_host, non_blocking=True)
            model(tok, use_cache=True, kv_cache=cache, output_hidden_states=False, logits_mode=LOGITS_LAST, logits_out=lg)
            model.argmax(lg, out=tok.view(-1))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        rollout(False)
    barrier()
    model.reset_stats()
    sampler = ClockSampler(local) if sample_clocks else None
    if rank == 0 and sampler:
        sampler.start()
    ms = 0.0
    for _ in range(steps):
        ms += rollout(False)[0]
    barrier()
    clocks = sampler.stop() if (rank == 0 and sampler) else None
    stt = model.stats()
    launches = stt["launches"]
    e2e_ms = 0.0
    for _ in range(steps):
        e2e_ms += rollout(True)[1]
    model.set_decode_graph(False)
    ms, e2e_ms = dp.max_over_ranks([ms, e2e_ms], device=dev)
    toks = B * NT * replicas * steps
    if rank != 0:
        return None
    from llm_np_cp_b200 import shapes
    L = cfg.num_hidden_layers
    kv = cfg.num_key_value_heads * cfg.head_dim
    t_avg = P + NT / 2.0
    bytes_step = 2.0 * shapes.matmul_params(cfg) + B * t_avg * L * 2 * kv * 2   # bf16 weights once + KV read
    if tp > 1:                                                                   # per GPU: 1/tp of everything but the embedding gather
        bytes_step /= tp
    step_s = ms / 1e3 / (steps * NT)
    gbs = bytes_step / step_s / 1e9
    out_tp = None
    if tp > 1:
        H = cfg.hidden_size
        Lc = cfg.num_hidden_layers      # the graph-replayed steps do not pass through the host-side counters: static count
        out_tp = {"world": tp, "nccl_calls_per_token_step": 2 * Lc + 1,
                  "payload_bytes_per_token_step": 2 * Lc * B * H * 4 + B * cfg.vocab_size * 4,
                  "algorithmic_nvlink_bytes_per_allreduce": 2.0 * B * H * 4 * (tp - 1) / tp,
                  "note": "2 all-reduces of [B,H] fp32 partial sums per layer (after o_proj and down_proj) + 1 all-gather of logits"}
    return {
        "metric": "decode_tokens_per_sec", "value": toks / (ms / 1e3), "unit": "tokens/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak" if tp <= 1 else "strong",
        "tp": out_tp,
        "config": {"batch_per_gpu": B if tp <= 1 else None, "batch": B * replicas, "prompt": P, "new_tokens": NT,
                   "parallelism": (f"replicas x{world} (no collective)" if tp <= 1 else f"tp{tp} (heads / FFN columns / vocabulary split, NCCL all-reduce over NVLink)"),
                   "sampling": "greedy argmax on device (b2_argmax)",
                   "cuda_graph": not os.environ.get("B2_NO_GRAPH"),
                   "l2": "all weights (2.47 GB for 1B) streamed every token step >> 126 MB L2; no flush needed"},
        "ms_per_token_step": step_s * 1e3,
        "roofline": {"bound": "hbm", "kernel": "whole token step (weight-streaming GEMMs at M = batch + split-KV attention)",
                     "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": None,
                     "bytes_per_token_step": bytes_step, "peak_source": pk["src"] + " hbm_gbs"},
        "e2e": {"value": toks / (e2e_ms / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": B * 8 * NT * world,
                "d2h_bytes_per_step": B * 8 * NT * world,
                "what": "per token: sampled ids D2H, ids H2D, one-token forward through model(), greedy pick"},
        "gpu_launches": int(launches), "clocks": clocks}


def bench_decode(a, cfg, model, dev, dist, rank, world, local, pk, workload):
    """`--workload decode`: the decode rollout as the headline line."""
    line = measure_decode(cfg, model, dev, dist, rank, world, local, pk, a.batch, a.prompt, a.new_tokens, a.steps, a.warmup, tp=a.tp)
    if rank == 0:
        line["config"] = dict({"workload": workload}, **line["config"])
        line.update({"vs_baseline": None, "dtype": "bf16", "data": "synthetic"})
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def prefill_config(cfg, fam, workload, B, S, world):
    """The `config` object of the prefill line: identical in the engine's arm and in `--impl reference`."""
    return {"workload": workload, "model": f"{fam} shapes, random-init bf16 weights (std 0.02)",
            "global_batch": B * world, "seq_len": S, "tokens_per_step_per_gpu": B * S,
            "parallelism": f"dp{world} (batch-sharded, weights replicated, no collective)",
            "logits": "full [B,S,V] fp32 written to HBM every step",
            "l2": "per-step working set (2.5 GB weights + 8.4 GB logits + activations) >> 126 MB L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="prefill", choices=["prefill", "decode"])
    ap.add_argument("--model", default="1b", choices=["1b", "3b", "tiny", "gemma2-2b"])
    ap.add_argument("--hf-flags", action="store_true", help="gemma2: attention soft-cap + sliding window (HF semantics)")
    ap.add_argument("--prompt", type=int, default=16, help="decode: prompt length")
    ap.add_argument("--new-tokens", type=int, default=128, help="decode: generated tokens per sequence")
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU per step (default 8 prefill / 64 decode)")
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--tp", type=int, default=0, help="tensor-parallel degree (= --gpus): one batch split by heads / FFN columns over the GPUs")
    ap.add_argument("--cpu-sample-tokens", type=int, default=int(os.environ.get("B2_CPU_SAMPLE_TOKENS", 16)))
    ap.add_argument("--cpu-long-sample", type=int, default=int(os.environ.get("B2_CPU_LONG_SAMPLE", 128)),
                    help="--impl reference: one extra forward at this many tokens (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="prefill workload: skip the appended config-#5 decode rollout")
    ap.add_argument("--no-precise", action="store_true", help="prefill workload: skip the parity-mode (hi/lo split) measurement")
    ap.add_argument("--event-every", type=int, default=4, help="per-launch CUDA events on every k-th timed step only")
    a = ap.parse_args()

    from llm_np_cp_b200 import shapes          # presets / FLOP arithmetic live in the product package; oracle/ is only
    cfg = shapes.gemma2_config("2b") if a.model == "gemma2-2b" else shapes.llama_config(a.model)   # imported by the CPU legs
    if a.batch <= 0:
        a.batch = 8 if a.workload == "prefill" else 64
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    fam = "gemma-2-2b" if a.model == "gemma2-2b" else f"llama-3.2-{a.model}"
    workload = (f"{fam} prefill seq={a.seq}" + (" hf-flags" if a.hf_flags else "") if a.workload == "prefill" else
                f"{fam} decode batch={a.batch} prompt={a.prompt} new={a.new_tokens}")
    pk = peaks()
    f_tok = shapes.flops_per_token(cfg, a.seq)

    if a.impl == "reference":
        # the reference's own CPU implementation of the path, on this box's host cores; rank 0 only.  Each step is a
        # bounded sample of the workload: ONE sequence of --cpu-sample-tokens tokens (default 16 = BASELINE config #1,
        # ~10 s on 128 cores) so that --steps 20 --warmup 5 ends within a few minutes; one extra forward at a longer sample
        # (--cpu-long-sample, default 128 tokens) is reported beside it.
        if rank != 0:
            return
        tps, dt, kind = cpu_reference_sample(cfg, a.cpu_sample_tokens, a.steps, a.warmup)
        cores = os.cpu_count()
        sample = (f"1 sequence x {a.cpu_sample_tokens} tokens per step through "
                  f"{'the reference module (LlamaForCausalLM_np, oracle/_ref build)' if kind == 'reference' else 'the NumPy port (reference N-D np.dot calls)'}"
                  f", NumPy {np.__version__}, {cores} BLAS threads")
        line = {
            "impl": "reference", "metric": "prefill_tokens_per_sec", "value": tps, "unit": "tokens/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32/f64 (NumPy)", "data": "synthetic",
            "config": prefill_config(cfg, fam, workload, a.batch, a.seq, a.gpus),
            "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": kind, "sample": sample,
                             "sample_tokens": a.cpu_sample_tokens},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
        if a.cpu_long_sample > a.cpu_sample_tokens:
            tps2, dt2, kind2 = cpu_reference_sample(cfg, a.cpu_long_sample, 1, 0)
            line["cpu_baseline_long"] = {"value": tps2, "unit": "tokens/s", "cores": cores, "kind": kind2, "seconds": dt2,
                                         "sample": f"1 sequence x {a.cpu_long_sample} tokens, 1 forward", "sample_tokens": a.cpu_long_sample}
        print(json.dumps(line))
        return

    import torch
    import llm_np_cp_b200 as M
    from llm_np_cp_b200.model import LOGITS_ALL, LOGITS_ALL_LAST_TO_HOST
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B, S, V = a.batch, a.seq, cfg.vocab_size
    w = shapes.synth_checkpoint(cfg, seed=0, device=dev)
    tp = a.tp if a.tp > 1 else 1
    if tp > 1 and tp != world:
        raise SystemExit(f"--tp {tp} needs exactly {tp} ranks (torchrun --nproc-per-node {tp}), got {world}")
    tpkw = dict(tensor_parallel=True) if tp > 1 else {}
    if cfg.family == "gemma2":
        model = M.Gemma2ForCausalLM(M.AttributeDict(cfg), w, device=local, max_positions=max(S, a.prompt + a.new_tokens, 16), hf_flags=a.hf_flags, **tpkw)
    else:
        model = M.LlamaForCausalLM(M.AttributeDict(cfg), w, device=local, max_positions=max(S, a.prompt + a.new_tokens, 16), **tpkw)
    if a.workload == "decode":
        del w
        return bench_decode(a, cfg, model, dev, dist, rank, world, local, pk, workload)
    rng = np.random.default_rng(1234 + (rank if tp <= 1 else 0))    # tensor parallel: every rank feeds the same batch
    ids_host = torch.from_numpy(rng.integers(0, V, (B, S)).astype(np.int64)).pin_memory()
    ids_dev = ids_host.to(dev)
    logits = None

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        nonlocal logits
        _, logits, _, _, _ = model(ids_dev, use_cache=False, output_hidden_states=False, output_attentions=False,
                                   logits_mode=LOGITS_ALL)

    # ---- value: inputs resident in HBM ----
    # Per-launch CUDA events (the roofline figures) are recorded on every --event-every-th step only: events between
    # launches defeat the programmatic-dependent-launch overlap, so most timed steps run exactly what e2e runs.
    for _ in range(a.warmup):
        step_resident()
    logits = None
    barrier()
    model.reset_stats()
    model.set_timing(max(1, a.event_every))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_resident()            # one extra untimed step (every rank: a tensor-parallel step contains collectives) so that
                               # nvidia-smi is sampling under load before t0
    model.set_timing(max(1, a.event_every))   # restart the every-k-th counter; first timed step carries events
    model.reset_stats()        # only the timed steps below count towards the roofline figures
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step_resident()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    st = model.stats()
    model.set_timing(False)
    fast_last = logits[:, -1, :].clone() if (world == 1 and not a.no_precise) else None
    logits = None
    torch.cuda.empty_cache()

    # ---- e2e: the C-ABI call with HOST buffers (ids H2D, forward with all logits, last-position logits D2H) ----
    out_host = torch.empty((B, V), dtype=torch.float32).pin_memory().numpy()   # pinned: the D2H of the result is a DMA
    ids_np = ids_host.numpy()
    for _ in range(max(1, min(a.warmup, 2))):
        model.forward_host(ids_np, logits_mode=LOGITS_ALL_LAST_TO_HOST, out=out_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        model.forward_host(ids_np, logits_mode=LOGITS_ALL_LAST_TO_HOST, out=out_host)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3

    tm = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tm[0]), float(tm[1])
    tokens = B * S * (world if tp <= 1 else 1) * a.steps
    value = tokens / (ms / 1e3)
    e2e = tokens / (e2e_ms / 1e3)

    # ---- decode: BASELINE's metric ends "...; decode tok/s" -- a short config-#5 rollout rides along in the same line ----
    decode = None
    if not a.no_decode and cfg.family == "llama" and tp <= 1:
        decode = measure_decode(cfg, model, dev, dist, rank, world, local, pk, 64, 16, 128, steps=3, warmup=1, sample_clocks=False)

    # ---- parity mode: what BASELINE's literal "1e-2 max-abs" costs.  Same kernels, every bf16 activation carried as hi + lo
    # ---- planes (logits within ~2e-4 of the reference, tests/test_gpu_model.py).  Labelled, never the headline.
    precise = None
    if fast_last is not None and cfg.head_dim in (64, 128, 256):
        try:
            cls = M.Gemma2ForCausalLM if cfg.family == "gemma2" else M.LlamaForCausalLM
            kw = dict(hf_flags=a.hf_flags) if cfg.family == "gemma2" else {}
            pm = cls(M.AttributeDict(cfg), w, device=local, max_positions=max(S, 16), precise=True, **kw)
            for _ in range(2):
                _, plog, _, _, _ = pm(ids_dev, use_cache=False, output_hidden_states=False, logits_mode=LOGITS_ALL)
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            psteps = 3
            p0.record()
            for _ in range(psteps):
                _, plog, _, _, _ = pm(ids_dev, use_cache=False, output_hidden_states=False, logits_mode=LOGITS_ALL)
            p1.record()
            torch.cuda.synchronize()
            pms = p0.elapsed_time(p1) / psteps
            diff = (plog[:, -1, :] - fast_last).abs()
            precise = {"parity_mode": "hi/lo split-bf16 activations (b2_config.precise): logits within 1e-2 max-abs of the reference",
                       "value": B * S / (pms / 1e3), "unit": "tokens/s", "ms_per_step": pms, "steps": psteps,
                       "slowdown_vs_fast": pms / (ms / a.steps),
                       "fast_vs_precise_last_logits": {"max_abs": float(diff.max()), "mean_abs": float(diff.mean()),
                                                       "logit_std": float(plog[:, -1, :].std())},
                       "note": "not the headline: never quote throughput from this mode without the label"}
            del pm, plog
        except Exception as ex:      # e.g. a head_dim the precise attention does not cover
            precise = {"parity_mode": "unavailable", "why": str(ex)[:200]}
        torch.cuda.empty_cache()
    del w

    if rank == 0:
        nt = max(1, int(st["timed_forwards"]))      # steps whose launches carried events
        gemm_tf = st["gemm_flops"] / (st["gemm_ms"] * 1e-3) / 1e12 if st["gemm_ms"] > 0 else None
        attn_tf = st["attn_flops"] / (st["attn_ms"] * 1e-3) / 1e12 if st["attn_ms"] > 0 else None
        peak = pk["tf_sustained"]
        traffic, tsrc = None, None
        for tname in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if a.model == "1b" and S == 2048 and B == 8 and os.path.exists(tpath):   # the ncu capture is for exactly this workload
                with open(tpath) as f:
                    traffic, tsrc = json.load(f)["avg_dram_bytes_per_gemm_launch"], tname
                break
        step_ms = ms / a.steps
        line = {
            "metric": "prefill_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": prefill_config(cfg, fam, workload, B, S, world),
            "model_flops_per_token": f_tok,
            "model_tflops_per_gpu": value / world * f_tok / 1e12,
            "model_frac_of_peak": value / world * f_tok / 1e12 / peak,
            "roofline": {"bound": "tensor", "kernel": "gemm2_bf16_kernel (tcgen05 cta_group::2, all fused-epilogue variants)",
                         "achieved": gemm_tf, "peak": peak, "unit": "TFLOP/s",
                         "frac": (gemm_tf / peak) if gemm_tf else None, "traffic": traffic,
                         "traffic_note": f"avg DRAM bytes per GEMM launch from profiles/{tsrc} (ncu --set full); algorithmic 507 MB",
                         "peak_source": pk["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                         "events_on_steps": f"{nt} of {a.steps} timed steps (every {max(1, a.event_every)}th)",
                         "gemm_ms_per_step": st["gemm_ms"] / nt, "gemm_share_of_step": st["gemm_ms"] / nt / step_ms,
                         "attention": {"achieved": attn_tf, "ms_per_step": st["attn_ms"] / nt,
                                       "share_of_step": st["attn_ms"] / nt / step_ms}},
            "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": B * S * 8 * world, "d2h_bytes_per_step": B * V * 4 * world,
                    "ms_per_step": e2e_ms / a.steps,
                    "what": "b2_forward_host: ids from host, all-position logits kept on device (like the reference's CuPy twin), "
                            "last-position logits [B,V] fp32 copied to host"},
            "gpu_launches": int(st["launches"]),
            "clocks": clocks,
        }
        if tp > 1:
            H = cfg.hidden_size
            line["scaling"] = "strong"
            line["config"]["parallelism"] = f"tp{tp} (heads / FFN columns / vocabulary split over {tp} GPUs, NCCL all-reduce over NVLink)"
            line["config"]["global_batch"] = B
            line["model_tflops_per_gpu"] = value * f_tok / 1e12 / tp
            line["model_frac_of_peak"] = value * f_tok / 1e12 / tp / peak
            line["tp"] = {"world": tp, "nccl_calls_per_step": st["comm_calls"] / a.steps, "payload_bytes_per_step": st["comm_bytes"] / a.steps,
                          "comm_ms_per_step": st["comm_ms"] / nt, "comm_share_of_step": st["comm_ms"] / nt / step_ms,
                          "algorithmic_nvlink_bytes_per_allreduce": 2.0 * B * S * H * 4 * (tp - 1) / tp,
                          "allreduce_payload_bytes": B * S * H * 4,
                          "note": "2 all-reduces of the [M,H] fp32 partial sums per layer (after o_proj and down_proj) + the all-gather of the logits"}
        if decode is not None:
            line["decode"] = decode
        if precise is not None:
            line["precise"] = precise
        if world == 1 and not a.no_cpu_baseline:
            tps, dt, kind = cpu_reference_sample(cfg, a.cpu_sample_tokens, 1, 0)
            line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": os.cpu_count(), "kind": kind,
                                    "sample": f"1 sequence x {a.cpu_sample_tokens} tokens, 1 timed forward of "
                                              + ("the reference module itself (LlamaForCausalLM_np from the oracle/_ref build)" if kind == "reference"
                                                 else "the NumPy port issuing the reference's own N-D np.dot calls")
                                              + f" (NumPy {np.__version__})", "seconds": dt}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
