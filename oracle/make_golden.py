"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF.

Run here (CPU container, /root/reference mounted):   python -m oracle.make_golden [--skip-1b]

The reference ships no golden vectors, so parity is pinned on outputs of
the unmodified reference modules (loaded by oracle/ref_harness.py) for seeded synthetic
checkpoints.  Each case stores the inputs' recipe (config name, seeds) and the reference's
outputs; tests/test_oracle_golden.py replays the recipe through oracle/np_oracle.py and the
GPU parity tests replay it through the CUDA engine.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import np_oracle as O          # noqa: E402
from oracle import ref_harness as R        # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def ids_for(cfg, B, S, seed=0):
    return np.random.default_rng(seed).integers(0, cfg.vocab_size, (B, S)).astype(np.int64)


def run_case(mod, cfg, wseed, std, jitter, B, S, decode_steps=0, ids_seed=0):
    """Prefill S tokens (with a cache iff decode_steps > 0), then decode_steps 1-token calls."""
    w = O.synth_weights(cfg, seed=wseed, std=std, norm_jitter=jitter)
    model, rcfg = R.build_reference_model(mod, cfg, w)
    ids = ids_for(cfg, B, S + decode_steps, ids_seed)
    cache = mod.KVCache() if decode_steps else None
    t = time.perf_counter()
    _, logits, cache, hs, att = model(ids[:, :S], use_cache=bool(decode_steps), kv_cache=cache)
    dt = time.perf_counter() - t
    out = {"ids": ids, "logits": np.asarray(logits), "hidden": np.stack([np.asarray(h) for h in hs]),
           "attn0": np.asarray(att[0])}
    steps = []
    for i in range(decode_steps):
        _, lg, cache, _, _ = model(ids[:, S + i:S + i + 1], use_cache=True, kv_cache=cache)
        steps.append(np.asarray(lg))
    if steps:
        out["decode_logits"] = np.stack(steps)
    return out, dt, str(np.asarray(logits).dtype)


def save(name, meta, arrays):
    os.makedirs(OUT, exist_ok=True)
    meta = dict(meta, numpy=np.__version__)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **arrays)
    print(f"[golden] {name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in arrays.items()), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-1b", action="store_true")
    ap.add_argument("--only-1b", action="store_true")
    a = ap.parse_args()

    if not a.only_1b:
        lm = R.load_llama_numpy()
        for name, cname, B, S, dec in [("llama_tiny_prefill", "tiny", 2, 12, 0),
                                       ("llama_tiny_decode", "tiny", 2, 5, 3),
                                       ("llama_tiny_qlen2", "tiny", 1, 2, 0),
                                       ("llama_tiny_qlen1", "tiny", 1, 1, 0),
                                       ("llama_tiny128_prefill", "tiny128", 1, 9, 0),
                                       ("llama_tiny_long", "tiny", 1, 200, 0)]:
            cfg = O.llama_config(cname)
            out, dt, dty = run_case(lm, cfg, wseed=1, std=0.08, jitter=0.1, B=B, S=S, decode_steps=dec)
            meta = dict(family="llama", config=cname, wseed=1, std=0.08, jitter=0.1, B=B, S=S,
                        decode_steps=dec, ids_seed=0, logits_dtype=dty, ref="llama3.2_model_numpy.py")
            save(name, meta, out)
        # chunk-on-cache must raise in the reference: record the fact
        cfg = O.llama_config("tiny")
        w = O.synth_weights(cfg, seed=1, std=0.08, norm_jitter=0.1)
        model, _ = R.build_reference_model(lm, cfg, w)
        ids = ids_for(cfg, 1, 8)
        cache = lm.KVCache()
        model(ids[:, :4], use_cache=True, kv_cache=cache)
        try:
            model(ids[:, 4:8], use_cache=True, kv_cache=cache)
            raised = ""
        except ValueError as e:
            raised = str(e)
        save("llama_tiny_chunk_on_cache", dict(family="llama", config="tiny", raised=raised), {"ids": ids})

        gm = R.load_gemma2()
        for name, B, S, dec in [("gemma2_tiny_prefill", 2, 10, 0), ("gemma2_tiny_decode", 1, 6, 3)]:
            cfg = O.gemma2_config("tiny")
            out, dt, dty = run_case(gm, cfg, wseed=2, std=0.05, jitter=0.1, B=B, S=S, decode_steps=dec)
            meta = dict(family="gemma2", config="tiny", wseed=2, std=0.05, jitter=0.1, B=B, S=S,
                        decode_steps=dec, ids_seed=0, logits_dtype=dty, ref="gemma2_model.py (cupy->numpy shim)")
            save(name, meta, out)

    if not a.skip_1b:
        # BASELINE.json config #1: Llama-3.2-1B shapes, 16-token prompt, random-init weights.
        lm = R.load_llama_numpy()
        cfg = O.llama_config("1b")
        out, dt, dty = run_case(lm, cfg, wseed=0, std=0.02, jitter=0.0, B=1, S=16)
        lg = out["logits"]
        arrays = {
            "ids": out["ids"],
            "logits_sub": lg[:, :, ::61].astype(np.float32),          # every 61st vocab column
            "logits_last": lg[:, -1, :].astype(np.float32),           # full last-position row
            "argmax": lg.argmax(-1).astype(np.int64),
            "logits_stats": np.array([lg.mean(), lg.std(), np.abs(lg).max()], dtype=np.float64),
            "hidden_sub": out["hidden"][:, :, :, ::8].astype(np.float32),
            "hidden_absmax": np.abs(out["hidden"]).max(axis=(1, 2, 3)),
        }
        meta = dict(family="llama", config="1b", wseed=0, std=0.02, jitter=0.0, B=1, S=16, decode_steps=0,
                    ids_seed=0, logits_dtype=dty, ref="llama3.2_model_numpy.py", ref_seconds=dt,
                    cores=os.cpu_count())
        save("llama_1b_s16", meta, arrays)


if __name__ == "__main__":
    main()
