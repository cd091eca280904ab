"""Generates tests/golden/hf_llama_tiny_rope3.npz by running the installed `transformers` LlamaForCausalLM (fp32, CPU,
eager attention) on a seeded synthetic checkpoint with llama3 RoPE frequency scaling.

TEST INFRASTRUCTURE.  This pins the HF-faithful OPTIONS the reference lacks (llama3 rope
scaling, causal mask for every q_len) against the real HF implementation, the same way make_golden.py pins the
reference-parity path against the reference.  Run here (transformers importable, no GPU needed):
    python oracle/make_golden_hf.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import np_oracle as O  # noqa: E402

ROPE_SCALING = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                "original_max_position_embeddings": 64}


def main():
    import transformers
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = O.llama_config("tiny")
    wseed, std, jitter = 7, 0.08, 0.1
    w = O.synth_weights(cfg, wseed, std, jitter)
    hf_cfg = LlamaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                         num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                         num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_norm_eps,
                         rope_theta=cfg.rope_theta, rope_scaling=dict(ROPE_SCALING), tie_word_embeddings=True,
                         attention_bias=False, mlp_bias=False, hidden_act="silu", max_position_embeddings=512,
                         attn_implementation="eager")
    model = LlamaForCausalLM(hf_cfg).to(torch.float32).eval()
    sd = {k: v.to(torch.float32) for k, v in w.items()}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    inv = model.model.rotary_emb.inv_freq.detach().numpy().astype(np.float32)
    B, S, dec = 2, 150, 3
    ids = np.random.default_rng(23).integers(0, cfg.vocab_size, (B, S + dec))
    with torch.no_grad():
        out = model(torch.from_numpy(ids[:, :S]), output_hidden_states=True, use_cache=True)
        logits = out.logits.numpy()
        hidden = np.stack([h.numpy() for h in out.hidden_states])
        past = out.past_key_values
        dlog = []
        for i in range(dec):
            o2 = model(torch.from_numpy(ids[:, S + i:S + i + 1]), past_key_values=past, use_cache=True)
            past = o2.past_key_values
            dlog.append(o2.logits.numpy())
        # q_len = 2 on an empty cache: HF masks causally, the reference does not (llama3.2_model_numpy.py:479)
        q2 = model(torch.from_numpy(ids[:, :2])).logits.numpy()
    meta = dict(family="llama", config="tiny", wseed=wseed, std=std, jitter=jitter, S=S, decode_steps=dec,
                rope_scaling=ROPE_SCALING, transformers=transformers.__version__, torch=torch.__version__,
                numpy=np.__version__, generator="oracle/make_golden_hf.py")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hf_llama_tiny_rope3.npz")
    np.savez_compressed(path, ids=ids, logits=logits.astype(np.float32), hidden=hidden.astype(np.float32),
                        decode_logits=np.stack(dlog).astype(np.float32), qlen2_logits=q2.astype(np.float32), inv_freq=inv,
                        meta=json.dumps(meta))
    print("wrote", path, os.path.getsize(path), "bytes; logits std", logits.std())


if __name__ == "__main__":
    main()
