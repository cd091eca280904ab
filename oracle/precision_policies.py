"""TEST INFRASTRUCTURE / ANALYSIS ONLY -- CPU emulation of operand-precision policies for the Llama forward pass.

Question: which is the cheapest operand policy that still meets BASELINE.json's literal "logits within
1e-2 max-abs of the NumPy reference" at logit-std ~0.9?  The engine implements two policies (all-bf16 operands = the fast path,
all hi+lo split-bf16 = the parity mode); this script emulates those and the policies in between on the CPU (PyTorch fp32
arithmetic with the chosen tensors rounded the way a tensor-core operand would be) at BASELINE config #1 (Llama-3.2-1B shapes,
16 tokens) and compares against the reference's own logits (tests/golden/llama_1b_s16.npz).  Method = SURVEY.md Appendix C.

    python oracle/precision_policies.py            # ~3 min on 8 cores -> table on stdout (kept as profiles/r02_precision_policies.md)

Rounding modes:  bf16 (8-bit mantissa, round-to-nearest-even) | hilo (hi + lo bf16 pair: ~16 mantissa bits, what the parity
mode carries) | tf32 (10-bit mantissa, what kind::tf32 operands keep) | fp32.
Operand groups:  gemm = A operands of the weight GEMMs (xn, att, h, final xn); qk = Q and K of S = Q.K^T; pv = P and V of O = P.V.
Cost model (last column): time of a 1B / S=2048 step relative to the fast path, from the measured shares of the fast step (GEMMs
76 %, attention 17 %, rest 7 %): a hilo GEMM operand doubles the contraction (2x), hilo attention operands need 3 products per
product, tf32 operands run the tensor pipe at half the bf16 rate (2x).  Measured: fast 1.00, all-hilo 2.05 (bench.py "precise").
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import np_oracle as O  # noqa: E402


def rnd(x, mode):
    if mode == "fp32":
        return x
    if mode == "bf16":
        return x.to(torch.bfloat16).float()
    if mode == "hilo":
        hi = x.to(torch.bfloat16).float()
        return hi + (x - hi).to(torch.bfloat16).float()
    if mode == "tf32":                      # keep 10 explicit mantissa bits (round to nearest)
        i = x.contiguous().view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32)
    raise KeyError(mode)


def forward(cfg, W, ids, pol):
    """The reference's Llama forward (llama3.2_model_numpy.py:526-586, :407-516) in fp32 torch with operand rounding."""
    B, S = ids.shape
    NH, NKV, HD, H = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.hidden_size
    x = W["model.embed_tokens.weight"][ids]
    pos = torch.arange(S, dtype=torch.float32)
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, HD, 2, dtype=torch.float32) / HD))
    fr = torch.outer(pos, inv)
    emb = torch.cat([fr, fr], -1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]

    def norm(t, g):
        return t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + cfg.rms_norm_eps) * g

    def rot(t):
        h = t.shape[-1] // 2
        return torch.cat([-t[..., h:], t[..., :h]], -1)

    mask = torch.tril(torch.ones(S, S, dtype=torch.bool))
    for li in range(cfg.num_hidden_layers):
        p = f"model.layers.{li}."
        a = rnd(norm(x, W[p + "input_layernorm.weight"]), pol["gemm"])
        q = (a @ W[p + "self_attn.q_proj.weight"].T).view(B, S, NH, HD).transpose(1, 2)
        k = (a @ W[p + "self_attn.k_proj.weight"].T).view(B, S, NKV, HD).transpose(1, 2)
        v = (a @ W[p + "self_attn.v_proj.weight"].T).view(B, S, NKV, HD).transpose(1, 2)
        q, k = q * cos + rot(q) * sin, k * cos + rot(k) * sin
        q, k, v = rnd(q, pol["qk"]), rnd(k, pol["qk"]), rnd(v, pol["pv"])
        k, v = k.repeat_interleave(NH // NKV, 1), v.repeat_interleave(NH // NKV, 1)
        s = (q @ k.transpose(-1, -2)) / np.sqrt(HD)
        if S > 2:
            s = s.masked_fill(~mask, float("-inf"))
        pr = rnd(torch.softmax(s, -1), pol["pv"])
        o = (pr @ v).transpose(1, 2).reshape(B, S, NH * HD)
        x = x + rnd(o, pol["gemm"]) @ W[p + "self_attn.o_proj.weight"].T
        a = rnd(norm(x, W[p + "post_attention_layernorm.weight"]), pol["gemm"])
        g = a @ W[p + "mlp.gate_proj.weight"].T
        u = a @ W[p + "mlp.up_proj.weight"].T
        x = x + rnd(torch.nn.functional.silu(g) * u, pol["gemm"]) @ W[p + "mlp.down_proj.weight"].T
    a = rnd(norm(x, W["model.norm.weight"]), pol["gemm"])
    return a @ W["model.embed_tokens.weight"].T


COST = {"bf16": 1.0, "hilo_gemm": 2.0, "hilo_attn": 3.0, "tf32": 2.0, "fp32": None}
POLICIES = [
    ("fast path: bf16 everywhere (implemented, benchmarked)", dict(gemm="bf16", qk="bf16", pv="bf16")),
    ("hi+lo GEMM operands, bf16 attention", dict(gemm="hilo", qk="bf16", pv="bf16")),
    ("bf16 GEMM operands, hi+lo attention", dict(gemm="bf16", qk="hilo", pv="hilo")),
    ("hi+lo GEMM + hi+lo Q,K; bf16 P,V", dict(gemm="hilo", qk="hilo", pv="bf16")),
    ("hi+lo GEMM + bf16 Q,K; hi+lo P,V", dict(gemm="hilo", qk="bf16", pv="hilo")),
    ("hi+lo GEMM + tf32 attention operands (kind::tf32)", dict(gemm="hilo", qk="tf32", pv="tf32")),
    ("tf32 everywhere", dict(gemm="tf32", qk="tf32", pv="tf32")),
    ("tf32 GEMM operands, hi+lo attention", dict(gemm="tf32", qk="hilo", pv="hilo")),
    ("parity mode: hi+lo everywhere (implemented, measured 2.05x)", dict(gemm="hilo", qk="hilo", pv="hilo")),
    ("fp32 everywhere (sanity of the emulation; reference is float64)", dict(gemm="fp32", qk="fp32", pv="fp32")),
]


def est_cost(pol):
    g = {"bf16": 1.0, "hilo": 2.0, "tf32": 2.0, "fp32": float("nan")}[pol["gemm"]]
    a = {"bf16": 1.0, "hilo": 3.0, "tf32": 2.0, "fp32": float("nan")}
    att = 0.5 * a[pol["qk"]] + 0.5 * a[pol["pv"]]
    return 0.76 * g + 0.17 * att + 0.07


def main():
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "llama_1b_s16.npz"))
    meta = json.loads(str(z["meta"]))
    cfg = O.llama_config(meta["config"])
    w = O.synth_weights(cfg, meta["wseed"], meta["std"], meta["jitter"])
    W = {k: v.float() for k, v in w.items()}
    ids = torch.from_numpy(z["ids"])
    sub, last, std = z["logits_sub"].astype(np.float64), z["logits_last"].astype(np.float64), float(z["logits_stats"][1])
    torch.set_num_threads(os.cpu_count() or 1)
    print(f"Llama-3.2-1B shapes, 16 tokens, reference logit std {std:.3f}; error vs the reference's own logits (2.1 M values)\n")
    print("| operand policy | max-abs | mean-abs | max/std | meets 1e-2 | est. step cost vs fast |")
    print("|---|---|---|---|---|---|")
    with torch.no_grad():
        for name, pol in POLICIES:
            lg = forward(cfg, W, ids, pol).double().numpy()
            d = np.concatenate([np.abs(lg[:, :, ::61] - sub).ravel(), np.abs(lg[:, -1, :] - last).ravel()])
            c = est_cost(pol)
            print(f"| {name} | {d.max():.2e} | {d.mean():.2e} | {d.max() / std:.4f} | {'yes' if d.max() <= 1e-2 else 'no'} | "
                  f"{'-' if c != c else f'{c:.2f}x'} |", flush=True)


if __name__ == "__main__":
    main()
