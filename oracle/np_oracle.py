"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference forward pass (the oracle).

This is a from-scratch functional restatement of what githubpradeep/llm_np_cp computes on
its hot path ``model(input_ids) -> logits`` (+ KV cache), written as one pure function per
block.  Every function cites the reference lines it follows (``file:line`` is relative to
/root/reference).  It intentionally reproduces the reference's *dtype flow* under NumPy >= 2
(fp32 until the first ``/ np.sqrt(head_dim)``, float64 afterwards -- SURVEY.md App. B Q5) and
its quirks (causal mask only when q_len > 2, un-stabilised softmax, tied lm_head, [q,q] mask
that cannot broadcast onto a non-empty cache).

Pinning status: PINNED.  ``oracle/make_golden.py`` executes the unmodified reference modules
(``oracle/ref_harness.py``) on seeded synthetic weights and stores their outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against them.
The reference itself ships no golden vectors or tests.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np


# --------------------------------------------------------------------------------------
# configs (public HF shapes)
# --------------------------------------------------------------------------------------
class Config(dict):
    """dict with attribute access, like the reference's AttributeDict
    (llama3.2_model_numpy.py:212-215)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def llama_config(name: str = "1b", **over) -> Config:
    base = dict(family="llama", vocab_size=128256, hidden_size=2048, intermediate_size=8192,
                num_hidden_layers=16, num_attention_heads=32, num_key_value_heads=8, head_dim=64,
                max_position_embeddings=131072, rope_theta=500000.0, rms_norm_eps=1e-5,
                hidden_act="silu", attention_dropout=0.0, use_cache=False)
    if name == "3b":
        base.update(hidden_size=3072, num_hidden_layers=28, num_attention_heads=24,
                    num_key_value_heads=8, head_dim=128)
    elif name == "tiny":
        base.update(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                    num_attention_heads=4, num_key_value_heads=2, head_dim=64)
    elif name == "tiny128":   # head_dim 128 like Llama-3.2-3B
        base.update(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                    num_attention_heads=2, num_key_value_heads=1, head_dim=128)
    elif name != "1b":
        raise KeyError(name)
    base.update(over)
    return Config(base)


def gemma2_config(name: str = "2b", **over) -> Config:
    base = dict(family="gemma2", vocab_size=256000, hidden_size=2304, intermediate_size=9216,
                num_hidden_layers=26, num_attention_heads=8, num_key_value_heads=4, head_dim=256,
                max_position_embeddings=8192, rope_theta=10000.0, rms_norm_eps=1e-6,
                hidden_act="gelu_pytorch_tanh", attention_dropout=0.0, use_cache=False,
                pad_token_id=0, query_pre_attn_scalar=256, final_logit_softcapping=30.0,
                attn_logit_softcapping=50.0, sliding_window=4096)
    if name == "tiny":
        base.update(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                    num_attention_heads=2, num_key_value_heads=1, head_dim=256,
                    query_pre_attn_scalar=256)
    elif name != "2b":
        raise KeyError(name)
    base.update(over)
    return Config(base)


def weight_names(cfg) -> List[str]:
    """HF tensor names in the fixed order used for seeded synthetic weights
    (names: llama3.2_model_numpy.py:164-168, :250-252, :377-380, :597;
    gemma2_model.py:590-591 adds the two feed-forward norms)."""
    names = ["model.embed_tokens.weight"]
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        names += [p + "input_layernorm.weight",
                  p + "self_attn.q_proj.weight", p + "self_attn.k_proj.weight",
                  p + "self_attn.v_proj.weight", p + "self_attn.o_proj.weight",
                  p + "post_attention_layernorm.weight"]
        if cfg.family == "gemma2":
            names += [p + "pre_feedforward_layernorm.weight",
                      p + "post_feedforward_layernorm.weight"]
        names += [p + "mlp.gate_proj.weight", p + "mlp.up_proj.weight", p + "mlp.down_proj.weight"]
    names.append("model.norm.weight")
    return names


def weight_shape(cfg, name: str):
    H, I = cfg.hidden_size, cfg.intermediate_size
    q = cfg.num_attention_heads * cfg.head_dim
    kv = cfg.num_key_value_heads * cfg.head_dim
    if name.endswith("embed_tokens.weight"):
        return (cfg.vocab_size, H)
    if name.endswith("norm.weight") or name.endswith("layernorm.weight"):
        return (H,)
    if name.endswith("q_proj.weight"):
        return (q, H)
    if name.endswith("k_proj.weight") or name.endswith("v_proj.weight"):
        return (kv, H)
    if name.endswith("o_proj.weight"):
        return (H, q)
    if name.endswith("gate_proj.weight") or name.endswith("up_proj.weight"):
        return (I, H)
    if name.endswith("down_proj.weight"):
        return (H, I)
    raise KeyError(name)


def synth_weights(cfg, seed: int = 0, std: float = 0.02, norm_jitter: float = 0.0):
    """Seeded synthetic checkpoint as {hf_name: torch.bfloat16 tensor}: every value is exactly
    representable in bf16, so the GPU engine and the fp32/f64 oracle share bit-identical weights
    (SURVEY.md section 8d config #1).  Norm gains are 1 (Llama) / 0 (Gemma, whose norm adds 1,
    gemma2_model.py:334) plus an optional bf16-rounded jitter so the gain path is exercised."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = {}
    for n in weight_names(cfg):
        shp = weight_shape(cfg, n)
        if len(shp) == 1:
            base = 0.0 if cfg.family == "gemma2" else 1.0
            w = torch.full(shp, base, dtype=torch.float32)
            if norm_jitter:
                w = w + norm_jitter * torch.randn(shp, generator=g, dtype=torch.float32)
        else:
            w = torch.randn(shp, generator=g, dtype=torch.float32) * std
        out[n] = w.to(torch.bfloat16)
    return out


def to_numpy_f32(weights) -> Dict[str, np.ndarray]:
    """What the reference's load_weights does to every tensor: upcast to fp32 NumPy
    (llama3.2_model_numpy.py:1024-1028)."""
    import torch
    return {k: v.to(torch.float32).numpy() for k, v in weights.items()}


# --------------------------------------------------------------------------------------
# array ops (L0)
# --------------------------------------------------------------------------------------
# Parity tests at the benchmarked shapes run hundreds of one-token steps over GB-sized weights: casting W to the
# activation dtype (float64 after the first attention, App. B Q5) on every call would dominate.  Opt-in cache of the
# cast copies, keyed by the weight array's identity; same values, so results are unchanged.
_CAST_CACHE: Dict[tuple, np.ndarray] = {}
CACHE_CASTS = False
BY_HEAD_MIN_ELEMS = 1 << 26      # score-matrix size above which attention() evaluates head by head (same arithmetic)


def clear_cast_cache():
    _CAST_CACHE.clear()


def _cast(w, dtype):
    if w.dtype == dtype:
        return w
    if not CACHE_CASTS:
        return w.astype(dtype)
    key = (id(w), np.dtype(dtype).str)
    hit = _CAST_CACHE.get(key)
    if hit is None:
        hit = _CAST_CACHE[key] = w.astype(dtype)
    return hit


def linear(x, w, faithful: bool):
    """y = x . W^T with W stored [out, in]; no bias is ever loaded.
    Reference: Linear_np.__call__, llama3.2_model_numpy.py:134-144 -- an N-D ``np.dot`` into a
    preallocated ``out`` of x's dtype.  ``faithful=True`` issues the same NumPy call (slow N-D
    path; this is what the CPU baseline times).  ``faithful=False`` runs the same product as a
    2-D GEMM in x's dtype (same maths, different summation order; used by the parity tests)."""
    if faithful:
        out = np.empty((*x.shape[:-1], w.shape[0]), dtype=x.dtype)
        np.dot(x, w.T, out=out)
        return out
    x2 = x.reshape(-1, x.shape[-1])
    y = x2 @ _cast(w, x.dtype).T
    return y.reshape(*x.shape[:-1], w.shape[0])


def rms_norm(x, gamma, eps, plus_one: bool = False):
    """x * 1/sqrt(mean(x^2) + eps) * gamma  (llama3.2_model_numpy.py:256-278);
    Gemma-2 uses (gamma + 1) (gemma2_model.py:334)."""
    g = gamma + 1.0 if plus_one else gamma
    n = x * np.reciprocal(np.sqrt(np.mean(np.square(x), axis=-1, keepdims=True) + eps))
    return n * g


def llama3_scaled_inv_freq(inv_freq, rope_scaling):
    """HF "llama3" RoPE frequency scaling, fp32 like transformers' modeling_rope_utils._compute_llama3_parameters
    (the reference never applies it; this is an Extras option pinned against transformers in
    tests/test_oracle_golden.py)."""
    f32 = np.float32
    factor, low_ff = f32(rope_scaling["factor"]), f32(rope_scaling.get("low_freq_factor", 1.0))
    high_ff = f32(rope_scaling.get("high_freq_factor", 4.0))
    old = f32(rope_scaling.get("original_max_position_embeddings", 8192))
    inv_freq = inv_freq.astype(f32)
    wavelen = f32(2 * np.pi) / inv_freq
    scaled = np.where(wavelen > old / low_ff, inv_freq / factor, inv_freq).astype(f32)
    smooth = ((old / wavelen - low_ff) / (high_ff - low_ff)).astype(f32)
    blended = ((f32(1) - smooth) * scaled / factor + smooth * scaled).astype(f32)
    medium = ~(wavelen < old / high_ff) & ~(wavelen > old / low_ff)
    return np.where(medium, blended, scaled).astype(f32)


def rope_tables(head_dim: int, base: float, position_ids, dtype, rope_scaling=None):
    """cos/sin [1,S,HD] from plain RoPE -- no llama3 frequency scaling unless the Extras option asks for it
    (llama3.2_model_numpy.py:42-60).  inv_freq and the angle product are fp32."""
    inv_freq = 1.0 / (base ** (np.arange(0, head_dim, 2).astype(np.float32) / head_dim))
    if rope_scaling:
        inv_freq = llama3_scaled_inv_freq(inv_freq, rope_scaling)
    inv = np.expand_dims(np.expand_dims(inv_freq, 0).astype(np.float32), -1)
    inv = inv.repeat(position_ids.shape[0], axis=0)
    pos = np.expand_dims(position_ids, 1).astype(np.float32)
    freqs = np.matmul(inv, pos).transpose(0, 2, 1)
    emb = np.concatenate((freqs, freqs), axis=-1)
    return np.cos(emb).astype(dtype), np.sin(emb).astype(dtype)


def rotate_half(x):
    """[-x2, x1] with the halves split at HD/2 (llama3.2_model_numpy.py:69-83)."""
    h = x.shape[-1] // 2
    return np.concatenate((-x[..., h:], x[..., :h]), axis=-1)


def apply_rope(q, k, cos, sin):
    """llama3.2_model_numpy.py:85-90 (cos/sin broadcast over the head axis)."""
    c, s = np.expand_dims(cos, 1), np.expand_dims(sin, 1)
    return q * c + rotate_half(q) * s, k * c + rotate_half(k) * s


def repeat_kv(x, n_rep: int):
    """[B,NKV,T,HD] -> [B,NKV*n_rep,T,HD]; kv head j serves q heads j*n_rep..j*n_rep+n_rep-1
    (llama3.2_model_numpy.py:188-204)."""
    if n_rep == 1:
        return x
    b, nkv, t, d = x.shape
    return np.tile(np.expand_dims(x, 2), (1, 1, n_rep, 1, 1)).reshape(b, nkv * n_rep, t, d)


def softmax_plain(x):
    """The live NumPy definition: exp(x)/sum(exp(x)), no max subtraction
    (llama3.2_model_numpy.py:915-919; App. B Q1)."""
    e = np.exp(x)
    return e / np.sum(e, axis=-1, keepdims=True)


def softmax_stable_f32(x):
    """The live CuPy definition: RawKernel with max subtraction, input cast to fp32
    (llama3.2_model.py:924-975, gemma2_model.py:987-1038)."""
    x = x.astype(np.float32)
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - m)
    return (e / np.sum(e, axis=-1, keepdims=True)).astype(np.float32)


def silu(x):
    """x * 1/(1+exp(-x))  (llama3.2_model_numpy.py:101-105)."""
    return x * (1 / (1 + np.exp(-x)))


def gelu_tanh(x):
    """0.5x(1+tanh(sqrt(2/pi)(x+0.044715x^3)))  (gemma2_model.py:168-169)."""
    return 0.5 * x * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * np.power(x, 3.0))))


ACT = {"silu": silu, "gelu_pytorch_tanh": gelu_tanh}


# --------------------------------------------------------------------------------------
# KV cache (L1)
# --------------------------------------------------------------------------------------
class KVCache:
    """Per-layer lists grown by concatenation along T; holds post-RoPE K and V, un-repeated
    (llama3.2_model_numpy.py:311-340)."""

    def __init__(self):
        self.key_cache: List[np.ndarray] = []
        self.value_cache: List[np.ndarray] = []

    def num_items(self) -> int:
        return 0 if not self.key_cache else self.key_cache[0].shape[-2]

    def update(self, k, v, layer_idx: int):
        if len(self.key_cache) <= layer_idx:
            self.key_cache.append(k)
            self.value_cache.append(v)
        else:
            self.key_cache[layer_idx] = np.concatenate([self.key_cache[layer_idx], k], axis=-2)
            self.value_cache[layer_idx] = np.concatenate([self.value_cache[layer_idx], v], axis=-2)
        return self.key_cache[layer_idx], self.value_cache[layer_idx]


# --------------------------------------------------------------------------------------
# blocks (L1/L2) and the model (L3)
# --------------------------------------------------------------------------------------
@dataclass
class Extras:
    """Options the reference does NOT implement but BASELINE.json asks
    for; semantics follow transformers 5.5 modeling_gemma2.py.  All default to the reference's
    behaviour (off)."""
    attn_softcap: Optional[float] = None       # tanh(s/cap)*cap after scaling, before the mask
    sliding_window: Optional[int] = None       # kv_idx > q_idx - window on "sliding" layers
    sliding_layers: object = "even"            # HF default layer_types pattern: even layers slide; or a per-layer list of
                                               # booleans / "sliding_attention" strings (config.layer_types)
    attn_scale: Optional[float] = None         # e.g. query_pre_attn_scalar ** -0.5
    always_causal: bool = False                # drop the q_len > 2 quirk
    rope_scaling: Optional[dict] = None        # HF llama3 frequency scaling {factor, low_freq_factor, high_freq_factor, original_max_position_embeddings}


def _slides(ex: Extras, li: int) -> bool:
    """Does layer li apply the sliding window?  (transformers modeling_gemma2: config.layer_types[li] ==
    "sliding_attention"; default pattern = even layers.)"""
    if ex.sliding_window is None:
        return False
    if isinstance(ex.sliding_layers, str):
        return (li % 2 == 0) if ex.sliding_layers == "even" else True
    t = ex.sliding_layers[li]
    return t == "sliding_attention" or t is True or t == 1


def _keep_mask(S, T, li, ex: Extras):
    """Boolean [S,T] keep-mask of the score matrix, or None when the reference applies no mask (q_len <= 2,
    llama3.2_model_numpy.py:479).  Same predicate as the in-line code of attention() below."""
    if not (S > 2 or ex.always_causal):
        return None
    if ex.always_causal or ex.sliding_window is not None:
        qi = np.arange(T - S, T)[:, None]
        ki = np.arange(T)[None, :]
        keep = ki <= qi
        if _slides(ex, li):
            keep &= ki > qi - ex.sliding_window
        return keep
    if S != T:   # the reference's [q,q] tril cannot broadcast onto [q,kv] (App. B Q4)
        raise ValueError(f"operands could not be broadcast together with shapes (1,1,{S},{S}) () (B,NH,{S},{T})")
    return np.tril(np.ones((S, S), dtype=np.float32)) != 0


def _attention_core_by_head(q, k, v, n_rep, HD, S, T, li, stable_softmax, ex: Extras, keep_attn):
    """scores -> mask -> softmax -> P.V of attention(), evaluated one query head at a time on a small thread pool
    with in-place ufuncs (NumPy releases the GIL inside them).  Exactly the per-element arithmetic of the block
    below -- repeat_kv becomes head indexing (kv head = q head // n_rep, :188-204), np.where becomes a masked
    assignment -- so long sequences (S = 8192: 12.9 GB of float64 scores per layer otherwise) stay tractable for
    the parity tests at the benchmarked shapes."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    B, NH = q.shape[0], q.shape[1]
    keep = _keep_mask(S, T, li, ex)
    drop = None if keep is None else ~keep
    out = np.empty((B, NH, S, HD), dtype=np.float32 if stable_softmax else np.float64)
    probs = np.empty((B, NH, S, T), dtype=out.dtype) if keep_attn else None

    def one(h):
        kh, vh = k[:, h // n_rep], v[:, h // n_rep]
        if ex.attn_scale is None:
            s = np.matmul(q[:, h], kh.transpose(0, 2, 1)) / np.sqrt(HD)          # float64 under NumPy 2 (:475-477)
        else:
            s = np.matmul(q[:, h], kh.transpose(0, 2, 1)) * np.float64(ex.attn_scale)
        if ex.attn_softcap is not None:
            s /= ex.attn_softcap
            np.tanh(s, out=s)
            s *= ex.attn_softcap
        if drop is not None:
            s[:, drop] = float("-inf")
        if stable_softmax:                                                       # CuPy twin: fp32, max-subtracted
            s = s.astype(np.float32)
            s -= np.max(s, axis=-1, keepdims=True)
        np.exp(s, out=s)
        s /= np.sum(s, axis=-1, keepdims=True)
        out[:, h] = np.matmul(s, vh)
        if probs is not None:
            probs[:, h] = s

    with ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 1) // 4))) as pool:
        list(pool.map(one, range(NH)))
    return out, probs


def attention(cfg, W, li, x, cos, sin, cache, faithful, stable_softmax, ex: Extras, keep_attn: bool = False):
    """LlamaAttention.__call__ (llama3.2_model_numpy.py:407-516) / Gemma2Attention
    (gemma2_model.py:480-582)."""
    B, S, _ = x.shape
    NH, NKV, HD = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    p = f"model.layers.{li}.self_attn."
    q = linear(x, W[p + "q_proj.weight"], faithful).reshape(B, S, NH, HD).transpose(0, 2, 1, 3)
    k = linear(x, W[p + "k_proj.weight"], faithful).reshape(B, S, NKV, HD).transpose(0, 2, 1, 3)
    v = linear(x, W[p + "v_proj.weight"], faithful).reshape(B, S, NKV, HD).transpose(0, 2, 1, 3)
    q, k = apply_rope(q, k, cos, sin)
    if cache is not None:
        k, v = cache.update(k, v, li)
    T = k.shape[-2]
    if not faithful and B * NH * S * T > BY_HEAD_MIN_ELEMS:
        o, pr = _attention_core_by_head(q, k, v, NH // NKV, HD, S, T, li, stable_softmax, ex, keep_attn)
        return linear(o.transpose(0, 2, 1, 3).reshape(B, S, NH * HD), W[p + "o_proj.weight"], faithful), pr
    kr, vr = repeat_kv(k, NH // NKV), repeat_kv(v, NH // NKV)
    if ex.attn_scale is None:
        # f32 / np.float64 scalar -> float64 under NumPy 2 (:475-477, App. B Q5)
        s = np.matmul(q, kr.transpose(0, 1, 3, 2)) / np.sqrt(HD)
    else:
        s = np.matmul(q, kr.transpose(0, 1, 3, 2)) * np.float64(ex.attn_scale)
    if ex.attn_softcap is not None:
        s = np.tanh(s / ex.attn_softcap) * ex.attn_softcap
    if S > 2 or ex.always_causal:
        # reference: tril(q_len, q_len) broadcast onto [B,NH,q_len,kv_len] (:479-486); with a
        # non-empty cache and q_len > 1 this raises, exactly like the reference (App. B Q4).
        if ex.always_causal or ex.sliding_window is not None:
            qi = np.arange(T - S, T)[:, None]
            ki = np.arange(T)[None, :]
            keep = ki <= qi
            if _slides(ex, li):
                keep &= ki > qi - ex.sliding_window
            s = np.where(keep[None, None], s, float("-inf"))
        else:
            tril = np.tril(np.ones((S, S), dtype=np.float32))
            s = np.where(tril[None, None, :, :] == 0, float("-inf"), s)
    pr = softmax_stable_f32(s) if stable_softmax else softmax_plain(s)
    o = np.matmul(pr, vr).transpose(0, 2, 1, 3).reshape(B, S, NH * HD)
    return linear(o, W[p + "o_proj.weight"], faithful), pr


def mlp(cfg, W, li, x, faithful):
    """down(act(gate(x)) * up(x))  (llama3.2_model_numpy.py:176-182; gemma2_model.py:260-267)."""
    p = f"model.layers.{li}.mlp."
    g = linear(x, W[p + "gate_proj.weight"], faithful)
    u = linear(x, W[p + "up_proj.weight"], faithful)
    return linear(ACT[cfg.hidden_act](g) * u, W[p + "down_proj.weight"], faithful)


@dataclass
class Output:
    logits: np.ndarray
    hidden_states: List[np.ndarray]
    attentions: List[np.ndarray] = field(default_factory=list)
    cache: Optional[KVCache] = None


def forward(cfg, W: Dict[str, np.ndarray], input_ids, cache: Optional[KVCache] = None, *,
            faithful: bool = False, keep_attn: bool = False,
            extras: Optional[Extras] = None, logit_positions=None, inputs_embeds=None,
            position_ids=None, use_cache: bool = True) -> Output:
    """model(input_ids) -> logits, L+1 hidden states (and optionally L attention maps).

    Llama: LlamaForCausalLM_np.__call__ -> LlamaModel.__call__ -> LlamaDecoderLayer.__call__
    (llama3.2_model_numpy.py:777-830, :624-732, :526-586).
    Gemma-2: gemma2_model.py:830-886, :680-785, :594-651 (4 norms, embed scale, (1+gamma),
    GeGLU, final tanh soft-cap; stable fp32 softmax because that file is CuPy-only)."""
    ex = extras or Extras()
    gem = cfg.family == "gemma2"
    if inputs_embeds is None:
        ids = np.asarray(input_ids)
        B, S = ids.shape
    else:
        B, S = inputs_embeds.shape[:2]
    # positions continue from the cache only when use_cache is set (:659-672); the cache is appended to either way
    t0 = cache.num_items() if (cache is not None and use_cache) else 0
    pos = np.expand_dims(np.arange(t0, t0 + S, dtype=np.int64), 0)         # [1,S]
    if position_ids is not None:
        pos = np.asarray(position_ids, dtype=np.int64).reshape(1, S)       # caller positions win (:657)
    x = W["model.embed_tokens.weight"][ids] if inputs_embeds is None else np.asarray(inputs_embeds)   # (:675-676)
    if gem:
        x = x * (cfg.hidden_size ** 0.5)                                   # gemma2_model.py:738-739
    hs, att = [], []
    cos = sin = None
    if not gem:
        cos, sin = rope_tables(cfg.head_dim, cfg.rope_theta, pos, x.dtype, ex.rope_scaling)  # once per forward (:682)
    for li in range(cfg.num_hidden_layers):
        hs.append(x)
        p = f"model.layers.{li}."
        if gem:
            # cos/sin are recomputed per layer in the activations' dtype (gemma2_model.py:506)
            h = rms_norm(x, W[p + "input_layernorm.weight"], cfg.rms_norm_eps, True)
            cos, sin = rope_tables(cfg.head_dim, cfg.rope_theta, pos, h.dtype, ex.rope_scaling)
            a, pr = attention(cfg, W, li, h, cos, sin, cache, faithful, True, ex, keep_attn)
            x = x + rms_norm(a, W[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps, True)
            h = rms_norm(x, W[p + "pre_feedforward_layernorm.weight"], cfg.rms_norm_eps, True)
            m = mlp(cfg, W, li, h, faithful)
            x = x + rms_norm(m, W[p + "post_feedforward_layernorm.weight"], cfg.rms_norm_eps, True)
        else:
            h = rms_norm(x, W[p + "input_layernorm.weight"], cfg.rms_norm_eps)
            a, pr = attention(cfg, W, li, h, cos, sin, cache, faithful, False, ex, keep_attn)
            x = x + a
            h = rms_norm(x, W[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
            x = x + mlp(cfg, W, li, h, faithful)
        if keep_attn:
            att.append(pr)
    x = rms_norm(x, W["model.norm.weight"], cfg.rms_norm_eps, gem)
    hs.append(x)
    if logit_positions is not None:        # the same lm_head product, evaluated at a subset of positions only
        x = x[:, np.asarray(logit_positions), :]
    logits = linear(x, W["model.embed_tokens.weight"], faithful)           # tied lm_head (:1025-1026)
    cap = cfg.get("final_logit_softcapping") if gem else None
    if cap is not None:
        logits = np.tanh(logits / cap) * cap                               # gemma2_model.py:867-870
    return Output(logits, hs, att, cache)


def min_p_probs(logits, p_base: float = 0.1):
    """The deterministic part of min_p_sampling (llama3.2_model_numpy.py:948-961): un-stabilised
    softmax, keep p >= p_base * p_max, renormalise.  The reference then draws with an unseeded
    torch.multinomial, so only this distribution is testable."""
    p = softmax_plain(logits)
    keep = p >= np.max(p, axis=-1, keepdims=True) * p_base
    p = p * keep.astype(np.float32)
    return p / p.sum(axis=-1, keepdims=True)


def flops_per_token(cfg, S: int) -> float:
    """Algorithmic FLOPs/token: 2*P_mm + causal attention 2*(S+1)*NH*HD*L."""
    H, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    q = cfg.num_attention_heads * cfg.head_dim
    kv = cfg.num_key_value_heads * cfg.head_dim
    p_mm = L * (H * q + 2 * H * kv + q * H + 3 * H * I) + cfg.vocab_size * H
    return 2.0 * p_mm + 2.0 * (S + 1) * cfg.num_attention_heads * cfg.head_dim * L
