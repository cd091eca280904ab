"""Model-shape presets and shape arithmetic for synthetic checkpoints (product side; no oracle dependency).

The reference reads its shapes from the checkpoint's config.json at run time (llama3.2_model_numpy.py:1016-1021);
with no network these presets restate the public HF configs of the models BASELINE.json names (SURVEY.md App. A;
Gemma-2-2B values are the ones quoted in gemma2_model.py:33-109).  tests/test_cabi.py checks them against the oracle's
independent copy.
"""
from __future__ import annotations

from typing import List

from .model import AttributeDict


def llama_config(name: str = "1b") -> AttributeDict:
    c = dict(family="llama", model_type="llama", vocab_size=128256, hidden_size=2048, intermediate_size=8192,
             num_hidden_layers=16, num_attention_heads=32, num_key_value_heads=8, head_dim=64, max_position_embeddings=131072,
             rope_theta=500000.0, rms_norm_eps=1e-5, hidden_act="silu", attention_dropout=0.0, use_cache=False)
    if name == "3b":
        c.update(hidden_size=3072, num_hidden_layers=28, num_attention_heads=24, num_key_value_heads=8, head_dim=128)
    elif name == "tiny":
        c.update(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=2, head_dim=64)
    elif name != "1b":
        raise KeyError(name)
    return AttributeDict(c)


def gemma2_config(name: str = "2b") -> AttributeDict:
    c = dict(family="gemma2", model_type="gemma2", vocab_size=256000, hidden_size=2304, intermediate_size=9216,
             num_hidden_layers=26, num_attention_heads=8, num_key_value_heads=4, head_dim=256, max_position_embeddings=8192,
             rope_theta=10000.0, rms_norm_eps=1e-6, hidden_act="gelu_pytorch_tanh", attention_dropout=0.0, use_cache=False,
             pad_token_id=0, query_pre_attn_scalar=256, final_logit_softcapping=30.0, attn_logit_softcapping=50.0,
             sliding_window=4096)
    if name == "tiny":
        c.update(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                 num_key_value_heads=1, head_dim=256)
    elif name != "2b":
        raise KeyError(name)
    return AttributeDict(c)


def weight_names(cfg) -> List[str]:
    """HF tensor names of a checkpoint (llama3.2_model_numpy.py:164-168, :250-252, :377-380, :597;
    gemma2_model.py:590-591 adds the two feed-forward norms)."""
    names = ["model.embed_tokens.weight"]
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        names += [p + "input_layernorm.weight", p + "self_attn.q_proj.weight", p + "self_attn.k_proj.weight",
                  p + "self_attn.v_proj.weight", p + "self_attn.o_proj.weight", p + "post_attention_layernorm.weight"]
        if cfg.family == "gemma2":
            names += [p + "pre_feedforward_layernorm.weight", p + "post_feedforward_layernorm.weight"]
        names += [p + "mlp.gate_proj.weight", p + "mlp.up_proj.weight", p + "mlp.down_proj.weight"]
    names.append("model.norm.weight")
    return names


def weight_shape(cfg, name: str):
    H, I = cfg.hidden_size, cfg.intermediate_size
    q, kv = cfg.num_attention_heads * cfg.head_dim, cfg.num_key_value_heads * cfg.head_dim
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


def matmul_params(cfg) -> int:
    """Weights that take part in a GEMM per token, tied lm_head included (P_mm)."""
    H, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    q, kv = cfg.num_attention_heads * cfg.head_dim, cfg.num_key_value_heads * cfg.head_dim
    return L * (H * q + 2 * H * kv + q * H + 3 * H * I) + cfg.vocab_size * H


def flops_per_token(cfg, S: int) -> float:
    """Algorithmic FLOPs per token at sequence length S: 2*P_mm + causal attention 2*(S+1)*NH*HD*L."""
    return 2.0 * matmul_params(cfg) + 2.0 * (S + 1) * cfg.num_attention_heads * cfg.head_dim * cfg.num_hidden_layers


def synth_checkpoint(cfg, seed: int = 0, device="cpu", std: float = 0.02):
    """Random-init bf16 checkpoint of the named architecture (there is no network for real weights): N(0, std) matrices,
    unit norm gains (Gemma stores gamma - 1 = 0).  Returns the reference's {hf_name: torch.Tensor} dict on the host."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for n in weight_names(cfg):
        shp = weight_shape(cfg, n)
        if len(shp) == 1:
            out[n] = torch.full(shp, 0.0 if cfg.family == "gemma2" else 1.0, dtype=torch.bfloat16)
        else:
            out[n] = (torch.randn(shp, generator=g, device=device, dtype=torch.float32) * std).to(torch.bfloat16).cpu()
    return out
