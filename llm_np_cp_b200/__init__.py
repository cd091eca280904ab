"""llm_np_cp_b200 -- B200-native drop-in for the forward path of githubpradeep/llm_np_cp.

Python host mirroring the reference's call surface (model(ids) -> 5-tuple, KVCache, generate,
min_p_sampling, load_model) over the C ABI in include/b200fwd.h; every op on the per-layer path is a
hand-written sm_100a kernel in csrc/.  No CPU fallback.
"""
from .model import (AttributeDict, Gemma2ForCausalLM, Gemma2ForCausalLM_np, KVCache, LlamaForCausalLM,
                    LlamaForCausalLM_np, generate, load_model, min_p_sampling, load_weights, weights)
from ._lib import B2Error

__all__ = ["AttributeDict", "KVCache", "LlamaForCausalLM", "LlamaForCausalLM_np", "Gemma2ForCausalLM",
           "Gemma2ForCausalLM_np", "generate", "min_p_sampling", "load_model", "load_weights", "weights", "B2Error"]
