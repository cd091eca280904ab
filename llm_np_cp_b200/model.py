"""Host-side mirror of the reference's Python call surface, driving libb200fwd.so.

Reference surface:
  tokenizer, model, config = load_model(name)                      llama3.2_model_numpy.py:1030-1047
  loss, logits, kv_cache, hidden_states, attentions = model(ids, use_cache=..., kv_cache=...)   :777-830
  text = generate(prompt, tokenizer, model, max_tokens, kv_cache=KVCache(), config=config)     :873-910

Same names, argument meaning and error behaviour; arrays that the reference's CuPy twin keeps on the
GPU (logits, hidden states, cache contents) are returned as torch CUDA tensors.  Deviations, all
forced by never materialising [B,NH,S,T] scores: ``attentions`` is None.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import B2Config, B2ForwardArgs, B2Stats, check

LOGITS_NONE, LOGITS_ALL, LOGITS_LAST, LOGITS_ALL_LAST_TO_HOST = 0, 1, 2, 3


class AttributeDict(dict):
    """dict with attribute access (llama3.2_model_numpy.py:212-215)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


# the reference's module-global injection seam (llama3.2_model_numpy.py:1023-1028)
weights: Dict[str, torch.Tensor] = dict()


def load_weights(key):
    """Name -> tensor lookup with the lm_head alias (llama3.2_model_numpy.py:1024-1028)."""
    if key == "lm_head.weight":
        key = "model.embed_tokens.weight"
    return weights[key]


_DTYPES = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}


def _cfg_get(config, key, default=None):
    """Optional config field: dict-style configs (AttributeDict raises KeyError from __getattr__, so getattr's default
    would never apply) and attribute-style ones (transformers configs) alike."""
    if isinstance(config, dict):
        return config.get(key, default)
    try:
        return getattr(config, key)
    except (AttributeError, KeyError):
        return default


class _LayerViews:
    """List-like view of the cache: ``cache.key_cache[i]`` copies only layer i off the engine's buffer (the
    reference's ``key_cache`` is a plain list of per-layer arrays, llama3.2_model_numpy.py:313-314)."""

    def __init__(self, cache, which):
        self._c, self._which = cache, which

    def __len__(self):
        c = self._c
        return 0 if c._h is None or c.num_items() == 0 else c._model._cfg.num_hidden_layers

    def __getitem__(self, i):
        n = len(self)
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(n))]
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("list index out of range")
        return self._c._read_layer(self._which, i)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class KVCache:
    """KVCache() of the reference (llama3.2_model_numpy.py:311-340) backed by device memory owned by the engine.

    ``KVCache()`` takes no arguments there.  Here, with no ``max_tokens`` the cache is PAGED (``page_size`` positions per
    page, block table per sequence) and its pool grows on demand, so memory follows the tokens actually cached -- the
    reference's lists grow the same way; with ``max_tokens`` it is one dense preallocated ``[L][2][B][NKV][max_tokens][HD]``
    buffer (what bench.py uses), or a paged pool of ``n_pages`` pages when ``paged=True``."""

    def __init__(self, max_batch: Optional[int] = None, max_tokens: Optional[int] = None, paged: Optional[bool] = None,
                 page_size: int = 128, n_pages: Optional[int] = None):
        self.max_batch, self.max_tokens = max_batch, max_tokens
        self.paged = (max_tokens is None) if paged is None else bool(paged)
        self.page_size, self.n_pages = page_size, n_pages
        self._table = None
        self._h = None
        self._model = None

    def set_block_table(self, table):
        """Place pages yourself: int32 [max_batch, ceil(max_tokens / page_size)] of distinct physical pages.  Applied when
        the cache binds to a model (first use); only valid for ``paged=True`` caches with explicit sizes."""
        self._table = np.ascontiguousarray(table, dtype=np.int32)

    def pool_info(self):
        """(page_size, pages in the pool, pages handed out); page_size 0 = dense cache."""
        if self._h is None:
            return (self.page_size if self.paged else 0, 0, 0)
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        check(_lib.lib().b2_cache_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def _bind(self, model, batch):
        if self._h is None:
            mb = self.max_batch or batch
            mt = self.max_tokens or min(model._cfg.max_positions, 8192)
            h = C.c_void_p()
            if self.paged:
                # no explicit pool size: start with one page per sequence and let the engine double the pool as needed
                npg = self.n_pages if self.n_pages is not None else (mb if self.max_tokens is None else 0)
                check(_lib.lib().b2_cache_create_paged(model._h, mb, mt, self.page_size, npg, C.byref(h)))
                if self._table is not None:
                    t = self._table.reshape(-1)
                    check(_lib.lib().b2_cache_set_block_table(h, t.ctypes.data_as(C.POINTER(C.c_int32)), t.size))
            else:
                check(_lib.lib().b2_cache_create(model._h, mb, mt, C.byref(h)))
            self._h, self._model, self.max_batch, self.max_tokens = h, model, mb, mt
        elif self._model is not model:
            raise ValueError("KVCache is bound to a different model")
        return self._h

    def num_items(self) -> int:
        return 0 if self._h is None else int(_lib.lib().b2_cache_len(self._h))

    def reset(self):
        if self._h is not None:
            check(_lib.lib().b2_cache_reset(self._h))

    def _read_layer(self, which, li, batch=None):
        m = self._model
        B = batch or self.max_batch
        t = torch.empty((B, m._cfg.num_key_value_heads, self.num_items(), m._cfg.head_dim),
                        dtype=torch.bfloat16, device=m.device)
        check(_lib.lib().b2_cache_read(self._h, li, which, B, t.data_ptr()))
        return t

    @property
    def key_cache(self):
        """list-like of [B,NKV,T,HD] post-RoPE keys, one per layer (bf16 device tensors), read lazily per layer."""
        return _LayerViews(self, 0)

    @property
    def value_cache(self):
        return _LayerViews(self, 1)

    def __del__(self):
        try:
            if self._h is not None:
                _lib.lib().b2_cache_destroy(self._h)
        except Exception:
            pass


class _CausalLM:
    """Common engine wrapper; subclasses set the family."""
    family = "llama"

    def __init__(self, config, weights_dict: Optional[Dict[str, torch.Tensor]] = None, device: int = 0,
                 max_positions: Optional[int] = None, *, attn_logit_softcapping=None, sliding_window=None,
                 use_query_pre_attn_scalar: bool = False, always_causal: bool = False, precise: bool = False,
                 rope_scaling: Optional[dict] = None, tensor_parallel: bool = False, tp_group=None):
        self.config = config
        L = _lib.lib()
        if L.b2_device_count() == 0:
            raise _lib.B2Error("no sm_100 CUDA device: llm_np_cp_b200 has no CPU fallback")
        gem = self.family == "gemma2"
        c = B2Config()
        c.vocab_size = int(config.vocab_size)
        c.hidden_size = int(config.hidden_size)
        c.intermediate_size = int(config.intermediate_size)
        c.num_hidden_layers = int(config.num_hidden_layers)
        c.num_attention_heads = int(config.num_attention_heads)
        c.num_key_value_heads = int(config.num_key_value_heads)
        c.head_dim = int(config.head_dim)                      # must exist in the config (App. B Q7)
        mp = int(_cfg_get(config, "max_position_embeddings", 8192))
        c.max_positions = int(max_positions or min(mp, 16384))
        c.rope_theta = float(config.rope_theta)
        c.rms_norm_eps = float(config.rms_norm_eps)
        act = _cfg_get(config, "hidden_act", None) or _cfg_get(config, "hidden_activation", "silu")
        if act not in ("silu", "gelu_pytorch_tanh"):
            raise KeyError(act)                                # ACT2FN_np lookup fails the same way (:111-116)
        c.hidden_act = 0 if act == "silu" else 1
        c.gemma = 1 if gem else 0
        cap = _cfg_get(config, "final_logit_softcapping", None) if gem else None
        c.final_logit_softcap = float(cap) if cap else 0.0
        # options beyond the reference (off by default = reference behaviour)
        c.attn_logit_softcap = float(attn_logit_softcapping or 0.0)
        c.sliding_window = int(sliding_window or 0)
        c.attn_scale = float(_cfg_get(config, "query_pre_attn_scalar", c.head_dim)) ** -0.5 if use_query_pre_attn_scalar else 0.0
        c.always_causal = 1 if always_causal else 0
        # parity mode: every bf16 activation is carried as hi + lo (3 tensor-core products per attention product,
        # 2x K in the weight GEMMs); meets 1e-2 max-abs vs the reference.  Never quote throughput from it.
        c.precise = 1 if precise else 0
        # HF "llama3" rope frequency scaling (config.json: rope_scaling = {rope_type: llama3, factor, low_freq_factor,
        # high_freq_factor, original_max_position_embeddings}); the reference ignores it, so it is opt-in here
        if rope_scaling:
            kind = rope_scaling.get("rope_type", rope_scaling.get("type", "llama3"))
            if kind != "llama3":
                raise ValueError(f"rope_scaling type {kind!r} is not supported (llama3 only)")
            c.rope_factor = float(rope_scaling["factor"])
            c.rope_low_freq_factor = float(rope_scaling.get("low_freq_factor", 1.0))
            c.rope_high_freq_factor = float(rope_scaling.get("high_freq_factor", 4.0))
            c.rope_original_max_positions = int(rope_scaling.get("original_max_position_embeddings", 8192))
        # optional tensor parallelism (BASELINE.json north_star; the reference has none): this process is one rank of an
        # initialised torch.distributed group, one process per GPU; heads / FFN columns / vocabulary rows are split, the
        # partial sums of o_proj and down_proj are all-reduced with NCCL inside the engine
        self.tp_rank, self.tp_world = 0, 1
        if tensor_parallel:
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("tensor_parallel=True needs an initialised torch.distributed process group (torchrun)")
            self.tp_rank, self.tp_world = dist.get_rank(tp_group), dist.get_world_size(tp_group)
            c.tp_rank, c.tp_world = self.tp_rank, self.tp_world
        self._cfg = c
        self.device = torch.device("cuda", device)
        self.vocab_size = c.vocab_size
        h = C.c_void_p()
        check(L.b2_engine_create(C.byref(c), device, C.byref(h)))
        self._h = h
        if self.tp_world > 1:
            import torch.distributed as dist
            uid = (C.c_uint8 * 128)()
            if self.tp_rank == 0:
                check(L.b2_tp_unique_id(uid))
            box = [bytes(uid)]
            src = dist.get_global_rank(tp_group, 0) if tp_group is not None else 0
            dist.broadcast_object_list(box, src=src, group=tp_group, device=self.device)
            check(L.b2_tp_init(self._h, box[0]))      # ncclCommInitRank: collective over the group
            # Peer-memory all-reduce fused into the norm kernel for the decode step (NVLink P2P, cudaIpc).  Measured on B200
            # (decode, B=64): 1.32 vs 1.51 ms per token step at TP=2, 1.50 vs 1.56 at TP=4; at TP=8 every rank pulls seven
            # peers' planes and NCCL's in-switch (NVLS) reduction wins (1.77 ms), so the default stops at 4 ranks.
            # B2_TP_P2P=1 forces it on, B2_TP_NO_P2P=1 off.
            use_p2p = (self.tp_world <= 4 or os.environ.get("B2_TP_P2P")) and not os.environ.get("B2_TP_NO_P2P")
            if use_p2p:
                hbuf = (C.c_uint8 * 64)()
                check(L.b2_tp_ipc_handle(self._h, hbuf))
                gathered = [None] * self.tp_world
                dist.all_gather_object(gathered, bytes(hbuf), group=tp_group)
                check(L.b2_tp_set_peers(self._h, b"".join(gathered), self.tp_world))
        self._graph_on = False
        self._step_logits, self._step_ids = {}, {}
        src = weights_dict if weights_dict is not None else weights
        self._load(src)
        check(L.b2_finalize(self._h))

    def _load(self, src):
        L = _lib.lib()
        for name, t in src.items():
            if name == "lm_head.weight":
                continue                                       # tied to the embedding (:1025-1026)
            if isinstance(t, np.ndarray):
                t = torch.from_numpy(t)
            t = t.detach().cpu().contiguous()
            if t.dtype not in _DTYPES:
                t = t.to(torch.float32)
            shape = (C.c_int64 * t.dim())(*t.shape)
            check(L.b2_load_weight(self._h, name.encode(), t.data_ptr(), shape, t.dim(), _DTYPES[t.dtype]))

    # ---- the reference's __call__ (llama3.2_model_numpy.py:777-830 / gemma2_model.py:830-886)
    # opt-in: also materialise the reference's `attentions` tuple (L tensors [B,NH,S,T], llama3.2_model_numpy.py:713-714)
    # through a separate debug kernel.  Off by default: the fused attention never forms the score matrix, and at
    # Llama-1B S=2048 the tuple alone is 8.6 GB.
    materialize_attentions = False

    def __call__(self, input_ids=None, attention_mask=None, position_ids=None, inputs_embeds=None, labels=None,
                 use_cache=None, output_attentions=True, output_hidden_states=True, return_dict=True, kv_cache=None,
                 logits_mode: int = LOGITS_ALL, logits_out=None):
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        use_cache = use_cache if use_cache is not None else _cfg_get(self.config, "use_cache", False)
        H, V, Lc = self._cfg.hidden_size, self._cfg.vocab_size, self._cfg.num_hidden_layers
        ids = emb = None
        if input_ids is not None:
            ids = input_ids
            if isinstance(ids, np.ndarray):
                ids = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64))
            ids = torch.as_tensor(ids, dtype=torch.int64)
            if ids.dim() != 2:
                raise ValueError("input_ids must be [batch, seq]")
            if ids.device.type != "cuda":
                if int(ids.min()) < 0 or int(ids.max()) >= self.vocab_size:
                    raise IndexError("token id out of range")      # NumPy fancy indexing raises the same (:675)
                ids = ids.to(self.device, non_blocking=True)
            ids = ids.contiguous()
            B, S = ids.shape
        else:                                                      # caller-provided embeddings (:645-652, :675-676)
            emb = inputs_embeds
            if isinstance(emb, np.ndarray):
                emb = torch.from_numpy(np.ascontiguousarray(emb, dtype=np.float32))
            emb = torch.as_tensor(emb, dtype=torch.float32).to(self.device).contiguous()
            if emb.dim() != 3 or emb.shape[-1] != H:
                raise ValueError(f"inputs_embeds must be [batch, seq, {H}]")
            B, S = emb.shape[:2]
        # The reference appends to kv_cache whenever one is passed; use_cache only decides whether RoPE positions
        # continue from the cache length (:659-672).  Caller position_ids win over both (:657).
        cache_h = kv_cache._bind(self, B) if kv_cache is not None else None
        pos = None
        if position_ids is not None:
            pos = torch.as_tensor(np.asarray(position_ids.cpu() if torch.is_tensor(position_ids) else position_ids)).reshape(-1)
            if pos.numel() != S:
                raise ValueError("position_ids must be [1, seq] (shared by the batch, like the reference's)")
        elif kv_cache is not None and not use_cache and kv_cache.num_items() > 0:
            pos = torch.arange(0, S)
        if pos is not None:
            if int(pos.min()) < 0 or int(pos.max()) >= self._cfg.max_positions:
                raise IndexError("position id outside the engine's RoPE table")
            pos = pos.to(torch.int32).to(self.device)
        graph_step = (self._graph_on and S == 1 and cache_h is not None and ids is not None and pos is None
                      and not output_hidden_states and logits_mode != LOGITS_NONE)
        logits = hidden = probs = None
        if logits_mode != LOGITS_NONE:
            shape = (B, S, V) if logits_mode == LOGITS_ALL else (B, V)
            if logits_out is not None:           # caller-owned buffer (stable pointer: lets decode steps replay one CUDA graph)
                if tuple(logits_out.shape) != shape or logits_out.dtype != torch.float32 or not logits_out.is_contiguous():
                    raise ValueError(f"logits_out must be a contiguous float32 tensor of shape {shape}")
                logits = logits_out
            elif graph_step:                     # engine-side graphs are keyed on pointers: reuse one buffer per (cache, B, mode)
                key = (id(kv_cache), B, logits_mode)
                logits = self._step_logits.get(key)
                if logits is None:
                    if len(self._step_logits) >= 8:
                        self._step_logits.clear(); self._step_ids.clear()
                    logits = self._step_logits[key] = torch.empty(shape, dtype=torch.float32, device=self.device)
            else:
                logits = torch.empty(shape, dtype=torch.float32, device=self.device)
        if graph_step:                           # same for the one-token id buffer
            key = (id(kv_cache), B)
            buf = self._step_ids.get(key)
            if buf is None:
                buf = self._step_ids[key] = torch.empty((B, 1), dtype=torch.int64, device=self.device)
            if buf.data_ptr() != ids.data_ptr():
                buf.copy_(ids)
                ids = buf
        if output_hidden_states:
            hidden = torch.empty((Lc + 1, B, S, H), dtype=torch.float32, device=self.device)
        T = (kv_cache.num_items() if kv_cache is not None else 0) + S
        if output_attentions and self.materialize_attentions:
            if self.tp_world > 1:
                raise NotImplementedError("materialize_attentions is a single-GPU debug path")
            probs = torch.empty((Lc, B, self._cfg.num_attention_heads, S, T), dtype=torch.float32, device=self.device)
        a = B2ForwardArgs()
        a.ids = ids.data_ptr() if ids is not None else None
        a.inputs_embeds = emb.data_ptr() if emb is not None else None
        a.position_ids = pos.data_ptr() if pos is not None else None
        a.B, a.S, a.logits_mode = B, S, logits_mode
        a.logits = logits.data_ptr() if logits is not None else None
        a.hidden = hidden.data_ptr() if hidden is not None else None
        a.attn_probs = probs.data_ptr() if probs is not None else None
        a.stream = torch.cuda.current_stream(self.device).cuda_stream
        check(_lib.lib().b2_forward_ex(self._h, cache_h, C.byref(a)))
        self._keep = (ids, emb, pos)                           # keep inputs alive until the stream is done
        if graph_step and logits_out is None:
            logits = logits.clone()                            # the step buffer is overwritten by the next replay
        hs = tuple(hidden[i] for i in range(Lc + 1)) if hidden is not None else None
        att = tuple(probs[i] for i in range(Lc)) if probs is not None else None
        return (None, logits, kv_cache, hs, att)

    def argmax(self, logits, out=None):
        """Greedy pick on the device: [B,V] fp32 logits -> [B] int64 (np.argmax semantics)."""
        lg = logits if logits.dim() == 2 else logits.reshape(-1, logits.shape[-1])
        B, V = lg.shape
        if out is None:
            out = torch.empty(B, dtype=torch.int64, device=self.device)
        check(_lib.lib().b2_argmax(self._h, lg.data_ptr(), B, lg.stride(0), out.data_ptr(),
                                   torch.cuda.current_stream(self.device).cuda_stream))
        return out

    # ---- instrumentation
    def set_timing(self, on):
        """True/1: CUDA events around every GEMM / attention launch; k > 1: only on every k-th forward; False/0: off."""
        check(_lib.lib().b2_set_timing(self._h, int(on)))

    def set_decode_graph(self, on: bool):
        self._graph_on = bool(on)
        check(_lib.lib().b2_set_decode_graph(self._h, 1 if on else 0))

    def set_layer_types(self, layer_types):
        """HF ``config.layer_types`` ("sliding_attention" / "full_attention" per layer, or booleans): which layers
        apply the sliding window.  Without this call the engine uses transformers' Gemma-2 default (even layers)."""
        flags = [1 if (t == "sliding_attention" or t is True or t == 1) else 0 for t in layer_types]
        arr = (C.c_int32 * len(flags))(*flags)
        check(_lib.lib().b2_set_layer_types(self._h, arr, len(flags)))

    def reset_stats(self):
        check(_lib.lib().b2_reset_stats(self._h))

    def stats(self) -> dict:
        s = B2Stats()
        check(_lib.lib().b2_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in B2Stats._fields_}

    def forward_host(self, ids_np: np.ndarray, kv_cache: Optional[KVCache] = None, logits_mode: int = LOGITS_LAST,
                     out: Optional[np.ndarray] = None):
        """The C-ABI call with HOST buffers (b2_forward_host): ids H2D, forward, logits D2H."""
        ids_np = np.ascontiguousarray(ids_np, dtype=np.int64)
        B, S = ids_np.shape
        V = self._cfg.vocab_size
        shape = (B, S, V) if logits_mode == LOGITS_ALL else (B, V)
        if out is None:
            out = np.empty(shape, dtype=np.float32)
        cache_h = kv_cache._bind(self, B) if kv_cache is not None else None
        check(_lib.lib().b2_forward_host(self._h, cache_h, ids_np.ctypes.data, B, S, logits_mode, out.ctypes.data, None))
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.lib().b2_engine_destroy(self._h)
                self._h = None
        except Exception:
            pass


class LlamaForCausalLM(_CausalLM):
    """LlamaForCausalLM_np (llama3.2_model_numpy.py:734-830)."""
    family = "llama"


class Gemma2ForCausalLM(_CausalLM):
    """Gemma2ForCausalLM_np (gemma2_model.py:787-886).  Reference-parity flags by default: no attention
    soft-cap, no sliding window, 1/sqrt(head_dim) scaling.  ``hf_flags=True``
    switches on the HF semantics BASELINE.json asks for (soft-cap, sliding window on even layers)."""
    family = "gemma2"

    def __init__(self, config, weights_dict=None, device: int = 0, max_positions=None, hf_flags: bool = False, **kw):
        if hf_flags:
            kw.setdefault("attn_logit_softcapping", _cfg_get(config, "attn_logit_softcapping", 50.0))
            kw.setdefault("sliding_window", _cfg_get(config, "sliding_window", 4096))
            kw.setdefault("use_query_pre_attn_scalar", True)
            kw.setdefault("always_causal", True)
        super().__init__(config, weights_dict, device, max_positions, **kw)
        lt = _cfg_get(config, "layer_types", None)
        if hf_flags and lt:
            self.set_layer_types(lt)


# the reference's class names, for drop-in imports
LlamaForCausalLM_np = LlamaForCausalLM
Gemma2ForCausalLM_np = Gemma2ForCausalLM


def min_p_sampling(logits, p_base: float = 0.1, generator: Optional[torch.Generator] = None, model=None,
                   return_probs: bool = False):
    """min_p_sampling (llama3.2_model_numpy.py:948-961): softmax -> keep p >= p_base*p_max -> renormalise
    -> draw.  Runs on the device; the reference's draw is an unseeded torch.multinomial, here it is an
    inverse-CDF lookup of a uniform from ``generator`` (pass a seeded generator for reproducibility)."""
    if model is None:
        raise ValueError("min_p_sampling needs the model (engine handle) that produced the logits")
    lg = torch.as_tensor(logits, dtype=torch.float32, device=model.device)
    one = lg.dim() == 1
    if one:
        lg = lg[None]
    lg = lg.contiguous()
    B, V = lg.shape
    u = torch.rand(B, generator=generator, dtype=torch.float32).to(model.device)
    out = torch.empty(B, dtype=torch.int64, device=model.device)
    probs = torch.empty((B, V), dtype=torch.float32, device=model.device) if return_probs else None
    stream = torch.cuda.current_stream(model.device).cuda_stream
    check(_lib.lib().b2_sample_min_p(model._h, lg.data_ptr(), B, V, float(p_base), u.data_ptr(), out.data_ptr(),
                                     probs.data_ptr() if probs is not None else None, stream))
    res = out[0:1] if one else out
    return (res, probs) if return_probs else res


def generate(prompt, tokenizer, model, max_tokens=20, streamer=None, kv_cache=None, config=None, generator=None):
    """The reference's decode loop, quirks included (llama3.2_model_numpy.py:873-910): no EOS stop, BOS
    only on the first call, and with a cache the text of the last sampled token is re-tokenised and fed
    back."""
    res = prompt
    print(prompt, end="", flush=True)
    for i in range(max_tokens):
        if config.use_cache or kv_cache is not None:
            input_ids = tokenizer(prompt, return_tensors="pt", add_special_tokens=(i == 0)).input_ids.numpy()
        else:
            input_ids = tokenizer(res, return_tensors="pt", add_special_tokens=True).input_ids.numpy()
        input_ids = np.asarray(input_ids)
        _, logits, kv_cache, _, _ = model(input_ids, use_cache=config.use_cache, kv_cache=kv_cache,
                                          output_hidden_states=False, output_attentions=False,
                                          logits_mode=LOGITS_LAST)
        val = min_p_sampling(logits[-1], generator=generator, model=model)
        prompt = tokenizer.decode(val.cpu())
        res += prompt
        print(prompt, end="", flush=True)
    return res


def load_sharded_safetensors_via_weight_map(model_dir, index_filename="model.safetensors.index.json"):
    """Local-directory twin of llama3.2_model_numpy.py:981-1021 (index -> shards, single-file fallback,
    config.json -> AttributeDict)."""
    from safetensors.torch import load_file
    all_tensors = {}
    index_path = os.path.join(model_dir, index_filename)
    if os.path.exists(index_path):
        with open(index_path) as f:
            weight_map = json.load(f).get("weight_map", {})
        if not weight_map:
            raise ValueError("The weight_map is missing or empty in the index file.")
        for shard in sorted(set(weight_map.values())):
            all_tensors.update(load_file(os.path.join(model_dir, shard)))
    else:
        all_tensors.update(load_file(os.path.join(model_dir, "model.safetensors")))
    with open(os.path.join(model_dir, "config.json")) as f:
        config = AttributeDict(json.load(f))
    return all_tensors, config


def load_model(model_name, device: int = 0, hf_faithful: bool = False):
    """load_model (llama3.2_model_numpy.py:1030-1047).  ``model_name`` is a local snapshot directory
    (there is no network here); a hub id is resolved through huggingface_hub when one is reachable.
    ``hf_faithful=True`` additionally applies what transformers does and the reference does not: the checkpoint's
    llama3 ``rope_scaling``, a causal mask for every q_len, Gemma-2's attention soft-cap / sliding window."""
    global weights
    from transformers import AutoTokenizer
    model_dir = model_name
    if not os.path.isdir(model_dir):
        from huggingface_hub import snapshot_download
        model_dir = snapshot_download(repo_id=model_name)
    tokenizer = AutoTokenizer.from_pretrained(model_dir, trust_remote_code=True)
    weights, config = load_sharded_safetensors_via_weight_map(model_dir)
    if "head_dim" not in config:
        config["head_dim"] = config["hidden_size"] // config["num_attention_heads"]
    is_gemma = "gemma" in str(config.get("model_type", "")).lower()
    kw = {}
    if hf_faithful:
        if is_gemma:
            kw["hf_flags"] = True
        else:
            kw["always_causal"] = True
            kw["rope_scaling"] = config.get("rope_scaling") or None
    model = (Gemma2ForCausalLM if is_gemma else LlamaForCausalLM)(config, weights, device=device, **kw)
    return tokenizer, model, config
