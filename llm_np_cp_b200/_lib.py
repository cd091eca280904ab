"""ctypes binding of libb200fwd.so (include/b200fwd.h).  No CPU fallback: if the library cannot be
loaded, or no sm_100 device is present when an engine is created, this raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200fwd.so")


class B2Config(C.Structure):
    _fields_ = [("vocab_size", C.c_int32), ("hidden_size", C.c_int32), ("intermediate_size", C.c_int32),
                ("num_hidden_layers", C.c_int32), ("num_attention_heads", C.c_int32),
                ("num_key_value_heads", C.c_int32), ("head_dim", C.c_int32), ("max_positions", C.c_int32),
                ("rope_theta", C.c_float), ("rms_norm_eps", C.c_float), ("hidden_act", C.c_int32),
                ("gemma", C.c_int32), ("final_logit_softcap", C.c_float), ("attn_logit_softcap", C.c_float),
                ("sliding_window", C.c_int32), ("attn_scale", C.c_float), ("always_causal", C.c_int32),
                ("precise", C.c_int32), ("rope_factor", C.c_float), ("rope_low_freq_factor", C.c_float),
                ("rope_high_freq_factor", C.c_float), ("rope_original_max_positions", C.c_int32),
                ("tp_rank", C.c_int32), ("tp_world", C.c_int32)]


class B2Stats(C.Structure):
    _fields_ = [("launches", C.c_int64), ("gemm_launches", C.c_int64), ("gemm_ms", C.c_double),
                ("gemm_flops", C.c_double), ("attn_ms", C.c_double), ("attn_flops", C.c_double),
                ("attn_launches", C.c_int64), ("timed_forwards", C.c_int64), ("comm_calls", C.c_int64),
                ("comm_bytes", C.c_double), ("comm_ms", C.c_double), ("p2p_fused", C.c_int64)]


class B2ForwardArgs(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("inputs_embeds", C.c_void_p), ("position_ids", C.c_void_p),
                ("B", C.c_int32), ("S", C.c_int32), ("logits_mode", C.c_int32),
                ("logits", C.c_void_p), ("hidden", C.c_void_p), ("attn_probs", C.c_void_p), ("stream", C.c_void_p)]


# every symbol include/b200fwd.h declares: (name, restype, argtypes)
_P, _I, _F = C.c_void_p, C.c_int, C.c_float
SYMBOLS = [
    ("b2_last_error", C.c_char_p, []),
    ("b2_device_count", _I, []),
    ("b2_engine_create", _I, [C.POINTER(B2Config), _I, C.POINTER(_P)]),
    ("b2_engine_destroy", None, [_P]),
    ("b2_load_weight", _I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I, _I]),
    ("b2_finalize", _I, [_P]),
    ("b2_cache_create", _I, [_P, _I, _I, C.POINTER(_P)]),
    ("b2_cache_create_paged", _I, [_P, _I, _I, _I, _I, C.POINTER(_P)]),
    ("b2_cache_set_block_table", _I, [_P, C.POINTER(C.c_int32), _I]),
    ("b2_cache_info", _I, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("b2_cache_destroy", None, [_P]),
    ("b2_cache_len", _I, [_P]),
    ("b2_cache_reset", _I, [_P]),
    ("b2_cache_read", _I, [_P, _I, _I, _I, _P]),
    ("b2_forward", _I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    ("b2_forward_ex", _I, [_P, _P, C.POINTER(B2ForwardArgs)]),
    ("b2_set_layer_types", _I, [_P, C.POINTER(C.c_int32), _I]),
    ("b2_tp_unique_id", _I, [_P]),
    ("b2_tp_init", _I, [_P, _P]),
    ("b2_tp_ipc_handle", _I, [_P, _P]),
    ("b2_tp_set_peers", _I, [_P, _P, _I]),
    ("b2_forward_host", _I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    ("b2_set_decode_graph", _I, [_P, _I]),
    ("b2_sample_min_p", _I, [_P, _P, _I, _I, _F, _P, _P, _P, _P]),
    ("b2_argmax", _I, [_P, _P, _I, _I, _P, _P]),
    ("b2_set_timing", _I, [_P, _I]),
    ("b2_get_stats", _I, [_P, C.POINTER(B2Stats)]),
    ("b2_reset_stats", _I, [_P]),
    ("b2_op_rmsnorm", _I, [_P, _P, _P, _P, _I, _I, _F, _P]),
    ("b2_op_add_rmsnorm", _I, [_P, _P, _P, _I, _I, _F, _P]),
    ("b2_op_gemm", _I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P]),
    ("b2_op_gemm_sk", _I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    ("b2_op_qkv_rope", _I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P]),
    ("b2_op_attention", _I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _I, _I, _I, _P]),
    ("b2_interleave_gate_up", _I, [_P, _P, _P, _I, _I]),
]

_lib = None


class B2Error(RuntimeError):
    pass


def lib():
    """Load (building first if the .so is missing or stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        _build.build()
    except Exception as e:  # no nvcc on this box: fall through to the prebuilt .so
        if not os.path.exists(LIB_PATH):
            raise B2Error(f"libb200fwd.so is missing and could not be built: {e}") from e
    try:
        l = C.CDLL(LIB_PATH)
    except OSError as e:
        raise B2Error(f"cannot load {LIB_PATH}: {e} -- there is no CPU fallback") from e
    for name, res, args in SYMBOLS:
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(rc: int):
    if rc != 0:
        msg = lib().b2_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)      # the reference raises ValueError for bad shapes / broadcast errors
        raise B2Error(f"b200fwd error {rc}: {msg}")
