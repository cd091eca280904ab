"""CPU: the C-ABI library builds, loads, exports every symbol include/b200fwd.h declares, and fails
loudly (no CPU fallback) when there is no sm_100 device.  No compute calls."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from llm_np_cp_b200 import _lib


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200fwd.h")).read()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == {n for n, _, _ in _lib.SYMBOLS}, declared ^ {n for n, _, _ in _lib.SYMBOLS}


def test_config_struct_matches_header():
    hdr = open(os.path.join(ROOT, "include", "b200fwd.h")).read()
    body = hdr[hdr.index("typedef struct b2_config {"):hdr.index("} b2_config;")]
    fields = re.findall(r"\b([a-z_]+)\s*[,;]", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert fields == [f for f, _ in _lib.B2Config._fields_]


@pytest.mark.parametrize("struct,cls", [("b2_forward_args", "B2ForwardArgs"), ("b2_stats", "B2Stats")])
def test_other_structs_match_header(struct, cls):
    """Field order of every struct that crosses the ABI: header declaration == ctypes mirror."""
    hdr = open(os.path.join(ROOT, "include", "b200fwd.h")).read()
    body = hdr[hdr.index(f"typedef struct {struct} {{"):hdr.index(f"}} {struct};")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b\*?\s*([a-zA-Z_][a-zA-Z0-9_]*)\s*[,;]", body)
    assert fields == [f for f, _ in getattr(_lib, cls)._fields_]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.lib()
    assert lib.b2_device_count() == 0
    cfg = _lib.B2Config(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=1,
                        num_attention_heads=4, num_key_value_heads=2, head_dim=64, max_positions=64,
                        rope_theta=10000.0, rms_norm_eps=1e-5)
    h = C.c_void_p()
    rc = lib.b2_engine_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -4 and b"no CPU fallback" in lib.b2_last_error()
    import llm_np_cp_b200 as M
    from oracle import np_oracle as O
    with pytest.raises(_lib.B2Error):
        M.LlamaForCausalLM(M.AttributeDict(O.llama_config("tiny")), {})


def test_interleave_gate_up_host_helper():
    import numpy as np
    I, H = 64, 8
    g = np.arange(I * H, dtype=np.uint16).reshape(I, H)
    u = g + 10000
    out = np.zeros((2 * I, H), dtype=np.uint16)
    rc = _lib.lib().b2_interleave_gate_up(g.ctypes.data, u.ctypes.data, out.ctypes.data, I, H)
    assert rc == 0
    assert (out[:32] == g[:32]).all() and (out[32:64] == u[:32]).all() and (out[64:96] == g[32:]).all()


def test_shape_presets_match_the_oracle_copy():
    """llm_np_cp_b200.shapes (what bench.py's GPU arm uses) restates the same public configs, tensor names / shapes and
    FLOP arithmetic as the oracle's independent copy."""
    from llm_np_cp_b200 import shapes
    from oracle import np_oracle as O
    for mine, ref in [(shapes.llama_config("1b"), O.llama_config("1b")), (shapes.llama_config("3b"), O.llama_config("3b")),
                      (shapes.llama_config("tiny"), O.llama_config("tiny")), (shapes.gemma2_config("2b"), O.gemma2_config("2b")),
                      (shapes.gemma2_config("tiny"), O.gemma2_config("tiny"))]:
        for k, v in ref.items():
            assert mine[k] == v, k
        assert shapes.weight_names(mine) == O.weight_names(ref)
        assert all(shapes.weight_shape(mine, n) == O.weight_shape(ref, n) for n in O.weight_names(ref))
        for S in (16, 2048, 8192):
            assert shapes.flops_per_token(mine, S) == O.flops_per_token(ref, S)
