"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference modules in this container.

/root/reference is read-only and absent on the GPU box, so this is used only here, by
``oracle/make_golden.py``, to pin the restatement in ``np_oracle.py``.
Nothing is copied: the reference files are exec'd from where they lie with their import-time
side effects neutralised:

* ``huggingface_hub.notebook_login`` is stubbed (llama3.2_model_numpy.py:972-973 calls it);
* ``gemma2_model.py`` is CuPy-only, so a fake ``cupy`` module (NumPy namespace + a RawKernel
  that performs the max-subtracted row softmax its CUDA string implements, gemma2_model.py:
  987-1017) is registered before import, and ``load_weights`` is replaced by a CPU version that
  keeps the lm_head -> embed_tokens alias (gemma2_model.py:1134-1138).
"""
from __future__ import annotations

import importlib.util
import sys
import types

import numpy as np

REF_DIR = "/root/reference"


def _stub_login():
    import huggingface_hub
    huggingface_hub.notebook_login = lambda *a, **k: None


def load_llama_numpy():
    """-> the module object of /root/reference/llama3.2_model_numpy.py."""
    _stub_login()
    spec = importlib.util.spec_from_file_location("ref_llama_np", f"{REF_DIR}/llama3.2_model_numpy.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _FakeRawKernel:
    """Stands in for cp.RawKernel(softmax source): y = softmax(x, axis) with max subtraction,
    args (x, y, axis_dim, inner_dim, outer_dim) as launched at gemma2_model.py:1032-1036."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, grid, block, args):
        x, y, axis_dim, inner_dim, outer_dim = args
        xv = np.asarray(x, dtype=np.float32).reshape(int(outer_dim), int(axis_dim), int(inner_dim))
        m = xv.max(axis=1, keepdims=True)
        e = np.exp(xv - m)
        y[...] = (e / e.sum(axis=1, keepdims=True)).reshape(y.shape)


def load_gemma2():
    """-> the module object of /root/reference/gemma2_model.py running on a NumPy-backed cupy shim."""
    _stub_login()
    fake = types.ModuleType("cupy")
    fake.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    fake.RawKernel = _FakeRawKernel
    fake.from_dlpack = lambda t: t.numpy()
    fake.asnumpy = np.asarray
    sys.modules["cupy"] = fake
    try:
        spec = importlib.util.spec_from_file_location("ref_gemma2", f"{REF_DIR}/gemma2_model.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.modules.pop("cupy", None)

    def load_weights_cpu(key):
        import torch
        if key == "lm_head.weight":
            key = "model.embed_tokens.weight"
        return mod.weights[key].to(torch.float32).numpy()

    mod.load_weights = load_weights_cpu
    return mod


def build_reference_model(mod, cfg, weights):
    """Inject synthetic weights through the module-global ``weights`` seam
    (llama3.2_model_numpy.py:1023-1028) and construct the reference's own model class."""
    mod.weights = weights
    rcfg = mod.AttributeDict({k: v for k, v in dict(cfg).items()})
    if cfg["family"] == "gemma2":
        return mod.Gemma2ForCausalLM_np(rcfg), rcfg
    return mod.LlamaForCausalLM_np(rcfg), rcfg
