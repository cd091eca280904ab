"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference modules in this container.

/root/reference is read-only and absent on the GPU box.  Here (container) the reference files are exec'd from
where they lie -- used by ``oracle/make_golden*.py`` to pin the restatement in ``np_oracle.py`` (SURVEY.md section
8c).  On the GPU box the same modules are loaded from the byte-compiled artefacts ``oracle/_ref/*.pyc`` that
``oracle/build_ref.py`` produced here (a build output, git-ignored; no reference source is copied), so that
bench.py's CPU legs can time the reference's own classes.  Import-time side effects are neutralised:

* ``huggingface_hub.notebook_login`` is stubbed (llama3.2_model_numpy.py:972-973 calls it);
* ``gemma2_model.py`` is CuPy-only, so a fake ``cupy`` module (NumPy namespace + a RawKernel
  that performs the max-subtracted row softmax its CUDA string implements, gemma2_model.py:
  987-1017) is registered before import, and ``load_weights`` is replaced by a CPU version that
  keeps the lm_head -> embed_tokens alias (gemma2_model.py:1134-1138).
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

REF_DIR = "/root/reference"
BUILT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    """Can the reference's Llama NumPy module be loaded here (source mounted, or oracle/_ref built)?"""
    return os.path.exists(f"{REF_DIR}/llama3.2_model_numpy.py") or os.path.exists(f"{BUILT_DIR}/llama3_2_model_numpy.pyc")


def _exec_reference(modname: str, src: str, built: str):
    """Module object of reference file `src`: from /root/reference if mounted, else from its byte-compiled build."""
    path = f"{REF_DIR}/{src}"
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location(modname, path)
    else:
        pyc = f"{BUILT_DIR}/{built}"
        if not os.path.exists(pyc):
            raise FileNotFoundError(f"neither {path} nor {pyc} exists (run oracle/build_ref.py where /root/reference is mounted)")
        loader = importlib.machinery.SourcelessFileLoader(modname, pyc)
        spec = importlib.util.spec_from_loader(modname, loader)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stub_login():
    import huggingface_hub
    huggingface_hub.notebook_login = lambda *a, **k: None


def load_llama_numpy():
    """-> the module object of /root/reference/llama3.2_model_numpy.py."""
    _stub_login()
    return _exec_reference("ref_llama_np", "llama3.2_model_numpy.py", "llama3_2_model_numpy.pyc")


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
        mod = _exec_reference("ref_gemma2", "gemma2_model.py", "gemma2_model.pyc")
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
