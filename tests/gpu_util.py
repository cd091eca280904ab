"""Thin torch-tensor wrappers over the C-ABI single-kernel entry points (tests only)."""
import ctypes as C
import math

import numpy as np
import torch

from llm_np_cp_b200 import _lib

DEV = "cuda:0"


def L():
    return _lib.lib()


def stream():
    return torch.cuda.current_stream().cuda_stream


def bf16_randn(*shape, std=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(torch.bfloat16).to(DEV)


def op_gemm(a, w, epi, bn=256, split_k=1, softcap=0.0, act=0, out=None):
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        if epi == 0:
            out = torch.zeros((M, N), dtype=torch.bfloat16, device=DEV)
        elif epi == 4:
            out = torch.zeros((M, N // 2), dtype=torch.bfloat16, device=DEV)
        else:
            out = torch.zeros((M, N), dtype=torch.float32, device=DEV)
    _lib.check(L().b2_op_gemm(a.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, epi, bn, split_k,
                              float(softcap), act, stream()))
    torch.cuda.synchronize()
    return out


def interleave_gate_up(gate, up):
    """[I,H] x2 -> [2I,H] in 32-row groups (what the engine does at load time)."""
    I, H = gate.shape
    g = gate.view(I // 32, 32, H)
    u = up.view(I // 32, 32, H)
    return torch.stack([g, u], dim=1).reshape(2 * I, H).contiguous()


def op_rmsnorm(x, g, eps, want_f32=False):
    M, H = x.shape
    ob = torch.empty((M, H), dtype=torch.bfloat16, device=DEV)
    of = torch.empty((M, H), dtype=torch.float32, device=DEV) if want_f32 else None
    _lib.check(L().b2_op_rmsnorm(x.data_ptr(), g.data_ptr(), ob.data_ptr(), of.data_ptr() if want_f32 else None,
                                 M, H, float(eps), stream()))
    torch.cuda.synchronize()
    return ob, of


def op_qkv_rope(a, w, B, S, NH, NKV, HD, Tmax, t0, theta, bn=256):
    H = a.shape[1]
    q = torch.zeros((B * S, NH * HD), dtype=torch.bfloat16, device=DEV)
    kc = torch.zeros((B, NKV, Tmax, HD), dtype=torch.bfloat16, device=DEV)
    vc = torch.zeros((B, NKV, Tmax, HD), dtype=torch.bfloat16, device=DEV)
    _lib.check(L().b2_op_qkv_rope(a.data_ptr(), w.data_ptr(), q.data_ptr(), kc.data_ptr(), vc.data_ptr(), B, S, H,
                                  NH, NKV, HD, Tmax, t0, float(theta), bn, stream()))
    torch.cuda.synchronize()
    return q, kc, vc


def op_attention(q, kc, vc, B, S, NH, NKV, HD, Tmax, t0, scale, softcap=0.0, causal=1, window=0, decode=0):
    out = torch.zeros((B * S, NH * HD), dtype=torch.bfloat16, device=DEV)
    _lib.check(L().b2_op_attention(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), out.data_ptr(), B, S, NH, NKV, HD,
                                   Tmax, t0, float(scale), float(softcap), causal, window, decode, stream()))
    torch.cuda.synchronize()
    return out


def ref_attention(q, kc, vc, B, S, NH, NKV, HD, t0, scale, softcap=0.0, causal=1, window=0):
    """fp32 torch restatement of the reference's attention on the same bf16 inputs
    (llama3.2_model_numpy.py:470-511), + the optional HF soft-cap / window."""
    T = t0 + S
    qf = q.float().view(B, S, NH, HD).permute(0, 2, 1, 3)
    k = kc[:, :, :T].float().repeat_interleave(NH // NKV, dim=1)
    v = vc[:, :, :T].float().repeat_interleave(NH // NKV, dim=1)
    s = torch.matmul(qf, k.transpose(-1, -2)) * scale
    if softcap > 0:
        s = torch.tanh(s / softcap) * softcap
    qi = torch.arange(t0, T, device=q.device)[:, None]
    ki = torch.arange(T, device=q.device)[None, :]
    keep = torch.ones((S, T), dtype=torch.bool, device=q.device)
    if causal:
        keep &= ki <= qi
    if window > 0:
        keep &= ki > qi - window
    s = s.masked_fill(~keep, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, v)
    return o.permute(0, 2, 1, 3).reshape(B * S, NH * HD)


def err_stats(got, want):
    d = (got.double() - want.double()).abs()
    return dict(max_abs=float(d.max()), mean_abs=float(d.mean()), ref_absmax=float(want.double().abs().max()),
                ref_std=float(want.double().std()))
