"""GPU parity of every hand-written kernel, called through the C ABI, against fp32 torch restatements
of the reference's ops fed the same bf16-rounded inputs (SURVEY.md section 4 'per-kernel parity').
Tolerances are stated per test."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import (DEV, bf16_randn, err_stats, interleave_gate_up, op_attention, op_gemm, op_qkv_rope,
                      op_rmsnorm, ref_attention)


@pytest.mark.parametrize("M,H", [(5, 256), (37, 2048), (300, 2304), (64, 3072)])
def test_rmsnorm(M, H):
    g = torch.Generator().manual_seed(M)
    x = (torch.randn(M, H, generator=g) * 3).to(DEV)
    w = (1 + 0.1 * torch.randn(H, generator=g)).to(DEV)
    ob, of = op_rmsnorm(x, w, 1e-5, want_f32=True)
    ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * w
    assert (of - ref).abs().max() < 1e-5 * ref.abs().max() + 1e-6          # fp32 path: rounding only
    assert (ob.float() - ref).abs().max() <= ref.abs().max() * 2 ** -8      # bf16 output rounding


GEMM_SHAPES = [
    (128, 256, 64), (128, 256, 256), (256, 512, 2048), (100, 256, 128), (1000, 768, 512),
    (2048, 3072, 2048), (333, 2304, 2304),
    (64, 512, 2048), (40, 256, 512), (32, 256, 256), (7, 256, 1024), (1, 512, 128),   # few-row path: 64- / 32-row A boxes
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("bn", [256, 128, 64])
def test_gemm_f32(M, N, K, bn):
    a, w = bf16_randn(M, K, seed=1), bf16_randn(N, K, seed=2)
    out = op_gemm(a, w, epi=1, bn=bn)
    ref = a.float() @ w.float().T
    st = err_stats(out, ref)
    # bf16 x bf16 products are exact in fp32; only accumulation order differs
    assert st["max_abs"] <= 2e-3 * math.sqrt(K / 64) , st


@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (77, 256, 128)])
def test_gemm_bf16_and_softcap(M, N, K):
    a, w = bf16_randn(M, K, seed=3), bf16_randn(N, K, seed=4)
    ref = a.float() @ w.float().T
    out = op_gemm(a, w, epi=0)
    assert (out.float() - ref).abs().max() <= ref.abs().max() * 2 ** -8 + 1e-3
    cap = 30.0
    out = op_gemm(a, w, epi=1, softcap=cap)
    assert (out - torch.tanh(ref / cap) * cap).abs().max() < 1e-3


@pytest.mark.parametrize("M,N,K,bn,split", [(256, 256, 512, 256, 1), (64, 512, 1024, 64, 4), (64, 2048, 8192, 64, 4),
                                            (16, 256, 2048, 32, 2), (2000, 2048, 2048, 256, 1)])
def test_gemm_residual(M, N, K, bn, split):
    a, w = bf16_randn(M, K, seed=5), bf16_randn(N, K, seed=6, std=0.05)
    x0 = torch.randn(M, N, generator=torch.Generator().manual_seed(7)).to(DEV)
    x = x0.clone()
    op_gemm(a, w, epi=2, bn=bn, split_k=split, out=x)
    ref = x0 + a.float() @ w.float().T
    assert (x - ref).abs().max() < 2e-3 * math.sqrt(K / 64), err_stats(x, ref)


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("M,I,K,bn", [(256, 512, 256, 256), (90, 8192, 2048, 256), (64, 512, 256, 64)])
def test_gemm_glu(M, I, K, bn, act):
    a = bf16_randn(M, K, seed=8)
    gate, up = bf16_randn(I, K, seed=9, std=0.05), bf16_randn(I, K, seed=10, std=0.05)
    out = op_gemm(a, interleave_gate_up(gate, up), epi=4, bn=bn, act=act)
    g, u = a.float() @ gate.float().T, a.float() @ up.float().T
    actf = torch.nn.functional.silu(g) if act == 0 else torch.nn.functional.gelu(g, approximate="tanh")
    ref = actf * u
    assert (out.float() - ref).abs().max() <= ref.abs().max() * 2 ** -7 + 2e-3, err_stats(out.float(), ref)


def _rope_ref(x, pos, HD, theta):
    """rotate-half RoPE with the reference's fp32 table arithmetic (llama3.2_model_numpy.py:42-90)."""
    inv = 1.0 / (theta ** (np.arange(0, HD, 2).astype(np.float32) / HD))
    ang = (inv[None, :] * pos.astype(np.float32)[:, None]).astype(np.float32)
    emb = np.concatenate([ang, ang], -1)
    cos, sin = torch.from_numpy(np.cos(emb)).to(x.device), torch.from_numpy(np.sin(emb)).to(x.device)
    h = HD // 2
    rot = torch.cat([-x[..., h:], x[..., :h]], -1)
    return x * cos[:, None, :] + rot * sin[:, None, :]


@pytest.mark.parametrize("B,S,H,NH,NKV,HD,t0,bn", [(2, 12, 256, 4, 2, 64, 0, 256), (1, 200, 256, 2, 1, 128, 7, 256),
                                                   (2, 9, 256, 2, 1, 256, 3, 256), (3, 1, 256, 4, 2, 64, 5, 64),
                                                   (1, 300, 2048, 32, 8, 64, 0, 256)])
def test_qkv_rope_and_cache_append(B, S, H, NH, NKV, HD, t0, bn):
    theta, Tmax = 500000.0, t0 + S + 5
    a = bf16_randn(B * S, H, seed=11)
    w = bf16_randn((NH + 2 * NKV) * HD, H, seed=12, std=0.05)
    q, kc, vc = op_qkv_rope(a, w, B, S, NH, NKV, HD, Tmax, t0, theta, bn=bn)
    y = (a.float() @ w.float().T).view(B, S, NH + 2 * NKV, HD)
    pos = np.arange(t0, t0 + S)
    tol = lambda ref: ref.abs().max() * 2 ** -7 + 1e-3
    for b in range(B):
        qr = _rope_ref(y[b, :, :NH], pos, HD, theta)
        kr = _rope_ref(y[b, :, NH:NH + NKV], pos, HD, theta)
        vr = y[b, :, NH + NKV:]
        assert (q.view(B, S, NH, HD)[b].float() - qr).abs().max() <= tol(qr)
        assert (kc[b, :, t0:t0 + S].float().permute(1, 0, 2) - kr).abs().max() <= tol(kr)
        assert (vc[b, :, t0:t0 + S].float().permute(1, 0, 2) - vr).abs().max() <= tol(vr)
    assert kc[:, :, :t0].abs().sum() == 0 and kc[:, :, t0 + S:].abs().sum() == 0    # nothing outside the appended rows


ATT_CASES = [
    # B, S, NH, NKV, HD, t0, causal, window, softcap
    (1, 128, 1, 1, 64, 0, 1, 0, 0.0),
    (2, 12, 4, 2, 64, 0, 1, 0, 0.0),
    (1, 2, 4, 2, 64, 0, 0, 0, 0.0),        # the reference's q_len <= 2 quirk: unmasked
    (1, 300, 4, 2, 64, 0, 1, 0, 0.0),
    (2, 520, 8, 2, 64, 0, 1, 0, 0.0),
    (1, 260, 2, 1, 128, 0, 1, 0, 0.0),
    (1, 200, 2, 1, 256, 0, 1, 0, 0.0),
    (1, 130, 4, 2, 64, 70, 1, 0, 0.0),     # chunk on a non-empty cache (always_causal extension)
    (1, 400, 2, 1, 64, 0, 1, 100, 0.0),    # sliding window
    (1, 300, 2, 1, 256, 0, 1, 64, 50.0),   # Gemma-2 HF flags: window + soft-cap
    (1, 2048, 2, 1, 64, 0, 1, 0, 0.0),
    # co-resident kernel (head_dim 64 / 128, S > 128): ragged S, non-causal, cache offsets that are not tile-aligned,
    # window / soft-cap combinations, GQA ratios 1 / 3 / 8, a batch that needs several waves
    (3, 129, 3, 3, 64, 0, 1, 0, 0.0),
    (1, 257, 8, 1, 64, 0, 0, 0, 0.0),
    (2, 200, 6, 2, 128, 333, 1, 0, 0.0),
    (1, 700, 4, 2, 128, 0, 1, 150, 30.0),
    (2, 384, 4, 4, 64, 129, 1, 97, 0.0),
    (1, 1000, 3, 1, 128, 0, 1, 0, 50.0),
    (16, 512, 8, 8, 64, 0, 1, 0, 0.0),
    (1, 4096, 2, 2, 128, 0, 1, 0, 0.0),
]


@pytest.mark.parametrize("B,S,NH,NKV,HD,t0,causal,window,softcap", ATT_CASES)
def test_attention_prefill(B, S, NH, NKV, HD, t0, causal, window, softcap):
    Tmax = t0 + S + 3
    q = bf16_randn(B * S, NH * HD, seed=13)
    kc, vc = bf16_randn(B, NKV, Tmax, HD, seed=14), bf16_randn(B, NKV, Tmax, HD, seed=15)
    scale = 1.0 / math.sqrt(HD)
    out = op_attention(q, kc, vc, B, S, NH, NKV, HD, Tmax, t0, scale, softcap, causal, window)
    ref = ref_attention(q, kc, vc, B, S, NH, NKV, HD, t0, scale, softcap, causal, window)
    st = err_stats(out.float(), ref)
    # P is rounded to bf16 before P.V and the output is bf16: ~2^-8 relative on O(1) values
    assert st["max_abs"] <= 2e-2, st


@pytest.mark.parametrize("B,NH,NKV,HD,t0,window,softcap", [(3, 4, 2, 64, 0, 0, 0.0), (2, 32, 8, 64, 143, 0, 0.0),
                                                          (2, 24, 8, 128, 77, 0, 0.0), (2, 8, 4, 256, 300, 64, 50.0),
                                                          (64, 32, 8, 64, 2100, 0, 0.0),
                                                          # split-KV tensor-core kernel: tile/slice boundaries, every
                                                          # head ratio, window, soft-cap, many slices
                                                          (1, 32, 8, 64, 511, 0, 0.0), (2, 32, 8, 64, 575, 0, 0.0),
                                                          (2, 32, 8, 64, 576, 0, 0.0), (1, 8, 8, 64, 1023, 0, 0.0),
                                                          (3, 16, 2, 64, 700, 0, 0.0), (2, 24, 8, 128, 3000, 0, 0.0),
                                                          (2, 8, 4, 256, 1500, 0, 0.0), (2, 8, 4, 256, 1500, 600, 50.0),
                                                          (2, 12, 4, 128, 900, 100, 30.0), (5, 32, 8, 64, 8000, 0, 0.0),
                                                          (2, 16, 1, 64, 200, 0, 0.0)])   # n_rep 16 > 8: per-head kernel
def test_attention_decode(B, NH, NKV, HD, t0, window, softcap):
    Tmax = t0 + 4
    q = bf16_randn(B, NH * HD, seed=16)
    kc, vc = bf16_randn(B, NKV, Tmax, HD, seed=17), bf16_randn(B, NKV, Tmax, HD, seed=18)
    scale = 1.0 / math.sqrt(HD)
    out = op_attention(q, kc, vc, B, 1, NH, NKV, HD, Tmax, t0, scale, softcap, 0, window, decode=1)
    ref = ref_attention(q, kc, vc, B, 1, NH, NKV, HD, t0, scale, softcap, 1, window)
    assert (out.float() - ref).abs().max() <= 1e-2, err_stats(out.float(), ref)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (256, 512, 256), (1000, 768, 512), (2048, 3072, 2048), (333, 2304, 2304)])
def test_gemm_2cta_f32(M, N, K):
    """cta_group::2 kernel (bn=512 selects it): 256 x 256 tiles per CTA pair."""
    a, w = bf16_randn(M, K, seed=21), bf16_randn(N, K, seed=22)
    out = op_gemm(a, w, epi=1, bn=512)
    ref = a.float() @ w.float().T
    assert (out - ref).abs().max() <= 2e-3 * math.sqrt(K / 64), err_stats(out, ref)


def test_gemm_2cta_epilogues():
    M, I, K = 700, 1024, 512
    a = bf16_randn(M, K, seed=23)
    gate, up = bf16_randn(I, K, seed=24, std=0.05), bf16_randn(I, K, seed=25, std=0.05)
    out = op_gemm(a, interleave_gate_up(gate, up), epi=4, bn=512, act=0)
    g, u = a.float() @ gate.float().T, a.float() @ up.float().T
    ref = torch.nn.functional.silu(g) * u
    assert (out.float() - ref).abs().max() <= ref.abs().max() * 2 ** -7 + 2e-3
    w = bf16_randn(512, K, seed=26, std=0.05)
    x0 = torch.randn(M, 512, generator=torch.Generator().manual_seed(7)).to(DEV)
    x = x0.clone()
    op_gemm(a, w, epi=2, bn=512, out=x)
    assert (x - (x0 + a.float() @ w.float().T)).abs().max() < 5e-3


@pytest.mark.parametrize("M,N,K,cluster", [(64, 2048, 2048, 4), (64, 2048, 8192, 4), (64, 2048, 2048, 0), (33, 256, 512, 2),
                                           (7, 512, 256, 1), (64, 3072, 3072, 3), (1, 256, 256, 4), (32, 2048, 2048, 2)])
def test_gemm_cluster_splitk_resid_norm(M, N, K, cluster):
    """Decode-step GEMM (gemm_sk_tcgen05.cuh): the K range split over the CTAs of a thread-block cluster, partial tiles
    reduced through distributed shared memory, epilogue = residual add + first half of the following RMSNorm:
    x += a.W^T (fp32), a' = bf16(x * gamma), ss = per-32-column sums of x^2."""
    from gpu_util import L, stream
    from llm_np_cp_b200 import _lib
    a = bf16_randn(M, K, std=1.0, seed=1)
    w = bf16_randn(N, K, std=0.05, seed=2)
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(M, N, generator=g).to(DEV)
    gamma = (1 + 0.2 * torch.randn(N, generator=g)).to(DEV)
    x = x0.clone()
    a_out = torch.zeros((M, N), dtype=torch.bfloat16, device=DEV)
    ss = torch.full((M, N // 32), -1.0, dtype=torch.float32, device=DEV)
    _lib.check(L().b2_op_gemm_sk(a.data_ptr(), w.data_ptr(), x.data_ptr(), a_out.data_ptr(), gamma.data_ptr(), ss.data_ptr(),
                                 M, N, K, cluster, stream()))
    torch.cuda.synchronize()
    ref = x0.double() + a.double() @ w.double().T
    tol = 2e-5 * math.sqrt(K) * float(ref.abs().max()) + 1e-5           # fp32 accumulation-order differences only
    assert (x.double() - ref).abs().max() <= tol, err_stats(x, ref.float())
    want_a = (x * gamma).to(torch.bfloat16)                              # from the kernel's own fp32 x: exact rounding expected
    assert (a_out.float() - want_a.float()).abs().max() <= want_a.float().abs().max() * 2 ** -7
    want_ss = x.double().pow(2).view(M, N // 32, 32).sum(-1)
    assert (ss.double() - want_ss).abs().max() <= 1e-5 * float(want_ss.max()) + 1e-6
    # determinism: the reduction order inside the cluster is fixed
    x2 = x0.clone()
    _lib.check(L().b2_op_gemm_sk(a.data_ptr(), w.data_ptr(), x2.data_ptr(), a_out.data_ptr(), gamma.data_ptr(), ss.data_ptr(),
                                 M, N, K, cluster, stream()))
    torch.cuda.synchronize()
    assert torch.equal(x, x2)
