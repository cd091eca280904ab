"""GPU end-to-end parity: model(ids) -> logits / hidden states / KV-cache decode through the public
Python surface (which drives the C ABI), against (a) the committed golden outputs of the reference
itself and (b) the NumPy oracle on the same seeded inputs.

Stated tolerance (bf16 tensor-core operands, fp32 accumulation/residual/logits), set at
~1.5x what is measured on B200 (max/std <= 0.066, mean/std <= 0.0115 over every case below):
    max-abs error  <= 0.10 * std(reference logits)
    mean-abs error <= 0.018 * std(reference logits)
BASELINE.json's literal 1e-2 max-abs is NOT reachable with bf16 operands at logit-std ~0.9 (SURVEY
section 7, 6.3e-2 for the best pure-bf16 policy); the tests print the achieved figures.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import load_golden
from oracle import np_oracle as O
import llm_np_cp_b200 as M

MAX_REL, MEAN_REL = 0.10, 0.018


def build(meta, **kw):
    gem = meta["family"] == "gemma2"
    cfg = (O.gemma2_config if gem else O.llama_config)(meta["config"])
    w = O.synth_weights(cfg, meta["wseed"], meta["std"], meta["jitter"])
    cls = M.Gemma2ForCausalLM if gem else M.LlamaForCausalLM
    return cfg, w, cls(M.AttributeDict(cfg), w, **kw)


def check_close(got, want, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    d = np.abs(got - want)
    std = want.std()
    print(f"[parity] {what}: max_abs={d.max():.4g} mean_abs={d.mean():.4g} ref_std={std:.4g} "
          f"max/std={d.max() / std:.4g} mean/std={d.mean() / std:.4g}")
    assert np.isfinite(got).all(), what
    assert d.max() <= MAX_REL * std, (what, d.max(), std)
    assert d.mean() <= MEAN_REL * std, (what, d.mean(), std)


@pytest.mark.parametrize("name", ["llama_tiny_prefill", "llama_tiny_qlen2", "llama_tiny_qlen1",
                                  "llama_tiny128_prefill", "llama_tiny_long", "gemma2_tiny_prefill"])
def test_prefill_vs_reference_golden(name):
    meta, arr = load_golden(name)
    cfg, w, model = build(meta)
    loss, logits, cache, hs, att = model(arr["ids"], use_cache=False)
    assert loss is None and cache is None and att is None
    assert len(hs) == cfg.num_hidden_layers + 1                       # L+1 hidden states
    assert tuple(logits.shape) == arr["logits"].shape
    check_close(logits.cpu().numpy(), arr["logits"], name + " logits")
    for i, h in enumerate(hs):
        check_close(h.cpu().numpy(), arr["hidden"][i], f"{name} hidden[{i}]")


@pytest.mark.parametrize("name", ["llama_tiny_decode", "gemma2_tiny_decode"])
def test_decode_with_cache_vs_reference_golden(name):
    meta, arr = load_golden(name)
    cfg, w, model = build(meta)
    S, dec = meta["S"], meta["decode_steps"]
    cache = M.KVCache(max_tokens=64)
    _, logits, cache2, _, _ = model(arr["ids"][:, :S], use_cache=True, kv_cache=cache)
    assert cache2 is cache and cache.num_items() == S                # same object mutated and returned
    check_close(logits.cpu().numpy(), arr["logits"], name + " prefill logits")
    for i in range(dec):
        _, lg, _, _, _ = model(arr["ids"][:, S + i:S + i + 1], use_cache=True, kv_cache=cache)
        check_close(lg.cpu().numpy(), arr["decode_logits"][i], f"{name} step {i}")
    assert cache.num_items() == S + dec
    assert tuple(cache.key_cache[0].shape) == (arr["ids"].shape[0], cfg.num_key_value_heads, S + dec, cfg.head_dim)
    # list-like, lazily read per layer (llama3.2_model_numpy.py:313-314 keeps plain per-layer lists)
    assert len(cache.key_cache) == len(cache.value_cache) == cfg.num_hidden_layers
    assert tuple(cache.value_cache[-1].shape) == tuple(cache.key_cache[0].shape)
    assert len(list(cache.key_cache)) == cfg.num_hidden_layers and len(M.KVCache().key_cache) == 0


@pytest.mark.parametrize("name", ["llama_tiny_decode", "gemma2_tiny_decode"])
@pytest.mark.parametrize("kind", ["default", "paged_full_pool", "paged_small_pool"])
def test_paged_cache_vs_reference_golden(name, kind):
    """Paged KV cache (block table + page pool) against the reference's outputs: KVCache() with no
    arguments (paged, pool growing from one page per sequence), an explicit full pool, and a pool that must grow."""
    meta, arr = load_golden(name)
    cfg, w, model = build(meta)
    S, dec = meta["S"], meta["decode_steps"]
    cache = {"default": M.KVCache(), "paged_full_pool": M.KVCache(max_tokens=512, paged=True),
             "paged_small_pool": M.KVCache(max_tokens=1024, paged=True, n_pages=1)}[kind]
    assert cache.paged
    _, logits, cache2, _, _ = model(arr["ids"][:, :S], use_cache=True, kv_cache=cache)
    assert cache2 is cache and cache.num_items() == S
    check_close(logits.cpu().numpy(), arr["logits"], f"{name} paged({kind}) prefill logits")
    for i in range(dec):
        _, lg, _, _, _ = model(arr["ids"][:, S + i:S + i + 1], use_cache=True, kv_cache=cache)
        check_close(lg.cpu().numpy(), arr["decode_logits"][i], f"{name} paged({kind}) step {i}")
    page, pool, used = cache.pool_info()
    B = arr["ids"].shape[0]
    assert page == 128 and used == B and pool >= B          # one logical page per sequence so far
    assert tuple(cache.key_cache[0].shape) == (B, cfg.num_key_value_heads, S + dec, cfg.head_dim)


@pytest.mark.parametrize("family", ["llama", "gemma2"])
def test_paged_cache_is_bit_identical_to_dense(family):
    """Same kernels, same data, different addresses: a paged cache -- with a SHUFFLED block table, and with a pool that grows
    twice while the sequence crosses page boundaries -- gives bit-identical logits and cache contents to the dense cache,
    through chunked prefill over several pages, graph-replayed decode steps across a page boundary, and cache.key_cache."""
    cfg = (O.gemma2_config if family == "gemma2" else O.llama_config)("tiny")
    w = O.synth_weights(cfg, 6, 0.08, 0.1)
    cls = M.Gemma2ForCausalLM if family == "gemma2" else M.LlamaForCausalLM
    model = cls(M.AttributeDict(cfg), w, always_causal=True)          # chunked prefill needs the HF mask semantics
    B, T0, T1, dec = 2, 300, 507, 8                                    # 507 + 8 crosses the 512 = 4-page boundary
    ids = np.random.default_rng(21).integers(0, cfg.vocab_size, (B, T1 + dec))
    rng = np.random.default_rng(5)

    def run(cache):
        model.set_decode_graph(True)
        outs = [model(ids[:, :T0], use_cache=True, kv_cache=cache, output_hidden_states=False)[1],
                model(ids[:, T0:T1], use_cache=True, kv_cache=cache, output_hidden_states=False)[1]]
        for t in range(T1, T1 + dec):
            outs.append(model(ids[:, t:t + 1], use_cache=True, kv_cache=cache, output_hidden_states=False, logits_mode=2)[1])
        model.set_decode_graph(False)
        return outs, cache.key_cache[1].clone(), cache.value_cache[0].clone()

    dense = run(M.KVCache(max_batch=B, max_tokens=1024))
    shuffled = M.KVCache(max_batch=B, max_tokens=1024, paged=True)
    shuffled.set_block_table(rng.permutation(B * 8).reshape(B, 8))
    growing = M.KVCache(max_batch=B, max_tokens=1024, paged=True, n_pages=2)
    for name, cache in (("shuffled block table", shuffled), ("growing pool", growing)):
        got = run(cache)
        for a, b in zip(dense[0], got[0]):
            assert torch.equal(a, b), name
        assert torch.equal(dense[1], got[1]) and torch.equal(dense[2], got[2]), name
    assert growing.pool_info()[1] >= 2 * 5 and growing.pool_info()[2] == 2 * 5      # 5 logical pages per row in use
    with pytest.raises(ValueError):
        bad = M.KVCache(max_batch=B, max_tokens=1024, paged=True)
        bad.set_block_table(np.zeros((B, 8), dtype=np.int32))                       # a page used twice
        model(ids[:, :4], use_cache=True, kv_cache=bad, output_hidden_states=False)


def test_decode_step_with_caller_position_ids():
    """A one-token step on a cache with a caller-supplied position (llama3.2_model_numpy.py:657): RoPE reads the table at
    that position while K/V are appended at the cache length -- through the 5-kernel decode path -- vs the oracle."""
    cfg = O.llama_config("tiny")
    w = O.synth_weights(cfg, 9, 0.08, 0.1)
    wn = O.to_numpy_f32(w)
    model = M.LlamaForCausalLM(M.AttributeDict(cfg), w)
    ids = np.random.default_rng(13).integers(0, cfg.vocab_size, (3, 7))
    oc, cache = O.KVCache(), M.KVCache(max_tokens=32)
    O.forward(cfg, wn, ids[:, :6], oc)
    model(ids[:, :6], use_cache=True, kv_cache=cache, output_hidden_states=False)
    want = O.forward(cfg, wn, ids[:, 6:7], oc, position_ids=[[20]]).logits
    _, got, _, _, _ = model(ids[:, 6:7], position_ids=[[20]], use_cache=True, kv_cache=cache, output_hidden_states=False)
    check_close(got.cpu().numpy(), want, "decode step with position_ids=[[20]] on a 6-token cache")
    plain = O.forward(cfg, wn, ids[:, 6:7], O.KVCache()).logits      # sanity: the position matters
    assert cache.num_items() == 7 and np.abs(want - plain).max() > 1e-2


def test_decode_graph_matches_plain():
    meta, arr = load_golden("llama_tiny_decode")
    cfg, w, model = build(meta)
    ids = np.random.default_rng(3).integers(0, cfg.vocab_size, (2, 12))
    outs = []
    for graph in (False, True):
        model.set_decode_graph(graph)
        cache = M.KVCache(max_tokens=32)
        model(ids[:, :6], use_cache=True, kv_cache=cache, output_hidden_states=False)
        tok = torch.empty((2, 1), dtype=torch.int64, device="cuda")
        lg_buf = []
        for t in range(6, 12):
            tok.copy_(torch.from_numpy(ids[:, t:t + 1]))
            _, lg, _, _, _ = model(tok, use_cache=True, kv_cache=cache, output_hidden_states=False, logits_mode=2)
            lg_buf.append(lg.clone())
        outs.append(torch.stack(lg_buf).cpu().numpy())
    assert np.array_equal(outs[0], outs[1])        # same kernels, same order: bit-identical


@pytest.mark.parametrize("family,graph", [("llama", False), ("llama", True), ("gemma2", False)])
def test_decode_long_cache_vs_oracle(family, graph):
    """Long caches decode through several T-slices of the split-KV tensor-core attention kernel (merged by the last
    CTA); same parity bar as every other path, against the NumPy oracle on the same seeded inputs."""
    cfg = (O.gemma2_config if family == "gemma2" else O.llama_config)("tiny")
    w = O.synth_weights(cfg, 5, 0.08, 0.1)
    model = (M.Gemma2ForCausalLM if family == "gemma2" else M.LlamaForCausalLM)(M.AttributeDict(cfg), w)
    model.set_decode_graph(graph)
    B, S0, dec = 2, 597, 4
    ids = np.random.default_rng(11).integers(0, cfg.vocab_size, (B, S0 + dec))
    wn, oc = O.to_numpy_f32(w), O.KVCache()
    O.forward(cfg, wn, ids[:, :S0], oc)
    cache = M.KVCache(max_tokens=1024)
    model(ids[:, :S0], use_cache=True, kv_cache=cache, output_hidden_states=False)
    for i in range(dec):
        want = O.forward(cfg, wn, ids[:, S0 + i:S0 + i + 1], oc).logits
        _, lg, _, _, _ = model(ids[:, S0 + i:S0 + i + 1], use_cache=True, kv_cache=cache, output_hidden_states=False)
        check_close(lg.cpu().numpy(), want, f"{family} long-cache decode step {i} (graph={graph})")
    assert cache.num_items() == S0 + dec


def test_hf_faithful_llama_vs_transformers_golden():
    """HF-faithful options the reference lacks (llama3 rope scaling, causal mask for every q_len, chunked prefill on a
    non-empty cache) against outputs of transformers' LlamaForCausalLM (tests/golden/hf_llama_tiny_rope3.npz)."""
    meta, arr = load_golden("hf_llama_tiny_rope3")
    cfg, w, model = build(meta, always_causal=True, rope_scaling=meta["rope_scaling"])
    S, dec = meta["S"], meta["decode_steps"]
    cache = M.KVCache(max_tokens=256)
    _, logits, _, hs, _ = model(arr["ids"][:, :S], use_cache=True, kv_cache=cache)
    check_close(logits.cpu().numpy(), arr["logits"], "hf llama3-rope prefill logits")
    for i, h in enumerate(hs):
        check_close(h.cpu().numpy(), arr["hidden"][i], f"hf llama3-rope hidden[{i}]")
    for i in range(dec):
        _, lg, _, _, _ = model(arr["ids"][:, S + i:S + i + 1], use_cache=True, kv_cache=cache)
        check_close(lg.cpu().numpy(), arr["decode_logits"][i], f"hf llama3-rope decode step {i}")
    _, q2, _, _, _ = model(arr["ids"][:, :2], use_cache=False)
    check_close(q2.cpu().numpy(), arr["qlen2_logits"], "hf q_len=2 (masked, unlike the reference)")
    # chunked prefill: 100 tokens, then 50 more on the non-empty cache (the reference raises here)
    cache2 = M.KVCache(max_tokens=256)
    model(arr["ids"][:, :100], use_cache=True, kv_cache=cache2, output_hidden_states=False)
    _, lg2, _, _, _ = model(arr["ids"][:, 100:S], use_cache=True, kv_cache=cache2, output_hidden_states=False)
    check_close(lg2.cpu().numpy(), arr["logits"][:, 100:S], "hf chunked prefill (second chunk)")
    # without the option the engine reproduces the reference, not HF
    _, _, plain = build(meta)
    _, lp, _, _, _ = plain(arr["ids"][:, :S], use_cache=False)
    assert np.abs(lp.cpu().numpy() - arr["logits"]).max() > 0.15 * arr["logits"].std()


def test_chunk_on_cache_raises_like_reference():
    meta, arr = load_golden("llama_tiny_chunk_on_cache")
    cfg = O.llama_config("tiny")
    w = O.synth_weights(cfg, 1, 0.08, 0.1)
    model = M.LlamaForCausalLM(M.AttributeDict(cfg), w)
    cache = M.KVCache(max_tokens=32)
    model(arr["ids"][:, :4], use_cache=True, kv_cache=cache)
    with pytest.raises(ValueError, match="broadcast"):
        model(arr["ids"][:, 4:8], use_cache=True, kv_cache=cache)


def test_llama_1b_config1_vs_reference_golden_and_oracle():
    """BASELINE.json config #1: Llama-3.2-1B shapes, 16-token prompt, random-init weights."""
    meta, arr = load_golden("llama_1b_s16")
    cfg, w, model = build(meta)
    _, logits, _, hs, _ = model(arr["ids"], use_cache=False)
    lg = logits.cpu().numpy()
    check_close(lg[:, :, ::61], arr["logits_sub"], "1B logits (sub-sampled golden)")
    check_close(lg[:, -1, :], arr["logits_last"], "1B last-position logits")
    agree = (lg.argmax(-1) == arr["argmax"]).mean()
    print(f"[parity] 1B argmax agreement {agree:.3f}")
    assert agree >= 0.98
    hid = torch.stack(hs).cpu().numpy()[:, :, :, ::8]
    for i in range(hid.shape[0]):
        d = np.abs(hid[i] - arr["hidden_sub"][i])
        assert d.max() <= 0.1 * arr["hidden_sub"][i].std() + 1e-3, (i, d.max())
    # batch rows are independent: B=3 copies give identical logits (data-parallel sharding invariant)
    ids3 = np.repeat(arr["ids"], 3, axis=0)
    _, lg3, _, _, _ = model(ids3, use_cache=False, output_hidden_states=False)
    assert torch.equal(lg3[0], lg3[1]) and torch.equal(lg3[0], lg3[2])
    _, lg1, _, _, _ = model(arr["ids"], use_cache=False, output_hidden_states=False)
    assert torch.equal(lg3[0], lg1[0])          # same sequence alone or inside a batch: bit-identical


def test_call_variants_vs_reference_golden():
    """The other call forms of the reference's __call__, against outputs of the reference itself:
    inputs_embeds, caller position_ids, a kv_cache passed with use_cache=False, and the `attentions` tuple through the
    opt-in attention-probability path."""
    meta, arr = load_golden("llama_tiny_call_variants")
    cfg, w, model = build(meta)
    model.materialize_attentions = True
    _, logits, _, _, att = model(arr["ids"], use_cache=False)
    check_close(logits.cpu().numpy(), arr["logits"], "variants: plain logits")
    assert len(att) == cfg.num_hidden_layers and tuple(att[0].shape) == arr["attn_all"][0].shape
    got = torch.stack(att).cpu().numpy()
    d = np.abs(got - arr["attn_all"])
    print(f"[parity] attentions (all {len(att)} layers): max_abs={d.max():.4g} (probabilities), row sums {got.sum(-1).min():.6f}..{got.sum(-1).max():.6f}")
    assert d.max() <= 0.02 and np.abs(got.sum(-1) - 1).max() < 1e-4
    assert (got[arr["attn_all"] == 0] == 0).all()                      # masked entries are exactly zero, like exp(-inf)
    model.materialize_attentions = False
    assert model(arr["ids"], use_cache=False)[4] is None
    _, le, _, _, _ = model(None, inputs_embeds=arr["embeds"], use_cache=False)
    check_close(le.cpu().numpy(), arr["embeds_logits"], "variants: inputs_embeds")
    _, lp, _, _, _ = model(arr["ids"], position_ids=arr["pos_ids"], use_cache=False)
    check_close(lp.cpu().numpy(), arr["pos_logits"], "variants: position_ids")
    with pytest.raises(ValueError):
        model(arr["ids"], inputs_embeds=arr["embeds"])
    with pytest.raises(ValueError):
        model(None)
    cache = M.KVCache(max_tokens=32)
    _, l1, c1, _, _ = model(arr["ids"][:, :5], use_cache=False, kv_cache=cache)
    assert c1 is cache and cache.num_items() == meta["nocache_items_after_call1"]      # appended although use_cache=False
    model.materialize_attentions = True
    _, l2, _, _, a2 = model(arr["ids"][:, 5:6], use_cache=False, kv_cache=cache)
    assert cache.num_items() == meta["nocache_items_after_call2"]
    check_close(l1.cpu().numpy(), arr["nocache_logits1"], "variants: use_cache=False call 1")
    check_close(l2.cpu().numpy(), arr["nocache_logits2"], "variants: use_cache=False call 2 (positions restart at 0)")
    assert np.abs(a2[0].cpu().numpy() - arr["nocache_attn2"]).max() <= 0.02


@pytest.mark.parametrize("tag", ["default", "swapped"])
def test_gemma2_hf_flags_vs_transformers_golden(tag):
    """Gemma-2 HF semantics (attention soft-cap, query_pre_attn_scalar, sliding window per layer_types) against outputs
    of transformers' Gemma2ForCausalLM (tests/golden/hf_gemma2_tiny_flags.npz, oracle/make_golden_r2.py)."""
    meta, arr = load_golden("hf_gemma2_tiny_flags")
    cfg = O.gemma2_config("tiny", sliding_window=meta["sliding_window"], query_pre_attn_scalar=meta["query_pre_attn_scalar"])
    cfg["layer_types"] = meta["layer_types"][tag]
    w = O.synth_weights(cfg, meta["wseed"], meta["std"], meta["jitter"])
    model = M.Gemma2ForCausalLM(M.AttributeDict(cfg), w, hf_flags=True)
    _, logits, _, hs, _ = model(arr["ids"], use_cache=False)
    check_close(logits.cpu().numpy(), arr[f"logits_{tag}"], f"gemma2 HF flags vs transformers ({tag} layer_types)")
    for i, h in enumerate(hs):
        check_close(h.cpu().numpy(), arr[f"hidden_{tag}"][i], f"gemma2 HF flags hidden[{i}] ({tag})")
    other = "swapped" if tag == "default" else "default"
    assert np.abs(logits.cpu().numpy() - arr[f"logits_{other}"]).max() > 0.15 * arr[f"logits_{other}"].std()


def test_decode_graph_survives_cache_recreation():
    """Captured decode graphs are keyed on a never-reused cache id, not on the cache's heap address: destroying a cache
    and creating another one (very likely at the same address) with the same stable id / logits buffers must not replay
    a graph that points into the freed cache."""
    meta, arr = load_golden("llama_tiny_decode")
    cfg, w, model = build(meta)
    ids = np.random.default_rng(8).integers(0, cfg.vocab_size, (2, 10))
    tok = torch.empty((2, 1), dtype=torch.int64, device="cuda")
    lg = torch.empty((2, cfg.vocab_size), dtype=torch.float32, device="cuda")

    def run(prompt_len, graph):
        model.set_decode_graph(graph)
        cache = M.KVCache(max_tokens=16)
        model(ids[:, :prompt_len], use_cache=True, kv_cache=cache, output_hidden_states=False)
        out = []
        for t in range(prompt_len, 10):
            tok.copy_(torch.from_numpy(ids[:, t:t + 1]))
            model(tok, use_cache=True, kv_cache=cache, output_hidden_states=False, logits_mode=2, logits_out=lg)
            out.append(lg.clone())
        return torch.stack(out).cpu().numpy()

    a = run(4, True)
    b = run(6, True)               # new cache, same tok/lg pointers, different history
    assert np.array_equal(a, run(4, False)) and np.array_equal(b, run(6, False))
    # without a caller logits buffer every step still replays one graph (model-owned step buffers)
    model.set_decode_graph(True)
    cache = M.KVCache(max_tokens=16)
    model(ids[:, :4], use_cache=True, kv_cache=cache, output_hidden_states=False)
    outs = [model(ids[:, t:t + 1], use_cache=True, kv_cache=cache, output_hidden_states=False, logits_mode=2)[1] for t in range(4, 10)]
    assert np.array_equal(torch.stack(outs).cpu().numpy(), a)


def test_llama_1b_s128_vs_reference_golden():
    """Llama-3.2-1B shapes, 128 tokens, all 16 layers, against outputs of the reference ITSELF (not the oracle): the mid-size
    point between config #1 (16 tokens) and the oracle-checked benchmarked shapes.  16 M logits are compared, so the
    maximum uses the large-sample tolerance of tests/test_gpu_baseline_shapes.py (0.16 std); the mean keeps 0.018."""
    meta, arr = load_golden("llama_1b_s128")
    cfg, w, model = build(meta)
    _, logits, _, hs, _ = model(arr["ids"], use_cache=False)
    lg = logits.cpu().numpy().astype(np.float64)
    std = arr["logits_stats"][1]
    for what, got, want in (("sub-sampled columns", lg[:, :, ::61], arr["logits_sub"]), ("last position", lg[:, -1, :], arr["logits_last"])):
        d = np.abs(got - want)
        print(f"[parity] 1B S=128 logits ({what}): max_abs={d.max():.4g} mean_abs={d.mean():.4g} ref_std={std:.4g} "
              f"max/std={d.max() / std:.4g} mean/std={d.mean() / std:.4g}")
        assert d.max() <= 0.16 * std and d.mean() <= MEAN_REL * std
    agree = (lg.argmax(-1) == arr["argmax"]).mean()
    print(f"[parity] 1B S=128 argmax agreement {agree:.3f}")
    assert agree >= 0.9
    hid = torch.stack(hs).cpu().numpy()[:, :, :, ::8]
    for i in range(hid.shape[0]):
        d = np.abs(hid[i] - arr["hidden_sub"][i])
        assert d.max() <= 0.1 * arr["hidden_sub"][i].std() + 1e-3, (i, d.max())


def test_gemma2_hf_flags_vs_oracle_extras():
    """Sliding window + attention soft-cap + query_pre_attn_scalar (absent from the reference; semantics
    of transformers' modeling_gemma2, restated in the oracle's Extras)."""
    cfg = O.gemma2_config("tiny", sliding_window=16, query_pre_attn_scalar=144)
    w = O.synth_weights(cfg, 4, 0.05, 0.1)
    ids = np.random.default_rng(1).integers(0, cfg.vocab_size, (2, 70))
    ex = O.Extras(attn_softcap=50.0, sliding_window=16, attn_scale=144 ** -0.5, always_causal=True)
    want = O.forward(cfg, O.to_numpy_f32(w), ids, extras=ex).logits
    model = M.Gemma2ForCausalLM(M.AttributeDict(cfg), w, hf_flags=True)
    _, logits, _, _, _ = model(ids, use_cache=False, output_hidden_states=False)
    check_close(logits.cpu().numpy(), want, "gemma2 HF flags")


def test_forward_host_and_sampling():
    meta, arr = load_golden("llama_tiny_prefill")
    cfg, w, model = build(meta)
    out = model.forward_host(arr["ids"], logits_mode=2)
    check_close(out, arr["logits"][:, -1, :], "forward_host last logits")
    # min-p: device distribution == oracle distribution; draws land on kept tokens
    lg = torch.from_numpy(arr["logits"][:, -1, :].astype(np.float32))
    tok, probs = M.min_p_sampling(lg, 0.1, generator=torch.Generator().manual_seed(0), model=model, return_probs=True)
    want = O.min_p_probs(arr["logits"][:, -1, :].astype(np.float32))
    assert np.abs(probs.cpu().numpy() - want).max() < 1e-5
    for b in range(lg.shape[0]):
        assert want[b, int(tok[b])] > 0
    # greedy pick == np.argmax (lowest index on ties)
    lgd = lg.to("cuda")
    lgd[1, 7] = lgd[1].max() + 1.0
    lgd[1, 300] = lgd[1, 7]
    assert model.argmax(lgd).cpu().tolist() == lgd.cpu().numpy().argmax(-1).tolist()
    # inverse-CDF property: frequencies follow the distribution
    cnt = np.zeros(cfg.vocab_size)
    g = torch.Generator().manual_seed(1)
    big = lg[0:1].repeat(2048, 1)
    t = M.min_p_sampling(big, 0.1, generator=g, model=model).cpu().numpy()
    np.add.at(cnt, t, 1)
    assert np.abs(cnt / 2048 - want[0]).max() < 0.06


# ---------------------------------------------------------------------------------------------
# Precise (hi/lo split-bf16) parity mode: BASELINE.json's literal target -- logits within 1e-2 max-abs of
# the reference's NumPy path.  Same kernels, every bf16 activation carried as hi + lo planes.
# ---------------------------------------------------------------------------------------------
PRECISE_TOL = 1e-2


@pytest.mark.parametrize("name", ["llama_tiny_prefill", "llama_tiny128_prefill", "llama_tiny_long", "llama_tiny_qlen2",
                                  "gemma2_tiny_prefill"])      # head_dim 64 / 128 / 256 (Gemma-2: 32-wide kv tiles)
def test_precise_prefill_within_1e2_of_reference(name):
    meta, arr = load_golden(name)
    cfg, w, model = build(meta, precise=True)
    _, logits, _, hs, _ = model(arr["ids"], use_cache=False)
    d = np.abs(logits.cpu().numpy().astype(np.float64) - arr["logits"])
    print(f"[parity-precise] {name}: max_abs={d.max():.3g} mean_abs={d.mean():.3g}")
    assert d.max() <= PRECISE_TOL
    for i, h in enumerate(hs):
        assert np.abs(h.cpu().numpy() - arr["hidden"][i]).max() <= PRECISE_TOL


@pytest.mark.parametrize("paged", [False, True])
@pytest.mark.parametrize("name", ["llama_tiny_decode", "gemma2_tiny_decode"])
def test_precise_decode_within_1e2_of_reference(name, paged):
    meta, arr = load_golden(name)
    cfg, w, model = build(meta, precise=True)
    S, dec = meta["S"], meta["decode_steps"]
    cache = M.KVCache(max_tokens=256, paged=True) if paged else M.KVCache(max_tokens=64)   # paged: hi and lo planes are page pools
    _, logits, _, _, _ = model(arr["ids"][:, :S], use_cache=True, kv_cache=cache)
    assert np.abs(logits.cpu().numpy() - arr["logits"]).max() <= PRECISE_TOL
    for i in range(dec):
        _, lg, _, _, _ = model(arr["ids"][:, S + i:S + i + 1], use_cache=True, kv_cache=cache, output_hidden_states=False)
        d = np.abs(lg.cpu().numpy() - arr["decode_logits"][i]).max()
        print(f"[parity-precise] decode step {i}: max_abs={d:.3g}")
        assert d <= PRECISE_TOL


def test_precise_llama_1b_config1_within_1e2():
    """BASELINE.json config #1 at the literal 1e-2 max-abs tolerance (logit std 0.9)."""
    meta, arr = load_golden("llama_1b_s16")
    cfg, w, model = build(meta, precise=True)
    _, logits, _, _, _ = model(arr["ids"], use_cache=False, output_hidden_states=False)
    lg = logits.cpu().numpy().astype(np.float64)
    d1 = np.abs(lg[:, :, ::61] - arr["logits_sub"]).max()
    d2 = np.abs(lg[:, -1, :] - arr["logits_last"]).max()
    print(f"[parity-precise] 1B config #1: max_abs sub={d1:.3g} last={d2:.3g}")
    assert max(d1, d2) <= PRECISE_TOL
    assert (lg.argmax(-1) == arr["argmax"]).all()
