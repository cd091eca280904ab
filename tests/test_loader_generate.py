"""load_model / load_sharded_safetensors_via_weight_map / generate: the checkpoint loader
(llama3.2_model_numpy.py:981-1047) on a synthetic HF-style snapshot written to a temp dir, and the reference's decode
loop with its quirks (:873-910) driven through the engine."""
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import np_oracle as O


def write_snapshot(d, sharded=True, seed=3):
    """Tiny Llama snapshot: config.json, (sharded) safetensors, a WordLevel fast tokenizer with BOS."""
    from safetensors.torch import save_file
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    from transformers import PreTrainedTokenizerFast
    cfg = O.llama_config("tiny")
    w = {k: v.contiguous() for k, v in O.synth_weights(cfg, seed, 0.08, 0.1).items()}
    names = sorted(w)
    if sharded:
        half = len(names) // 2
        shards = {"model-00001-of-00002.safetensors": names[:half], "model-00002-of-00002.safetensors": names[half:]}
        for fn, ks in shards.items():
            save_file({k: w[k] for k in ks}, os.path.join(d, fn))
        with open(os.path.join(d, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {}, "weight_map": {k: fn for fn, ks in shards.items() for k in ks}}, f)
    else:
        save_file(w, os.path.join(d, "model.safetensors"))
    hf = dict(model_type="llama", vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
              num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
              num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim, rope_theta=cfg.rope_theta,
              rms_norm_eps=cfg.rms_norm_eps, hidden_act="silu", max_position_embeddings=512, tie_word_embeddings=True,
              attention_dropout=0.0, use_cache=True,
              rope_scaling={"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                            "original_max_position_embeddings": 64})
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(hf, f)
    vocab = {"<s>": 0, "</s>": 1, "<unk>": 2}
    vocab.update({f"w{i}": i for i in range(3, cfg.vocab_size)})
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    tok.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 0)])
    PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<s>", eos_token="</s>", unk_token="<unk>").save_pretrained(d)
    return cfg, w


def test_loader_sharded_and_single_file(tmp_path):
    from llm_np_cp_b200.model import load_sharded_safetensors_via_weight_map
    a, b = tmp_path / "sharded", tmp_path / "single"
    a.mkdir(), b.mkdir()
    cfg, w = write_snapshot(str(a), sharded=True)
    write_snapshot(str(b), sharded=False)
    ta, ca = load_sharded_safetensors_via_weight_map(str(a))
    tb, cb = load_sharded_safetensors_via_weight_map(str(b))
    assert set(ta) == set(tb) == set(w)
    for k in w:
        assert ta[k].dtype == torch.bfloat16 and torch.equal(ta[k], w[k]) and torch.equal(tb[k], w[k])
    assert ca.hidden_size == cfg.hidden_size and ca["num_hidden_layers"] == cfg.num_hidden_layers   # attribute + item access
    assert "lm_head.weight" not in ta                         # tied: the engine aliases it to the embedding (:1025-1026)
    with open(a / "model.safetensors.index.json", "w") as f:
        json.dump({"weight_map": {}}, f)
    with pytest.raises(ValueError, match="weight_map"):      # same message as the reference (:1001-1002)
        load_sharded_safetensors_via_weight_map(str(a))


@pytest.mark.gpu
def test_load_model_and_generate(tmp_path, capsys):
    import llm_np_cp_b200 as M
    cfg, w = write_snapshot(str(tmp_path))
    tokenizer, model, config = M.load_model(str(tmp_path))
    assert isinstance(model, M.LlamaForCausalLM) and config.head_dim == cfg.head_dim
    prompt, n_new = "w5 w17 w300", 6
    ids0 = tokenizer(prompt, return_tensors="pt").input_ids.numpy()
    assert ids0.tolist() == [[0, 5, 17, 300]]                 # BOS on the first call
    # the engine's logits for the prompt agree with the oracle on the same checkpoint (reference semantics: the
    # checkpoint's rope_scaling is ignored unless hf_faithful=True)
    wn = O.to_numpy_f32(w)
    want = O.forward(cfg, wn, ids0).logits
    _, lg, _, _, _ = model(ids0, use_cache=False)
    assert np.abs(lg.cpu().numpy() - want).max() <= 0.15 * want.std()
    outs = []
    for rep in range(2):
        cache = M.KVCache(max_tokens=64)
        gen = torch.Generator().manual_seed(1234)
        text = M.generate(prompt, tokenizer, model, max_tokens=n_new, kv_cache=cache, config=config, generator=gen)
        assert text.startswith(prompt)
        new = re.findall(r"<s>|</s>|<unk>|w\d+", text[len(prompt):])
        assert len(new) == n_new                              # no EOS stop, one token per step
        assert cache.num_items() == ids0.shape[1] + n_new - 1     # the last sampled token is never fed back
        outs.append(text)
    assert outs[0] == outs[1]                                 # same seed, same draw
    assert capsys.readouterr().out.count(prompt) == 2         # streaming print of the prompt, then token by token
    # every sampled token survives the oracle's min-p filter (p >= 0.1 p_max) with slack for bf16 logits
    oc = O.KVCache()
    ids = ids0
    for t in tokenizer.convert_tokens_to_ids(re.findall(r"<s>|</s>|<unk>|w\d+", outs[0][len(prompt):])):
        logits = O.forward(cfg, wn, ids, oc).logits[0, -1]
        p = np.exp(logits - logits.max())
        p /= p.sum()
        assert p[t] >= 0.05 * p.max(), (t, p[t], p.max())
        ids = np.array([[t]])
    # hf_faithful picks the rope scaling up from config.json: different logits past the original context
    _, model_hf, _ = M.load_model(str(tmp_path), hf_faithful=True)
    long_ids = np.random.default_rng(0).integers(3, cfg.vocab_size, (1, 200))
    a = model(long_ids, use_cache=False)[1].cpu().numpy()
    b = model_hf(long_ids, use_cache=False)[1].cpu().numpy()
    ex = O.Extras(always_causal=True, rope_scaling=dict(config.rope_scaling))
    assert np.abs(b - O.forward(cfg, wn, long_ids, extras=ex).logits).max() <= 0.15 * b.std()
    assert np.abs(a - b).max() > 0.15 * b.std()
