"""Pin the GPT / BERT oracle (oracle/gpt_bert_ref.py, Megatron weight layout) against HF ``GPT2LMHeadModel`` and
``BertForMaskedLM`` in fp64 -- the baselines of the reference's own model tests (tests/models/test_model_correctness.py:111-117)."""
import pytest
import torch

from oracle import gpt_bert_ref as ref

CFG = dict(hidden=64, ffn=256, n_heads=4, head_dim=16, n_layers=2, vocab=256, seq=32, eps=1e-5)


def _randomize(w, seed):
    g = torch.Generator().manual_seed(seed)
    for name, t in list(w.items()):
        if torch.is_tensor(t) and t.dim() == 1:
            w[name] = t + 0.1 * torch.randn(t.shape, generator=g, dtype=t.dtype)
    for lw in w["layers"]:
        for k in lw:
            if lw[k].dim() == 1:
                lw[k] = lw[k] + 0.1 * torch.randn(lw[k].shape, generator=g, dtype=lw[k].dtype)
    leaves = [t for t in w.values() if torch.is_tensor(t)] + [t for lw in w["layers"] for t in lw.values()]
    for t in leaves:
        t.requires_grad_(True)
    return w


def test_gpt_oracle_matches_hf():
    transformers = pytest.importorskip("transformers")
    w = _randomize(ref.init_weights(CFG, "gpt", seed=3, std=0.05, dtype=torch.float64), 1)
    conf = transformers.GPT2Config(n_embd=CFG["hidden"], n_layer=CFG["n_layers"], n_head=CFG["n_heads"], vocab_size=CFG["vocab"],
                                   n_positions=CFG["seq"], n_inner=CFG["ffn"], layer_norm_epsilon=CFG["eps"], resid_pdrop=0.0, embd_pdrop=0.0,
                                   attn_pdrop=0.0, activation_function="gelu_new", tie_word_embeddings=False, attn_implementation="eager")
    hf = transformers.GPT2LMHeadModel(conf).double()
    sd = {"transformer.wte.weight": w["wte"], "transformer.wpe.weight": w["wpe"], "transformer.ln_f.weight": w["norm"],
          "transformer.ln_f.bias": w["norm_b"], "lm_head.weight": w["lm_head"]}
    for i, p in enumerate(w["layers"]):
        (q, k, v), (qb, kb, vb) = ref.split_qkv(p, CFG)
        pre = "transformer.h.%d." % i
        sd.update({pre + "ln_1.weight": p["ln1"], pre + "ln_1.bias": p["ln1_b"], pre + "ln_2.weight": p["ln2"], pre + "ln_2.bias": p["ln2_b"],
                   pre + "attn.c_attn.weight": torch.cat([q, k, v]).t(), pre + "attn.c_attn.bias": torch.cat([qb, kb, vb]),       # Conv1D: [in, out]
                   pre + "attn.c_proj.weight": p["dense"].t(), pre + "attn.c_proj.bias": p["dense_b"],
                   pre + "mlp.c_fc.weight": p["h_to_4h"].t(), pre + "mlp.c_fc.bias": p["h_to_4h_b"],
                   pre + "mlp.c_proj.weight": p["4h_to_h"].t(), pre + "mlp.c_proj.bias": p["4h_to_h_b"]})
    missing, unexpected = hf.load_state_dict({k: v.detach() for k, v in sd.items()}, strict=False)
    assert not [k for k in missing if "attn.bias" not in k and "masked_bias" not in k] and not unexpected, (missing, unexpected)
    tokens = torch.randint(0, CFG["vocab"], (3, 24))
    labels = torch.randint(0, CFG["vocab"], (3, 24))
    per_tok, loss = ref.gpt_forward_loss(w, tokens, labels, CFG, dtype=torch.float64)
    loss.backward()
    logits = hf(input_ids=tokens).logits
    hf_tok = torch.nn.functional.cross_entropy(logits.reshape(-1, CFG["vocab"]), labels.reshape(-1), reduction="none").view(3, 24)
    torch.testing.assert_close(per_tok.detach(), hf_tok.detach(), rtol=2e-6, atol=2e-6)
    hf_tok.mean().backward()
    torch.testing.assert_close(w["lm_head"].grad, hf.lm_head.weight.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(w["wpe"].grad, hf.transformer.wpe.weight.grad, rtol=1e-5, atol=1e-8)
    l0 = hf.transformer.h[0]
    torch.testing.assert_close(w["layers"][0]["dense"].grad, l0.attn.c_proj.weight.grad.t(), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(w["layers"][0]["h_to_4h_b"].grad, l0.mlp.c_fc.bias.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(w["layers"][0]["ln1_b"].grad, l0.ln_1.bias.grad, rtol=1e-5, atol=1e-8)
    n, hn = CFG["n_heads"], CFG["head_dim"]
    gq = w["layers"][0]["qkv"].grad.view(n, 3, hn, -1)
    torch.testing.assert_close(gq[:, 1].reshape(n * hn, -1), l0.attn.c_attn.weight.grad.t()[n * hn:2 * n * hn], rtol=1e-5, atol=1e-8)


def test_bert_oracle_matches_hf():
    transformers = pytest.importorskip("transformers")
    cfg = dict(CFG, gelu_tanh=False)          # HF BERT: exact GeLU
    w = _randomize(ref.init_weights(cfg, "bert", seed=5, std=0.05, dtype=torch.float64), 2)
    conf = transformers.BertConfig(hidden_size=cfg["hidden"], num_hidden_layers=cfg["n_layers"], num_attention_heads=cfg["n_heads"],
                                   intermediate_size=cfg["ffn"], vocab_size=cfg["vocab"], max_position_embeddings=cfg["seq"], type_vocab_size=2,
                                   layer_norm_eps=cfg["eps"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, hidden_act="gelu",
                                   tie_word_embeddings=False, attn_implementation="eager")
    hf = transformers.BertForMaskedLM(conf).double()
    sd = {"bert.embeddings.word_embeddings.weight": w["word"], "bert.embeddings.position_embeddings.weight": w["pos"],
          "bert.embeddings.token_type_embeddings.weight": w["type"], "bert.embeddings.LayerNorm.weight": w["emb_ln"],
          "bert.embeddings.LayerNorm.bias": w["emb_ln_b"], "cls.predictions.transform.dense.weight": w["transform"],
          "cls.predictions.transform.dense.bias": w["transform_b"], "cls.predictions.transform.LayerNorm.weight": w["transform_ln"],
          "cls.predictions.transform.LayerNorm.bias": w["transform_ln_b"], "cls.predictions.decoder.weight": w["decoder"],
          "cls.predictions.decoder.bias": w["decoder_b"], "cls.predictions.bias": w["decoder_b"]}
    for i, p in enumerate(w["layers"]):
        (q, k, v), (qb, kb, vb) = ref.split_qkv(p, cfg)
        pre = "bert.encoder.layer.%d." % i
        sd.update({pre + "attention.self.query.weight": q, pre + "attention.self.query.bias": qb, pre + "attention.self.key.weight": k,
                   pre + "attention.self.key.bias": kb, pre + "attention.self.value.weight": v, pre + "attention.self.value.bias": vb,
                   pre + "attention.output.dense.weight": p["dense"], pre + "attention.output.dense.bias": p["dense_b"],
                   pre + "attention.output.LayerNorm.weight": p["ln1"], pre + "attention.output.LayerNorm.bias": p["ln1_b"],
                   pre + "intermediate.dense.weight": p["h_to_4h"], pre + "intermediate.dense.bias": p["h_to_4h_b"],
                   pre + "output.dense.weight": p["4h_to_h"], pre + "output.dense.bias": p["4h_to_h_b"],
                   pre + "output.LayerNorm.weight": p["ln2"], pre + "output.LayerNorm.bias": p["ln2_b"]})
    have = set(hf.state_dict())
    missing, unexpected = hf.load_state_dict({k: v.detach() for k, v in sd.items() if k in have}, strict=False)
    assert not [k for k in missing if "position_ids" not in k], missing
    b, s = 3, 24
    tokens = torch.randint(0, cfg["vocab"], (b, s))
    labels = torch.randint(0, cfg["vocab"], (b, s))
    mask = torch.ones(b, s, dtype=torch.bool)
    mask[0, 18:] = False
    mask[2, 10:] = False
    tt = (torch.arange(s)[None] >= 9).long().expand(b, s)
    per_tok, loss = ref.bert_forward_loss(w, tokens, labels, cfg, dtype=torch.float64, attention_mask=mask, token_type_ids=tt)
    logits = hf(input_ids=tokens, attention_mask=mask.long(), token_type_ids=tt).logits
    hf_tok = torch.nn.functional.cross_entropy(logits.reshape(-1, cfg["vocab"]), labels.reshape(-1), reduction="none").view(b, s)
    valid = mask            # rows whose query is a padding position are unconstrained in both implementations' definitions: compare the rest
    torch.testing.assert_close(per_tok.detach()[valid], hf_tok.detach()[valid], rtol=2e-6, atol=2e-6)
    (per_tok * valid).sum().backward()
    (hf_tok * valid).sum().backward()
    torch.testing.assert_close(w["decoder"].grad, hf.cls.predictions.decoder.weight.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(w["type"].grad, hf.bert.embeddings.token_type_embeddings.weight.grad, rtol=1e-5, atol=1e-8)
    l0 = hf.bert.encoder.layer[0]
    torch.testing.assert_close(w["layers"][0]["dense"].grad, l0.attention.output.dense.weight.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(w["layers"][0]["ln2_b"].grad, l0.output.LayerNorm.bias.grad, rtol=1e-5, atol=1e-8)
    n, hn = cfg["n_heads"], cfg["head_dim"]
    gq = w["layers"][0]["qkv"].grad.view(n, 3, hn, -1)
    torch.testing.assert_close(gq[:, 0].reshape(n * hn, -1), l0.attention.self.query.weight.grad, rtol=1e-5, atol=1e-8)


def test_out_of_vocabulary_label_follows_megatron():
    """The -100 labels of BERT's unmasked positions: megatron's vocab-parallel CE subtracts the row maximum, then gives such a target
    the (shifted) predicted logit 0 (cross_entropy.py:22-60): the position contributes logsumexp - max."""
    logits = torch.randn(4, 2, 16, dtype=torch.float64)
    labels = torch.tensor([[1, -100], [3, 5], [-100, 15], [0, 2]])
    got = ref._token_loss(logits, labels, torch.float64)
    assert torch.allclose(got[0, 1], torch.logsumexp(logits[0, 1], -1) - logits[0, 1].max())
    assert torch.allclose(got[1, 0], torch.logsumexp(logits[1, 0], -1) - logits[1, 0, 3])


def test_gpt_oracle_with_tied_embeddings_matches_hf():
    """HF GPT-2's default: lm_head IS wte.  The oracle expresses it as one leaf used twice (what tests/_family_worker.py does for the
    tied runs of the product): its gradient must be HF's gradient of the shared matrix (embedding part + head part)."""
    transformers = pytest.importorskip("transformers")
    w = _randomize(ref.init_weights(CFG, "gpt", seed=5, std=0.05, dtype=torch.float64), 2)
    w["lm_head"] = w["wte"]
    conf = transformers.GPT2Config(n_embd=CFG["hidden"], n_layer=CFG["n_layers"], n_head=CFG["n_heads"], vocab_size=CFG["vocab"],
                                   n_positions=CFG["seq"], n_inner=CFG["ffn"], layer_norm_epsilon=CFG["eps"], resid_pdrop=0.0, embd_pdrop=0.0,
                                   attn_pdrop=0.0, activation_function="gelu_new", tie_word_embeddings=True, attn_implementation="eager")
    hf = transformers.GPT2LMHeadModel(conf).double()
    sd = {k: v.detach() for k, v in ref.to_hf_state_dict(w, CFG, "gpt").items()}
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "attn.bias" not in k and "masked_bias" not in k] and not unexpected, (missing, unexpected)
    assert hf.lm_head.weight is hf.transformer.wte.weight
    tokens = torch.randint(0, CFG["vocab"], (3, 24))
    labels = torch.randint(0, CFG["vocab"], (3, 24))
    _, loss = ref.gpt_forward_loss(w, tokens, labels, CFG, dtype=torch.float64)
    loss.backward()
    logits = hf(input_ids=tokens).logits
    hf_loss = torch.nn.functional.cross_entropy(logits.reshape(-1, CFG["vocab"]), labels.reshape(-1))
    torch.testing.assert_close(loss.detach(), hf_loss.detach(), rtol=2e-6, atol=2e-6)
    hf_loss.backward()
    torch.testing.assert_close(w["wte"].grad, hf.transformer.wte.weight.grad, rtol=1e-5, atol=1e-8)
