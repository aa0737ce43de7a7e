"""Pin the oracle: oracle/llama_ref.py (restatement of the reference's Llama layer math in Megatron weight layout) must
agree with HF ``LlamaForCausalLM`` -- the very baseline the reference's own GPU tests compare against
(tests/core/test_tp.py:60-121, HF->Galvatron weight split of galvatron/tools/checkpoint_convert_h2g.py:43-82)."""
import pytest
import torch

from oracle import llama_ref

CFG = dict(hidden=64, ffn=176, n_heads=4, n_kv_heads=2, head_dim=16, n_layers=2, vocab=256, eps=1e-5, rope_base=10000.0)


def _hf_model(weights):
    transformers = pytest.importorskip("transformers")
    conf = transformers.LlamaConfig(hidden_size=CFG["hidden"], intermediate_size=CFG["ffn"], num_attention_heads=CFG["n_heads"],
                                    num_key_value_heads=CFG["n_kv_heads"], num_hidden_layers=CFG["n_layers"],
                                    vocab_size=CFG["vocab"], rms_norm_eps=CFG["eps"], max_position_embeddings=64,
                                    rope_theta=CFG["rope_base"], attention_dropout=0.0, tie_word_embeddings=False,
                                    attn_implementation="eager")
    model = transformers.LlamaForCausalLM(conf).double()
    missing, unexpected = model.load_state_dict(llama_ref.to_hf_state_dict(weights, CFG), strict=False)
    assert not [k for k in missing if "rotary" not in k] and not unexpected
    return model


def test_oracle_matches_hf_forward_and_grads():
    torch.manual_seed(0)
    w = llama_ref.init_weights(CFG, seed=3, std=0.05, dtype=torch.float64)
    for lw in w["layers"]:
        lw["ln1"] = lw["ln1"] + 0.1 * torch.randn_like(lw["ln1"])
        lw["ln2"] = lw["ln2"] + 0.1 * torch.randn_like(lw["ln2"])
    leaves = [w["embed"], w["norm"], w["lm_head"]] + [t for lw in w["layers"] for t in lw.values()]
    for t in leaves:
        t.requires_grad_(True)
    tokens = torch.randint(0, CFG["vocab"], (3, 24))
    labels = torch.randint(0, CFG["vocab"], (3, 24))
    per_tok, loss = llama_ref.forward_loss(w, tokens, labels, CFG, dtype=torch.float64)
    loss.backward()

    hf = _hf_model({k: (v.detach() if torch.is_tensor(v) else [{kk: vv.detach() for kk, vv in lw.items()} for lw in v])
                    for k, v in w.items()})
    logits = hf(input_ids=tokens).logits                                         # [b, s, V]
    hf_tok = torch.nn.functional.cross_entropy(logits.reshape(-1, CFG["vocab"]), labels.reshape(-1), reduction="none").view(3, 24)
    torch.testing.assert_close(per_tok.detach(), hf_tok.detach(), rtol=2e-6, atol=2e-6)
    hf_tok.mean().backward()
    torch.testing.assert_close(w["lm_head"].grad, hf.lm_head.weight.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(w["embed"].grad, hf.model.embed_tokens.weight.grad, rtol=1e-5, atol=1e-8)
    l0 = hf.model.layers[0]
    torch.testing.assert_close(w["layers"][0]["dense"].grad, l0.self_attn.o_proj.weight.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(w["layers"][0]["4h_to_h"].grad, l0.mlp.down_proj.weight.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(w["layers"][0]["ln1"].grad, l0.input_layernorm.weight.grad, rtol=1e-5, atol=1e-8)
    hn, ng, r = CFG["head_dim"], CFG["n_kv_heads"], CFG["n_heads"] // CFG["n_kv_heads"]
    gq = w["layers"][0]["qkv"].grad.view(ng, (r + 2) * hn, -1)
    torch.testing.assert_close(gq[:, :r * hn].reshape(ng * r * hn, -1), l0.self_attn.q_proj.weight.grad, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(gq[:, r * hn:(r + 1) * hn].reshape(ng * hn, -1), l0.self_attn.k_proj.weight.grad, rtol=1e-5, atol=1e-8)


def test_bf16_mode_tracks_exact():
    w = llama_ref.init_weights(CFG, seed=5, std=0.05)
    tokens = torch.randint(0, CFG["vocab"], (2, 16))
    labels = torch.randint(0, CFG["vocab"], (2, 16))
    _, exact = llama_ref.forward_loss(w, tokens, labels, CFG, dtype=torch.float32)
    _, bf = llama_ref.forward_loss(w, tokens, labels, CFG, dtype=torch.bfloat16)
    assert abs(float(exact) - float(bf)) < 5e-3 * float(exact)   # the reference's own test tolerance (test_tp.py:121)
