#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle/): generates tests/golden/ckpt_gpt_tiny/ and tests/golden/ckpt_bert_tiny/ -- tiny HuggingFace GPT-2 and
BERT (masked-LM) checkpoints converted to Galvatron's layer-wise format BY THE REFERENCE'S OWN TOOL
(galvatron/tools/checkpoint_convert_h2g.py:6-41 ``convert_checkpoints_gpt``, :84-130 ``convert_checkpoints_bert_mlm``), plus the HF
models' losses on a fixed batch.  Runs only in the build container (needs /root/reference and transformers); the output is committed.
The families' loaders (hetu-galvatron_b200/{gpt_hf/GPTModel,bert_hf/BertModel}_checkpoint.py) must reproduce HF's weights bit-exactly at
any tensor-parallel degree and HF's loss within the reference's 5e-3 (tests/test_checkpoint_families.py).

    python oracle/gen_golden_ckpt_families.py
"""
import importlib.util
import json
import os
import shutil
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TOOL = "/root/reference/galvatron/tools/checkpoint_convert_h2g.py"
GPT = dict(n_layer=2, n_embd=128, n_head=4, vocab_size=512, n_positions=64)                      # tests/_family_worker.py TINY["gpt"]
BERT = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=4, vocab_size=512, max_position_embeddings=64, layer_norm_eps=1e-5)


def _bf16_exact(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1:                    # non-trivial norm weights and biases (HF initialises them to 1 / 0)
                base = 1.0 if name.endswith("weight") else 0.0
                p.copy_(base + 0.1 * torch.randn(p.shape, generator=g))
            p.copy_(p.to(torch.bfloat16).float())   # bf16-exact values: stored in bf16 without losing a bit


def _convert(fn_name, model, out):
    spec = importlib.util.spec_from_file_location("ref_h2g", REF_TOOL)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    src = tempfile.mkdtemp()
    torch.save({k: v.to(torch.bfloat16) if v.is_floating_point() else v for k, v in model.state_dict().items()}, os.path.join(src, "pytorch_model.bin"))
    shutil.rmtree(out, ignore_errors=True)
    getattr(ref, fn_name)(src, out)                 # the reference's converter, unmodified
    shutil.rmtree(src)


def gpt():
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(20240922)
    conf = GPT2Config(**GPT, n_inner=4 * GPT["n_embd"], layer_norm_epsilon=1e-5, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0,
                      activation_function="gelu_new", tie_word_embeddings=True, attn_implementation="eager")
    model = GPT2LMHeadModel(conf).float().eval()
    _bf16_exact(model, 1)
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, GPT["vocab_size"], (4, GPT["n_positions"] + 1), generator=g)
    with torch.no_grad():
        logits = model(x[:, :-1]).logits.float()
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), x[:, 1:].reshape(-1))
    out = os.path.join(ROOT, "tests", "golden", "ckpt_gpt_tiny")
    _convert("convert_checkpoints_gpt", model, out)
    meta = {"spec": GPT, "token_seed": 11, "batch": [4, GPT["n_positions"] + 1], "hf_loss_fp32": float(loss), "tied_lm_head": True,
            "generator": "oracle/gen_golden_ckpt_families.py", "converter": "galvatron/tools/checkpoint_convert_h2g.py:convert_checkpoints_gpt",
            "files": sorted(os.listdir(out))}
    json.dump(meta, open(os.path.join(out, "expected.json"), "w"), indent=2)
    print(json.dumps(meta, indent=2))


def bert():
    from transformers import BertConfig, BertForMaskedLM
    torch.manual_seed(20240923)
    conf = BertConfig(**BERT, intermediate_size=4 * BERT["hidden_size"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                      hidden_act="gelu_new", type_vocab_size=2, tie_word_embeddings=True, attn_implementation="eager")
    model = BertForMaskedLM(conf).float().eval()
    _bf16_exact(model, 2)
    g = torch.Generator().manual_seed(11)
    seq = BERT["max_position_embeddings"]
    x = torch.randint(0, BERT["vocab_size"], (4, seq + 1), generator=g)
    tokens = x[:, :-1]
    lengths = torch.randint(seq // 2, seq + 1, (4,), generator=g)           # the batch of tests/_family_worker.py (same generator order)
    mask = torch.arange(seq)[None, :] < lengths[:, None]
    tt = (torch.arange(seq)[None, :] >= (lengths[:, None] // 2)).long() * mask.long()
    labels = torch.where(torch.rand(4, seq, generator=g) < 0.15, tokens, torch.full_like(tokens, -100))
    labels = torch.where(mask, labels, torch.full_like(labels, -100))
    with torch.no_grad():
        logits = model(input_ids=tokens, attention_mask=mask.long(), token_type_ids=tt).logits.float()
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), ignore_index=-100)
    out = os.path.join(ROOT, "tests", "golden", "ckpt_bert_tiny")
    _convert("convert_checkpoints_bert_mlm", model, out)
    meta = {"spec": BERT, "token_seed": 11, "batch": [4, seq + 1], "hf_loss_fp32": float(loss), "generator": "oracle/gen_golden_ckpt_families.py",
            "converter": "galvatron/tools/checkpoint_convert_h2g.py:convert_checkpoints_bert_mlm", "files": sorted(os.listdir(out))}
    json.dump(meta, open(os.path.join(out, "expected.json"), "w"), indent=2)
    print(json.dumps(meta, indent=2))


if __name__ == "__main__":
    gpt()
    bert()
