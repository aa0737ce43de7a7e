#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle/): generates tests/golden/ckpt_llama_tiny/ -- a tiny HuggingFace Llama checkpoint converted to
Galvatron's layer-wise format BY THE REFERENCE'S OWN TOOL (galvatron/tools/checkpoint_convert_h2g.py:47-87,
``convert_checkpoints_llama``), plus the HF model's loss on a fixed token batch.  Runs only in the build container (needs
/root/reference and transformers); the output is committed.  The product's loader
(hetu-galvatron_b200/llama_hf/LlamaModel_checkpoint.py) must reproduce HF's weights bit-exactly at any tensor-parallel degree
and HF's loss within the reference's 5e-3.

    python oracle/gen_golden_ckpt.py
"""
import importlib.util
import json
import os
import shutil
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TOOL = "/root/reference/galvatron/tools/checkpoint_convert_h2g.py"
OUT = os.path.join(ROOT, "tests", "golden", "ckpt_llama_tiny")
# the TINY spec of tests/smoke_model.py
SPEC = dict(hidden_size=128, intermediate_size=352, num_attention_heads=4, num_key_value_heads=2, num_hidden_layers=2,
            rms_norm_eps=1e-5, vocab_size=512, max_position_embeddings=64, rope_theta=10000.0)


def main():
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(20240921)
    cfg = LlamaConfig(**SPEC, tie_word_embeddings=False, attention_bias=False, mlp_bias=False)
    cfg._attn_implementation = "eager"
    model = LlamaForCausalLM(cfg).float().eval()
    with torch.no_grad():
        for p in model.parameters():            # bf16-exact values: the files are stored in bf16 (1 MB) without losing a bit
            p.copy_(p.to(torch.bfloat16).float())
        for name, p in model.named_parameters():
            if name.endswith("norm.weight"):    # non-trivial norm weights (HF initialises them to 1)
                p.copy_((1.0 + 0.1 * torch.randn_like(p)).to(torch.bfloat16).float())
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, SPEC["vocab_size"], (4, SPEC["max_position_embeddings"] + 1), generator=g)
    tokens, labels = x[:, :-1], x[:, 1:]
    with torch.no_grad():
        logits = model(tokens).logits.float()
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1))

    spec = importlib.util.spec_from_file_location("ref_h2g", REF_TOOL)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    src = tempfile.mkdtemp()
    torch.save({k: v.to(torch.bfloat16) for k, v in model.state_dict().items()}, os.path.join(src, "pytorch_model.bin"))
    shutil.rmtree(OUT, ignore_errors=True)
    ref.convert_checkpoints_llama(src, OUT)          # the reference's converter, unmodified
    shutil.rmtree(src)
    meta = {"spec": SPEC, "token_seed": 11, "batch": [4, SPEC["max_position_embeddings"] + 1], "hf_loss_fp32": float(loss),
            "generator": "oracle/gen_golden_ckpt.py", "converter": "galvatron/tools/checkpoint_convert_h2g.py:convert_checkpoints_llama",
            "files": sorted(os.listdir(OUT))}
    json.dump(meta, open(os.path.join(OUT, "expected.json"), "w"), indent=2)
    print(json.dumps(meta, indent=2))
    print("bytes:", sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
