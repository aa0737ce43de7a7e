"""Checkpoint format compatibility (galvatron/models/llama_hf/LlamaModel_checkpoint.py): a HuggingFace Llama
checkpoint converted BY THE REFERENCE'S OWN h2g tool (tests/golden/ckpt_llama_tiny/, made by oracle/gen_golden_ckpt.py) must
load into the hybrid-parallel model bit-exactly at any tensor-parallel / sharding layout and reproduce HF's loss (5e-3, the
reference's criterion tests/models/test_model_correctness.py:111-117); the distributed layout written by save_llama_module
must have the reference's directory / key structure and load back to the same weights."""
import json
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_host_runtime import launch  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_llama_tiny")
EXPECTED = json.load(open(os.path.join(GOLDEN, "expected.json")))

CASES = {
    "tp1": (1, dict()),
    "tp2": (2, dict(global_tp_deg=2, vocab_tp=2)),
    "tp2_megatron_sp": (2, dict(global_tp_deg=2, vocab_tp=2, sequence_parallel=True)),
    "dp2_zero3": (2, dict(sdp=1, embed_sdp=1)),
    "pp2": (2, dict(pp_deg=2, chunks=2)),
    # vocabulary rows tp 1 (data parallel over both ranks), layers tp 2: every layer row is wrapped in Module_with_relocation
    "vtp1_layers_tp2": (2, dict(_strategy_json={"pp_deg": 1, "tp_sizes_enc": "2,2", "tp_consecutive_flags": "1,1", "dp_types_enc": "0,0",
                                                "use_sp": "0,0", "checkpoint": "0,0", "global_bsz": 4, "chunks": 1,
                                                "default_dp_type": "zero2", "vtp": 1, "vsp": 0})),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_load_hf_layered_checkpoint(name):
    world, over = CASES[name]
    rep = launch(world, dict(over, load=GOLDEN, _golden_ckpt=GOLDEN))
    assert rep["ckpt_tensors_bit_exact"] == 21          # embed, norm, head + 2 layers x 9 HF tensors
    assert abs(rep["loss"] - EXPECTED["hf_loss_fp32"]) <= 5e-3 * EXPECTED["hf_loss_fp32"], rep


@pytest.mark.parametrize("name", ["tp2", "dp2_zero3", "vtp1_layers_tp2"])
def test_save_then_load_distributed_checkpoint(name, tmp_path):
    world, over = CASES[name]
    out = str(tmp_path / "ckpt")
    launch(world, dict(over, load=GOLDEN, _golden_ckpt=GOLDEN, _save_to=out, save=out))
    # the reference's layout (LlamaModel_checkpoint.py:156-216)
    assert os.path.isfile(os.path.join(out, "hybrid_parallel_configs.json"))
    it = os.path.join(out, "iter_0")
    assert os.path.isfile(os.path.join(it, "opt_param_scheduler.json"))
    tp = 2 if name in ("tp2", "vtp1_layers_tp2") else 1
    for d in ("model_embed_tokens", "model_layers_0", "model_layers_1", "model_norm", "lm_head"):
        n_files = tp if (name != "vtp1_layers_tp2" or d.startswith("model_layers")) else 1
        assert sorted(os.listdir(os.path.join(it, d))) == ["%d.pt" % r for r in range(n_files)], d
    assert sorted(os.listdir(os.path.join(it, "optimizer"))) == ["%d.pt" % r for r in range(world)]
    layer = torch.load(os.path.join(it, "model_layers_1", "0.pt"), weights_only=True)
    assert sorted(layer) == ["attention.LayerNorm.weight", "attention.attention.dense.weight",
                             "attention.attention.query_key_value.weight", "mlp.LayerNorm.weight",
                             "mlp.mlp.dense_4h_to_h.weight", "mlp.mlp.dense_h_to_4h.weight"]
    assert layer["attention.attention.query_key_value.weight"].dtype == torch.float32
    assert layer["attention.attention.query_key_value.weight"].shape == ((4 + 2 * 2) * 32 // tp, 128)
    assert list(torch.load(os.path.join(it, "model_embed_tokens", "0.pt"), weights_only=True)) == ["embed_tokens.weight"]
    # ... and it loads back (load_distributed_checkpoint :27-45) to HF's weights
    rep = launch(world, dict(over, load=out, distributed_checkpoint=True, _golden_ckpt=GOLDEN))
    assert rep["ckpt_tensors_bit_exact"] == 21


@pytest.mark.parametrize("name,optimizer", [("dp2_zero3", "torch"), ("dp2_zero3", "fused"), ("tp2", "torch")])
def test_resume_is_bit_identical(name, optimizer, tmp_path):
    """Train 2 steps, save (weights + per-rank optimizer state + scheduler, core/runtime/utils.py:152-165), train a third; a
    fresh job that loads the checkpoint and trains that third step must see the same loss to the last bit."""
    world, over = CASES[name]
    over = dict(over, fused_optimizer=optimizer == "fused")
    out = str(tmp_path / "ckpt")
    a = launch(world, dict(over, load=GOLDEN, save=out, _save_to=out, _save_after=2, _iters=3))
    assert len(a["losses"]) == 3 and a["losses"][2] != a["losses"][0]
    b = launch(world, dict(over, load=out, distributed_checkpoint=True, load_iteration=2, _skip_batches=2, _iters=1))
    assert b["losses"][0] == a["losses"][2], (a["losses"], b["losses"])
