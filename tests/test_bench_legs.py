"""bench.py's path legs (the fixed-strategy runs of the collectives north_star names) before they meet a GPU: every leg's
Galvatron strategy must expand under the reference's config rules at its world size, and the tiny-model version of the same
strategy -- the one whose parity verdict the leg reports -- must reproduce the oracle on the CPU (gloo) backend."""
import importlib.util
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_host_runtime import launch  # noqa: E402


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _family(bench, leg):
    fam = bench.family_of(leg["model"])
    if fam == "llama":
        from hetu_galvatron_b200.llama_hf import config_from_meta, set_model_config
        from hetu_galvatron_b200.llama_hf.LlamaModel_hybrid_parallel import estimate_arena_bytes, get_hybrid_parallel_configs
        from hetu_galvatron_b200.llama_hf.meta_configs import _SPECS
        seq_key = "n_positions"
    elif fam == "gpt":
        from hetu_galvatron_b200.gpt_hf import config_from_meta, set_model_config
        from hetu_galvatron_b200.gpt_hf.GPTModel_hybrid_parallel import estimate_arena_bytes, get_hybrid_parallel_configs
        from hetu_galvatron_b200.gpt_hf.meta_configs import _SPECS
        seq_key = "n_positions"
    else:
        from hetu_galvatron_b200.bert_hf import config_from_meta, set_model_config
        from hetu_galvatron_b200.bert_hf.BertModel_hybrid_parallel import estimate_arena_bytes, get_hybrid_parallel_configs
        from hetu_galvatron_b200.bert_hf.meta_configs import _SPECS
        seq_key = "max_position_embeddings"
    return fam, config_from_meta, set_model_config, estimate_arena_bytes, get_hybrid_parallel_configs, dict(_SPECS[leg["model"]], **{seq_key: leg.get("seq", 8192)})


@pytest.mark.parametrize("n", [2, 4, 8])
def test_leg_strategies_expand(n):
    from hetu_galvatron_b200.core.runtime import world as _world
    from hetu_galvatron_b200.core.runtime.arguments import initialize_galvatron
    bench = _bench()
    legs = bench.leg_catalog(n)
    assert len(legs) >= 7                      # the five collectives legs + BASELINE configs 3 and 4 (+ config 5 at N = 8)
    for name, leg in legs.items():
        s = dict(leg["strategy"])
        fam, config_from_meta, set_model_config, estimate_arena_bytes, get_hybrid_parallel_configs, spec = _family(bench, leg)
        with _world.simulated(0, n):
            args = initialize_galvatron(galvatron_config_path=s, mixed_precision="bf16", fused_optimizer=True,
                                        sequence_parallel=bool(s.get("sequence_parallel", 0)), use_ulysses=False, vocab_tp=s.get("vtp", 1),
                                        default_dp_type=s["default_dp_type"], chunks=s["chunks"], global_train_batch_size=s["global_bsz"],
                                        pp_deg=s["pp_deg"], make_vocab_size_divisible_by=128)
            args.vocab_size = spec["vocab_size"]
            config = set_model_config(config_from_meta(spec), args)
            hp = get_hybrid_parallel_configs(config, args)
            assert len(hp["tp_sizes_enc"]) == config.num_hidden_layers, name
            arena = estimate_arena_bytes(config, args, hp)
            assert 0 < arena < 120 << 30, (name, arena)


@pytest.mark.parametrize("n", [2, 4])
def test_leg_tiny_strategies_match_the_oracle(n):
    from test_families import launch as launch_family
    bench = _bench()
    for name, leg in bench.leg_catalog(n).items():
        if leg["tiny"] is None:
            continue
        tiny = json.loads(json.dumps(leg["tiny"]))
        tiny.pop("_env", None)
        rep = (launch_family if "_family" in tiny else launch)(n, tiny)
        assert rep["max_grad_err"] < 3e-2, (name, rep)
