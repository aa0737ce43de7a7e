"""args/JSON -> hybrid_parallel_configs, bit-exact vs the reference's get_hybrid_parallel_configs_api
(goldens: tests/golden/hp_config_api.json from oracle/gen_golden_groups.py) + strategy codec round trips."""
import copy
import json
import os
import types

import pytest

from hetu_galvatron_b200.core.runtime import world
from hetu_galvatron_b200.core.runtime.hybrid_parallel_config import (check_hp_config, get_chunks,
                                                                     get_hybrid_parallel_configs_api,
                                                                     layer_shapes_dtypes_whole_model)
from hetu_galvatron_b200.utils import (array2str, config2strategy, form_strategy, str2array, strategy2config,
                                       strategy_str2list)

with open(os.path.join(os.path.dirname(__file__), "golden", "hp_config_api.json")) as _f:
    _GOLD = json.load(_f)["cases"]


def _info(n):
    return lambda config, args: types.SimpleNamespace(layernums=lambda: [n])


@pytest.mark.parametrize("case", _GOLD, ids=[f"{c['mode']}_{i}" for i, c in enumerate(_GOLD)])
def test_configs_api_bit_exact(case):
    if case["mode"] == "GLOBAL":
        args = types.SimpleNamespace(**copy.deepcopy(case["args"]))
    else:
        base = dict(local_rank=1, galvatron_config_path=copy.deepcopy(case["json"]), pp_deg=1, global_tp_deg=1,
                    global_cp_deg=1, sdp=0, global_checkpoint=0, use_ulysses=False, vocab_tp=1, vocab_cp=1, vocab_sp=0,
                    global_train_batch_size=32, chunks=-1, pipeline_type="gpipe", default_dp_type="ddp", embed_sdp=0,
                    distributed_checkpoint=False, load=None, mixed_precision="bf16")
        args = types.SimpleNamespace(**base)
    with world.simulated(0, case["world"]):
        got = get_hybrid_parallel_configs_api(None, args, _info(case["layers"]))
        chunks = get_chunks(args)
    assert got == case["result"]
    assert chunks == case["chunks"]
    for key, val in case["args_written"].items():
        assert getattr(args, key) == val, key
    if case["mode"] == "JSON":
        # the search engine's JSON has no cp_sizes_enc (search_engine.py:651-661): must load as all-ones
        assert list(config2strategy(case["json"])) == case["config2strategy"]
        with world.simulated(0, case["world"]):
            check_hp_config(got, [case["layers"]])


def test_codec_round_trip():
    strategies = [[2, 2, 2, {"tp": 1, "fsdp": 1, "cpt": 1}], [2, 4, 1, {"sp": 1}], [2, 1, 4, {"fsdp": 0}],
                  [2, 2, 2, {"tp": 0, "fsdp": 0}]]
    for s in strategies:
        assert strategy_str2list(form_strategy(s)) == s
    cfg = strategy2config(strategies)
    assert cfg == {"pp_deg": 2, "tp_sizes_enc": "2,4,1,2", "tp_consecutive_flags": "1,1,1,0",
                   "dp_types_enc": "1,0,0,0", "use_sp": "0,1,0,0"}
    pp, tp, cp, consec, dpt, use_sp, vtp, vsp, vcp = config2strategy(cfg)
    assert (pp, tp, cp, consec, dpt, use_sp, vtp, vsp, vcp) == (2, [2, 4, 1, 2], [1] * 4, [1, 1, 1, 0], [1, 0, 0, 0],
                                                                 [0, 1, 0, 0], 1, 0, 1)
    assert str2array(array2str([3, 1, 2])) == [3, 1, 2]
    assert form_strategy([1, 2, 4, {"tp": 1, "fsdp": 1, "cpt": 1, "sp": 1}]) == "1-2*-4f-c-sp"


def test_layer_count_mismatch_raises():
    js = {"pp_deg": 1, "tp_sizes_enc": "1,1", "tp_consecutive_flags": "1,1", "dp_types_enc": "0,0", "global_bsz": 8,
          "chunks": 1}
    args = types.SimpleNamespace(local_rank=1, galvatron_config_path=js, vocab_tp=1, vocab_cp=1, vocab_sp=0,
                                 pipeline_type="gpipe", default_dp_type="ddp", embed_sdp=0, chunks=1,
                                 global_train_batch_size=8, pp_deg=1)
    with world.simulated(0, 2), pytest.raises(ValueError):
        get_hybrid_parallel_configs_api(None, args, _info(3))


def test_shapes_dtypes_whole():
    mt = ["embed"] + ["gpt_dec"] * 3 + ["norm", "cls"]
    shapes, dtypes = layer_shapes_dtypes_whole_model(mt, [3], [[[8, -1, 4]]], [["bf16"]])
    assert shapes == [None, [[8, -1, 4]], [[8, -1, 4]], [[8, -1, 4]], None, None]
    assert dtypes == [None, ["bf16"], ["bf16"], ["bf16"], None, None]
