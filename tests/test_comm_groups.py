"""Strategy -> process-group mapping, bit-exact against the reference's own gen_comm_groups.

Goldens: tests/golden/comm_groups.json, produced by oracle/gen_golden_groups.py executing
galvatron/core/runtime/comm_groups.py:416 under a fake process group for every rank of 100 strategies
(uniform grids at world 1/2/4/8, the reference's test_hybrid / test_redistributed corpora, BASELINE configs 2-5).
"""
import copy
import json
import os

import pytest

from hetu_galvatron_b200.core.runtime import world
from hetu_galvatron_b200.core.runtime.comm_groups import CommGroup, gen_comm_groups
from hetu_galvatron_b200.core.runtime.hybrid_parallel_config import hp_config_whole_model

with open(os.path.join(os.path.dirname(__file__), "golden", "comm_groups.json")) as _f:
    _GOLD = json.load(_f)
KEYS = _GOLD["group_keys"]


def _ranks(g):
    return None if g is None else list(g.ranks)


@pytest.mark.parametrize("case", _GOLD["cases"], ids=[c["name"] for c in _GOLD["cases"]])
def test_mapping_bit_exact(case):
    W = case["world"]
    hpw = case["hp_configs_whole"]
    for rank in range(W):
        with world.simulated(rank, W):
            res = gen_comm_groups(list(hpw["tp_sizes_whole"]), list(hpw["sp_sizes_whole"]), list(hpw["cp_sizes_whole"]),
                                  hpw["pp_deg"], list(hpw["tp_consec_whole"]))
        assert len(res) == 16
        want = case["groups_per_rank"][rank]
        for key, val in zip(KEYS, res):
            got = [_ranks(g) for g in val] if isinstance(val, list) else _ranks(val)
            assert got == want[key], (case["name"], rank, key)


@pytest.mark.parametrize("case", _GOLD["cases"], ids=[c["name"] for c in _GOLD["cases"]])
def test_whole_model_expansion_bit_exact(case):
    with world.simulated(0, case["world"]):
        got = hp_config_whole_model(case["module_types"], copy.deepcopy(case["hp_configs"]), embed_sdp=case["embed_sdp"],
                                    embed_ckpt=0, vocab_tp=case["vocab_tp"], vocab_sp=case["vocab_sp"],
                                    vocab_cp=case["vocab_cp"])
    assert got == case["hp_configs_whole"]


def test_known_answers_from_survey():
    # cfg(3) PP2xTP2xSDP2, cfg(4) Ulysses SP4xDP2, cfg(5) SDP8, TP2xCP2xDP2 (world 8)
    def run(rank, tp, sp, cp, pp):
        n = len(tp)
        with world.simulated(rank, 8):
            return gen_comm_groups(tp, sp, cp, pp, [1] * n)
    r = run(5, [2] * 3, [1] * 3, [1] * 3, 2)
    assert r[0].ranks == [1, 5] and r[1][0].ranks == [4, 5] and r[4][0].ranks == [5, 7] and r[5][0].ranks == [5, 7]
    assert r[14].ranks == [1, 5] and r[15].ranks == [5, 7]
    r = run(6, [1] * 3, [4] * 3, [1] * 3, 1)
    assert r[2][0].ranks == [4, 5, 6, 7] and r[4][0].ranks == [2, 6] and r[5][0].ranks == list(range(8))
    r = run(3, [1] * 3, [1] * 3, [1] * 3, 1)
    assert r[4][0].ranks == list(range(8)) and r[1][0].ranks == [3]
    r = run(3, [2] * 3, [1] * 3, [2] * 3, 1)
    assert r[1][0].ranks == [2, 3] and r[3][0].ranks == [1, 3] and r[4][0].ranks == [3, 7] and r[5][0].ranks == [1, 3, 5, 7]


def test_consec_flag_rewritten_in_place_and_strided_tp_rejected():
    flags = [0, 0, 1]
    with world.simulated(0, 8):
        gen_comm_groups([1, 8, 2], [1, 1, 1], [1, 1, 1], 1, flags)
    assert flags == [1, 1, 1]  # comm_groups.py:427-428
    with world.simulated(0, 8), pytest.raises(ValueError):
        gen_comm_groups([2, 2], [1, 1], [1, 1], 1, [0, 1])
    with world.simulated(0, 8), pytest.raises(AssertionError):
        gen_comm_groups([2], [2], [1], 1, [1])  # Ulysses x Megatron-TP on one layer


def test_commgroup_value_semantics():
    g = CommGroup([3, 1, 1, 2])
    assert g.ranks == [1, 2, 3] and g.size == 3 and g.group is g
    assert g.has_rank(2) and g.intra_group_id == 1 and not g.has_rank(0)
    with world.simulated(2, 4):
        a = gen_comm_groups([2, 2], [1, 1], [1, 1], 1, [1, 1])
        b = gen_comm_groups([2, 2], [1, 1], [1, 1], 1, [1, 1])
    assert a[1][0] is b[1][1]  # interned: O(1) creation, one device handle per distinct rank list
