#!/usr/bin/env python
"""Golden-vector generator for the strategy -> process-group mapping (TEST INFRASTRUCTURE ONLY).

Runs the UNMODIFIED reference code from /root/reference under a fake process group and
dumps what it produced, so the product's own group builder / config expansion can be held
bit-exact against it:

  * galvatron/core/runtime/comm_groups.py:416      gen_comm_groups   (16-tuple of groups)
  * galvatron/core/runtime/hybrid_parallel_config.py:232  hp_config_whole_model
  * galvatron/core/runtime/hybrid_parallel_config.py:17   get_hybrid_parallel_configs_api
  * galvatron/utils/config_utils.py:22              config2strategy

Only runs in the build container (needs /root/reference); the emitted JSON under
tests/golden/ is what travels.  Usage:

    python oracle/gen_golden_groups.py          # rewrites tests/golden/*.json

The reference needs apex/amp_C/dropout_layer_norm at import time; oracle/ref_shim/ holds
4-symbol stand-ins (nothing on the group/config path touches them).
"""
import copy
import json
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "ref_shim"), REF, os.path.join(REF, "galvatron", "site_package")]
warnings.filterwarnings("ignore")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import galvatron.core.runtime.comm_groups as ref_cg  # noqa: E402
import galvatron.core.runtime.hybrid_parallel_config as ref_hc  # noqa: E402
import galvatron.core as ref_core  # noqa: E402
from galvatron.utils.config_utils import config2strategy as ref_config2strategy  # noqa: E402


class _FakeWorld:
    rank = 0
    world = 1


def _install_fake_pg():
    dist.get_rank = lambda group=None: _FakeWorld.rank
    dist.get_world_size = lambda group=None: _FakeWorld.world
    dist.new_group = lambda ranks=None, **kw: ("fake_pg", tuple(ranks))
    # hp_config_whole_model prints on local_rank 0 via galvatron.core.get_args
    ref_core.get_args = lambda: types.SimpleNamespace(local_rank=1)


def _ranks(g):
    return None if g is None else list(g.ranks)


GROUP_KEYS = [
    "pp_group", "tp_groups", "sp_groups", "cp_groups", "dp_groups", "seq_data_groups",
    "allgather_tp_sp_groups", "split_tp_sp_groups", "allgather_cp_groups", "split_cp_groups",
    "allgather_tp_sp_cp_groups", "split_tp_sp_cp_groups", "fused_allgather_groups",
    "fused_split_groups", "embedding_group", "vtp_data_group",
]


def run_ref_groups(world, pp, tp, sp, cp, consec):
    """Per-rank outputs of the reference gen_comm_groups, as plain rank lists."""
    out = []
    for rank in range(world):
        _FakeWorld.rank, _FakeWorld.world = rank, world
        stdout = sys.stdout
        sys.stdout = open(os.devnull, "w")
        try:
            res = ref_cg.gen_comm_groups(list(tp), list(sp), list(cp), pp, list(consec), show_rank=-1)
        finally:
            sys.stdout.close()
            sys.stdout = stdout
        rec = {}
        for key, val in zip(GROUP_KEYS, res):
            rec[key] = [_ranks(g) for g in val] if isinstance(val, list) else _ranks(val)
        out.append(rec)
    return out


def whole(world, pp, tp_enc, use_sp=None, cp_enc=None, consec=None, vtp=1, vsp=0, vcp=1,
          embed_sdp=0, dp_types=None, ckpt=None, pp_division=None, family="llama"):
    """Build hp_configs and expand with the reference's hp_config_whole_model."""
    L = len(tp_enc)
    use_sp = use_sp or [0] * L
    cp_enc = cp_enc or [1] * L
    consec = consec or [1] * L
    dp_types = dp_types or [0] * L
    ckpt = ckpt or [0] * L
    if pp_division is None:
        avg = L // pp
        pp_division = [avg] * (pp - 1) + [L - avg * (pp - 1)]
    hp = {
        "pp_deg": pp, "tp_sizes_enc": list(tp_enc), "tp_consecutive_flags": list(consec),
        "cp_sizes_enc": list(cp_enc), "dp_types_enc": list(dp_types), "checkpoint_flags_enc": list(ckpt),
        "pp_ranks_enc": ref_hc.get_pp_ranks_enc(pp_division), "pp_division": list(pp_division),
        "use_sp": list(use_sp),
    }
    if family == "bert":
        module_types = ["embed"] + ["bert_enc"] * L + ["mlm_head"]
    else:
        module_types = ["embed"] + ["gpt_dec"] * L + ["norm", "cls"]
    _FakeWorld.rank, _FakeWorld.world = 0, world
    hp_whole = ref_hc.hp_config_whole_model(module_types, copy.deepcopy(hp), embed_sdp=embed_sdp, embed_ckpt=0,
                                            vocab_tp=vtp, vocab_sp=vsp, vocab_cp=vcp)
    return module_types, hp, hp_whole


def strategy_corpus():
    """(name, world, kwargs-for-whole) tuples.  Sources cited per block."""
    C = []
    # uniform grids, world 1/2/4/8: every (pp, tp|sp, cp) that divides
    for world in (1, 2, 4, 8):
        for pp in (1, 2, 4, 8):
            if world % pp:
                continue
            per_stage = world // pp
            for t in (1, 2, 4, 8):
                if per_stage % t:
                    continue
                for c in (1, 2, 4):
                    if (per_stage // t) % c:
                        continue
                    L = max(pp, 2)
                    for ulysses in (0, 1):
                        if ulysses and t == 1:
                            continue
                        for vt in sorted({1, t}):
                            C.append((f"uniform_w{world}_pp{pp}_{'sp' if ulysses else 'tp'}{t}_cp{c}_vtp{vt}", world,
                                      dict(pp=pp, tp_enc=[t] * L, use_sp=[ulysses] * L, cp_enc=[c] * L,
                                           vtp=vt, vsp=ulysses if vt > 1 else 0, vcp=c)))
    # tests/core/test_hybrid.py:122-183 (4 JSON strategies, world 8)
    hyb = [
        dict(pp=1, tp_enc=[1, 2, 4, 8], dp_types=[0, 1, 0, 1], use_sp=[0, 1, 0, 1], ckpt=[0, 0, 1, 1], pp_division=[4], vtp=2, vsp=0),
        dict(pp=1, tp_enc=[1, 2, 4, 8], dp_types=[1, 0, 1, 0], use_sp=[0, 1, 0, 1], ckpt=[0, 0, 1, 1], pp_division=[4], vtp=4, vsp=1),
        dict(pp=2, tp_enc=[1, 2, 4, 2], dp_types=[0, 1, 0, 1], use_sp=[0, 1, 0, 1], ckpt=[0, 0, 1, 1], pp_division=[3, 1], vtp=2, vsp=0),
        dict(pp=2, tp_enc=[1, 2, 4, 2], dp_types=[1, 0, 1, 0], use_sp=[0, 1, 0, 1], ckpt=[0, 0, 1, 1], pp_division=[2, 2], vtp=4, vsp=1),
    ]
    for i, kw in enumerate(hyb):
        C.append((f"test_hybrid_{i}", 8, kw))
    # tests/core/test_redistributed.py:141-145 (per-layer tp lists x vocab_tp), both SP flavours
    for tp, vt in (([1, 2, 4, 8], 8), ([2, 8, 2, 1], 4), ([8, 4, 1, 2], 2)):
        C.append((f"test_redistributed_tp{''.join(map(str, tp))}", 8, dict(pp=1, tp_enc=tp, vtp=vt)))
        C.append((f"test_redistributed_sp{''.join(map(str, tp))}", 8, dict(pp=1, tp_enc=tp, use_sp=[1] * 4, vtp=vt, vsp=1)))
    # SURVEY.md 8(c) mixed known answer: pp=2, whole-model tp [2,1,2,4,2,2,2] (vtp=2 rows + 4 layers)
    C.append(("survey_mixed_pp2", 8, dict(pp=2, tp_enc=[1, 2, 4, 2], vtp=2, pp_division=[2, 2])))
    # cp changes between layers (exercises the swapped split_cp/allgather_cp unpack, comm_groups.py:306 vs :483)
    C.append(("cp_varying", 8, dict(pp=1, tp_enc=[2, 1, 2, 1], cp_enc=[1, 2, 2, 4], vtp=1, vcp=1)))
    C.append(("cp_sp_mixed", 8, dict(pp=1, tp_enc=[2, 2, 1, 4], use_sp=[1, 0, 0, 1], cp_enc=[2, 2, 4, 1], vtp=2, vsp=1, vcp=2)))
    C.append(("cp_pp2", 8, dict(pp=2, tp_enc=[1, 2, 2, 1], cp_enc=[2, 2, 1, 4], vtp=2, vcp=1)))
    # BASELINE.json configs (2)-(5), full depth
    C.append(("baseline2_llama8b_sdp8", 8, dict(pp=1, tp_enc=[1] * 32, dp_types=[1] * 32)))
    C.append(("baseline2_llama8b_tp2_pp2", 8, dict(pp=2, tp_enc=[2] * 32, dp_types=[0] * 32, vtp=2)))
    C.append(("baseline3_gpt67b_pp2_tp2_zero2", 8, dict(pp=2, tp_enc=[2] * 32, pp_division=[16, 16], vtp=2)))
    C.append(("baseline4_bert_ulysses4_dp2", 8, dict(pp=1, tp_enc=[4] * 24, use_sp=[1] * 24, vtp=4, vsp=1, family="bert")))
    C.append(("baseline5_llama70b_zero3_ckpt", 8, dict(pp=1, tp_enc=[1] * 80, dp_types=[1] * 80, ckpt=[1] * 80)))
    # shipped example JSONs are 2-node (world 16): keep one for the >8 mapping
    C.append(("world16_pp2_tp_mixed", 16, dict(pp=2, tp_enc=[8, 4, 2, 1, 1, 2, 4, 8], vtp=8)))
    C.append(("world16_pp4_sp", 16, dict(pp=4, tp_enc=[4, 2, 4, 2], use_sp=[1, 1, 0, 0], vtp=2)))
    return C


def gen_groups_golden():
    cases = []
    for name, world, kw in strategy_corpus():
        module_types, hp, hp_whole = whole(world, **kw)
        per_rank = run_ref_groups(world, hp_whole["pp_deg"], hp_whole["tp_sizes_whole"], hp_whole["sp_sizes_whole"],
                                  hp_whole["cp_sizes_whole"], hp_whole["tp_consec_whole"])
        cases.append({
            "name": name, "world": world, "module_types": module_types, "hp_configs": hp,
            "embed_sdp": kw.get("embed_sdp", 0), "vocab_tp": kw.get("vtp", 1), "vocab_sp": kw.get("vsp", 0),
            "vocab_cp": kw.get("vcp", 1), "hp_configs_whole": hp_whole, "groups_per_rank": per_rank,
        })
    return cases


def gen_config_api_golden():
    """get_hybrid_parallel_configs_api in GLOBAL and JSON mode + config2strategy."""
    out = []

    def base_args(**kw):
        a = dict(local_rank=1, galvatron_config_path=None, pp_deg=1, global_tp_deg=1, global_cp_deg=1, sdp=0,
                 global_checkpoint=0, use_ulysses=False, vocab_tp=1, vocab_cp=1, vocab_sp=0,
                 global_train_batch_size=32, chunks=-1, pipeline_type="gpipe", default_dp_type="ddp",
                 embed_sdp=0, distributed_checkpoint=False, load=None, mixed_precision="bf16")
        a.update(kw)
        return a

    def run(world, nlayers, args_kw):
        _FakeWorld.rank, _FakeWorld.world = 0, world
        args = types.SimpleNamespace(**base_args(**args_kw))
        info = lambda config, a: types.SimpleNamespace(layernums=lambda: [nlayers])  # noqa: E731
        res = ref_hc.get_hybrid_parallel_configs_api(None, args, info)
        chunks = ref_hc.get_chunks(args)
        return res, {k: getattr(args, k) for k in
                     ("vocab_tp", "vocab_sp", "vocab_cp", "pp_deg", "chunks", "global_train_batch_size",
                      "pipeline_type", "default_dp_type", "embed_sdp")}, chunks

    global_cases = [
        (8, 4, dict()),
        (8, 4, dict(pp_deg=2, global_tp_deg=2, sdp=1, global_checkpoint=1, default_dp_type="zero2", global_train_batch_size=64)),
        (8, 7, dict(pp_deg=2, global_tp_deg=4, use_ulysses=True, vocab_tp=4)),
        (8, 32, dict(pp_deg=4, global_tp_deg=2, global_cp_deg=1, vocab_tp=2, global_train_batch_size=16, pipeline_type="pipedream_flush")),
        (8, 6, dict(global_tp_deg=2, global_cp_deg=2, vocab_tp=2, vocab_cp=2)),
        (4, 5, dict(pp_deg=4, global_train_batch_size=8, chunks=3)),
        (1, 12, dict(global_train_batch_size=8)),
        (2, 3, dict(global_tp_deg=2, vocab_tp=2, global_train_batch_size=4)),
    ]
    for world, nl, kw in global_cases:
        res, wr, chunks = run(world, nl, kw)
        out.append({"mode": "GLOBAL", "world": world, "layers": nl, "args": base_args(**kw), "result": res,
                    "args_written": wr, "chunks": chunks})

    json_cases = [
        (8, {"pp_deg": 1, "tp_sizes_enc": "1,2,4,8", "tp_consecutive_flags": "1,1,1,1", "dp_types_enc": "0,1,0,1",
             "use_sp": "0,1,0,1", "checkpoint": "0,0,1,1", "global_bsz": 32, "chunks": 2, "pp_division": "4",
             "pipeline_type": "pipedream_flush", "default_dp_type": "zero2", "vtp": 2, "vsp": 0}),
        (8, {"pp_deg": 2, "tp_sizes_enc": "1,2,4,2", "tp_consecutive_flags": "1,1,1,1", "dp_types_enc": "1,0,1,0",
             "use_sp": "0,1,0,1", "checkpoint": "0,0,1,1", "global_bsz": 32, "chunks": 4, "pp_division": "2,2",
             "pipeline_type": "pipedream_flush", "default_dp_type": "zero2", "vtp": 4, "vsp": 1}),
        (8, {"pp_deg": 2, "tp_sizes_enc": "2,2,2,2,2", "tp_consecutive_flags": "1,1,1,1,1", "dp_types_enc": "0,0,0,0,0",
             "global_bsz": 16, "chunks": 8}),
        (8, {"pp_deg": 1, "tp_sizes_enc": "2,2,4,4", "cp_sizes_enc": "2,2,1,1", "tp_consecutive_flags": "1,1,1,1",
             "dp_types_enc": "0,0,1,1", "global_bsz": 8, "chunks": 1, "vtp": 2, "vcp": 2, "embed_sdp": 1}),
    ]
    for world, js in json_cases:
        js_ref = dict(js)
        nl = len(js["tp_sizes_enc"].split(","))
        # HEAD's loader requires cp_sizes_enc (config_utils.py:37); the search engine never writes it
        js_ref.setdefault("cp_sizes_enc", ",".join(["1"] * nl))
        res, wr, chunks = run(world, nl, dict(galvatron_config_path=js_ref))
        out.append({"mode": "JSON", "world": world, "layers": nl, "json": js, "result": res, "args_written": wr,
                    "chunks": chunks, "config2strategy": list(ref_config2strategy(js_ref))})
    return out


def main():
    _install_fake_pg()
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    cases = gen_groups_golden()
    with open(os.path.join(gold, "comm_groups.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden_groups.py", "reference": "PKU-DAIR/Hetu-Galvatron v2.4.1 (14f59728)",
                   "group_keys": GROUP_KEYS, "cases": cases}, f, separators=(",", ":"))
    api = gen_config_api_golden()
    with open(os.path.join(gold, "hp_config_api.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden_groups.py", "cases": api}, f, indent=1)
    print("cases:", len(cases), "api cases:", len(api))


if __name__ == "__main__":
    main()
