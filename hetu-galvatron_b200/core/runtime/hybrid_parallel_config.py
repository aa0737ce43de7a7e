"""args / strategy JSON -> per-layer hybrid-parallel config (integer plane of the drop-in boundary).

Mirrors ``galvatron/core/runtime/hybrid_parallel_config.py``: ``get_hybrid_parallel_configs_api`` (:17),
``ModelInfo`` (:161), ``check_hp_config`` (:190), ``hp_config_whole_model`` (:232),
``layer_shapes_dtypes_whole_model`` (:322), ``get_chunks`` (:351), ``get_pp_ranks_enc`` (:9).  Results are
held bit-exact against the reference by ``tests/test_hp_config.py`` (goldens made by running the reference).

Differences, all from SURVEY 8(g): a JSON without ``cp_sizes_enc``/``vcp`` loads (all-ones), a layer-count
mismatch between JSON and model raises (the reference's ``assert (False, "...")`` :78 never fires), and
``check_hp_config`` also validates ``cp_sizes_enc``/``use_sp``.
"""
import json
import math
import os

from ...utils import config2strategy, read_json_config, str2array
from . import world as _world


def get_pp_ranks_enc(pp_divide):
    return [stage for stage, n_layers in enumerate(pp_divide) for _ in range(n_layers)]


def _default_pp_division(total_layer_num, pp_deg):
    """Even split, remainder on the last stage (:85-88)."""
    avg = total_layer_num // pp_deg
    return [avg] * (pp_deg - 1) + [total_layer_num - avg * (pp_deg - 1)]


def get_hybrid_parallel_configs_api(config, args, model_info):
    world_size = _world.get_world_size()
    json_mode = args.galvatron_config_path not in (None, "None")
    total_layer_num = sum(model_info(config, args).layernums())

    if not json_mode:  # GLOBAL mode (:26-42)
        pp_deg = args.pp_deg
        tp_sizes_enc = [args.global_tp_deg if args.global_tp_deg > 0 else 1] * total_layer_num
        tp_consecutive_flags = [1] * total_layer_num
        global_cp = getattr(args, "global_cp_deg", 1)
        cp_sizes_enc = [global_cp if global_cp > 0 else 1] * total_layer_num
        dp_types_enc = [args.sdp] * total_layer_num
        checkpoint_flags_enc = [args.global_checkpoint] * total_layer_num
        pp_divide = None
        args.vocab_sp = 1 if args.use_ulysses else 0
        use_sp = [args.vocab_sp] * total_layer_num
        if not hasattr(args, "vocab_cp"):
            args.vocab_cp = 1
    else:  # JSON mode (:43-84)
        src = args.galvatron_config_path
        galvatron_config = read_json_config(src) if isinstance(src, str) else src
        pp_deg, tp_sizes_enc, cp_sizes_enc, tp_consecutive_flags, dp_types_enc, use_sp, vtp, vsp, vcp = \
            config2strategy(galvatron_config)
        if len(tp_sizes_enc) != total_layer_num:
            raise ValueError("Layer_num in json config (%d) does not match layer_num in the model (%d)!"
                             % (len(tp_sizes_enc), total_layer_num))
        checkpoint_flags_enc = (str2array(galvatron_config["checkpoint"]) if "checkpoint" in galvatron_config
                                else [0] * total_layer_num)
        pp_divide = str2array(galvatron_config["pp_division"]) if "pp_division" in galvatron_config else None
        for key in ("pipeline_type", "default_dp_type", "embed_sdp"):
            if key in galvatron_config:
                setattr(args, key, galvatron_config[key])
        args.global_train_batch_size = galvatron_config["global_bsz"]
        args.chunks = galvatron_config["chunks"]
        args.pp_deg, args.vocab_tp, args.vocab_sp, args.vocab_cp = pp_deg, vtp, vsp, vcp

    if pp_divide is None:
        pp_divide = _default_pp_division(total_layer_num, pp_deg)
    if len(pp_divide) != pp_deg or sum(pp_divide) != total_layer_num:
        raise ValueError("pp_division %s does not describe %d layers on %d stages" % (pp_divide, total_layer_num, pp_deg))
    pp_ranks_enc = get_pp_ranks_enc(pp_divide)
    min_tp = min(min(tp_sizes_enc), args.vocab_tp)
    min_cp = min(min(cp_sizes_enc), args.vocab_cp)
    assert args.global_train_batch_size % (world_size // pp_deg // min_tp // min_cp) == 0, \
        "global_train_batch_size should be multiple of world_size//pp_deg//min_tp//min_cp!"

    hybrid_parallel_configs = {
        "pp_deg": pp_deg,
        "tp_sizes_enc": tp_sizes_enc,
        "tp_consecutive_flags": tp_consecutive_flags,
        "cp_sizes_enc": cp_sizes_enc,
        "dp_types_enc": dp_types_enc,
        "checkpoint_flags_enc": checkpoint_flags_enc,
        "pp_ranks_enc": pp_ranks_enc,
        "pp_division": pp_divide,
        "use_sp": use_sp,
        "vocab_tp": args.vocab_tp,
        "vocab_sp": args.vocab_sp,
        "vocab_cp": args.vocab_cp,
        "default_dp_type": args.default_dp_type,
        "global_train_batch_size": args.global_train_batch_size,
    }

    if getattr(args, "distributed_checkpoint", False):  # strategy-equality check on resume (:112-124)
        with open(os.path.join(args.load, "hybrid_parallel_configs.json"), "r") as fp:
            saved = json.load(fp)
        assert hybrid_parallel_configs.keys() == saved.keys(), \
            "Hybrid parallel configs are not equal, %s vs %s" % (hybrid_parallel_configs.keys(), saved.keys())
        for key, val in hybrid_parallel_configs.items():
            assert val == saved[key], f"Hybrid parallel configs are not equal for key {key}, {val} vs {saved[key]}"

    if getattr(args, "local_rank", 1) == 0:
        print("======================== Galvatron Parallel Config =============================")
        print("Galvatron parallel config mode: [%s config mode]" % ("JSON" if json_mode else "GLOBAL"))
        print("   global_batch_size: %d, chunks: %d, pp_deg: %d" % (args.global_train_batch_size, args.chunks, pp_deg))
        print("   pipeline_type: %s, default_dp_type: %s, dtype: %s%s" % (
            args.pipeline_type, args.default_dp_type, getattr(args, "mixed_precision", "bf16"),
            ", embed_sdp: 1" if args.embed_sdp else ""))
        print_hp_configs(hybrid_parallel_configs)
    return hybrid_parallel_configs


class ModelInfo:
    """What a model family tells the core about itself (:161-187)."""

    def __init__(self):
        self.layernum_list = self.layer_shapes_list = self.layer_dtypes_list = self.layer_module_types = None

    def set_layernums(self, info):
        self.layernum_list = info

    def set_shapes(self, info):
        self.layer_shapes_list = info

    def set_dtypes(self, info):
        self.layer_dtypes_list = info

    def set_module_types(self, info):
        self.layer_module_types = info

    def layernums(self):
        return self.layernum_list

    def shapes(self):
        return self.layer_shapes_list

    def dtypes(self):
        return self.layer_dtypes_list

    def module_types(self):
        return self.layer_module_types


def check_hp_config(hp_configs, layernum_list):
    total = sum(layernum_list)
    pp_deg = hp_configs["pp_deg"]
    per_layer = ("tp_sizes_enc", "tp_consecutive_flags", "dp_types_enc", "pp_ranks_enc", "checkpoint_flags_enc",
                 "cp_sizes_enc", "use_sp")
    for key in per_layer:
        assert key in hp_configs and len(hp_configs[key]) == total, \
            "hp_configs[%r] must have one entry per layer (%d)" % (key, total)
    per_stage = _world.get_world_size() // pp_deg
    for tp, cp in zip(hp_configs["tp_sizes_enc"], hp_configs["cp_sizes_enc"]):
        assert 1 <= tp <= per_stage and per_stage % tp == 0, "Wrong tp_size!"
        assert cp >= 1 and per_stage % (tp * cp) == 0, "Wrong cp_size!"
    assert all(f in (0, 1) for f in hp_configs["tp_consecutive_flags"]), "Wrong tp_consec!"
    assert all(d in (0, 1, None) for d in hp_configs["dp_types_enc"]), "Wrong dp_type!"
    assert all(0 <= r <= pp_deg - 1 for r in hp_configs["pp_ranks_enc"]), "Wrong pp_rank!"
    assert all(c in (0, 1) for c in hp_configs["checkpoint_flags_enc"]), "Wrong checkpoint_flag!"
    assert all(u in (0, 1) for u in hp_configs["use_sp"]), "Wrong use_sp!"


def print_hp_config(key, val):
    if isinstance(val, (list, tuple)):
        print("   " + key + ":" + " " * max(28 - len(key), 0), val)


def print_hp_configs(hp_configs):
    for key, val in hp_configs.items():
        print_hp_config(key, val)
    print("================================================================================")


def _is_transformer_layer(module_type):
    return module_type[-3:] in ("enc", "dec")


def hp_config_whole_model(module_types, hp_configs, embed_sdp=0, embed_ckpt=0, vocab_tp=1, vocab_sp=0, vocab_cp=1):
    """Expand per-transformer-layer lists to one row per whole-model module (embed / layers / norm / cls).

    A layer with ``use_sp==1`` re-interprets its ``tp_sizes_enc`` entry as the Ulysses degree (:261-266);
    non-layer rows take the vocab_* degrees, ``embed_sdp``, and the pp stage of the neighbouring layer
    (:273-287).  ``dp_sizes_whole = world/pp/tp/sp/cp`` (:290-293)."""
    cols = {k: [] for k in ("tp_sizes_whole", "sp_sizes_whole", "cp_sizes_whole", "tp_consec_whole", "dp_types_whole",
                            "pp_ranks_whole", "checkpoint_flags_whole")}
    pp_ranks_enc = hp_configs["pp_ranks_enc"]
    cursor = 0  # next transformer layer
    for module_type in module_types:
        if _is_transformer_layer(module_type):
            degree, ulysses = hp_configs["tp_sizes_enc"][cursor], hp_configs["use_sp"][cursor] == 1
            row = (1 if ulysses else degree, degree if ulysses else 1, hp_configs["cp_sizes_enc"][cursor],
                   hp_configs["tp_consecutive_flags"][cursor], hp_configs["dp_types_enc"][cursor],
                   pp_ranks_enc[cursor], hp_configs["checkpoint_flags_enc"][cursor])
            cursor += 1
        else:
            ulysses = vocab_sp == 1
            row = (1 if ulysses else vocab_tp, vocab_tp if ulysses else 1, vocab_cp, 1, embed_sdp,
                   pp_ranks_enc[cursor] if cursor < len(pp_ranks_enc) else pp_ranks_enc[-1], embed_ckpt)
        for key, val in zip(("tp_sizes_whole", "sp_sizes_whole", "cp_sizes_whole", "tp_consec_whole", "dp_types_whole",
                             "pp_ranks_whole", "checkpoint_flags_whole"), row):
            cols[key].append(val)

    whole = {"pp_deg": hp_configs["pp_deg"]}
    whole.update(cols)
    stage_ranks = _world.get_world_size() // hp_configs["pp_deg"]
    whole["dp_sizes_whole"] = [stage_ranks // t // s // c for t, s, c in
                               zip(cols["tp_sizes_whole"], cols["sp_sizes_whole"], cols["cp_sizes_whole"])]
    return whole


def get_enc_groups(groups_whole, module_types):
    assert len(groups_whole) == len(module_types)
    return [g for g, m in zip(groups_whole, module_types) if _is_transformer_layer(m)]


def mixed_precision_dtype(mixed_precision):
    import torch
    return {"fp32": torch.float, "fp16": torch.float16, "bf16": torch.bfloat16}[mixed_precision]


def layer_shapes_dtypes_whole_model(module_types, layernum_list, layer_shapes_list, layer_dtypes_list):
    """Stage-boundary tensor shapes/dtypes per whole-model row (:322-348): a non-layer row sitting before the
    first / after the last transformer layer has none; one in between inherits the next layer's."""
    assert len(layernum_list) == len(layer_shapes_list) == len(layer_dtypes_list)
    shapes_enc = [shape for n, shape in zip(layernum_list, layer_shapes_list) for _ in range(n)]
    dtypes_enc = [dt for n, dt in zip(layernum_list, layer_dtypes_list) for _ in range(n)]
    shapes_whole, dtypes_whole, cursor = [], [], 0
    for module_type in module_types:
        is_layer = ("enc" in module_type) or ("dec" in module_type)
        if not is_layer and cursor in (0, len(shapes_enc)):
            shapes_whole.append(None)
            dtypes_whole.append(None)
            continue
        shapes_whole.append(shapes_enc[cursor])
        dtypes_whole.append(dtypes_enc[cursor])
        cursor += 1 if is_layer else 0
    return shapes_whole, dtypes_whole


def get_chunks(args):
    """chunks == -1 -> heuristic: 1 without PP, else ceil(local_bsz / 4) microbatches (:351-361)."""
    if args.chunks == -1:
        args.chunks = 1
        if args.pp_deg > 1:
            local_bsz = args.global_train_batch_size // (_world.get_world_size() // args.pp_deg)
            args.chunks = max(int(math.ceil(local_bsz / 4)), 1)
    return args.chunks
