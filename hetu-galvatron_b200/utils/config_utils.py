"""Strategy-JSON codec: the file format the Galvatron Search Engine emits and the runtime loads.

Mirrors ``galvatron/utils/config_utils.py:22-57`` (``config2strategy`` / ``strategy2config``) and the
``str2array`` / ``array2str`` helpers (:9-13).  One deliberate difference: the reference
loader requires ``cp_sizes_enc`` (:37) although neither the Search Engine (``search_engine.py:651-661``)
nor any shipped JSON writes it -- here a missing ``cp_sizes_enc`` means "all ones" and a missing ``vcp``
means 1, which is what makes the Search Engine's output loadable as-is.
"""
import json

STRATEGY_KEYS = ("pp_deg", "tp_sizes_enc", "tp_consecutive_flags", "dp_types_enc", "use_sp")


def str2array(s):
    if isinstance(s, (list, tuple)):
        return [int(v) for v in s]
    s = str(s).strip()
    return [int(tok) for tok in s.split(",")] if s else []


def array2str(a):
    return ",".join(str(int(v)) for v in a)


def read_json_config(path):
    with open(path, "r", encoding="utf-8") as fp:
        return json.load(fp)


def write_json_config(config, path):
    with open(path, "w") as fp:
        json.dump(config, fp, indent=4)


def config2strategy(config):
    """JSON dict -> (pp_deg, tp_sizes_enc, cp_sizes_enc, tp_consecutive_flags, dp_types_enc, use_sp, vtp, vsp, vcp).

    Same tuple order as the reference (config_utils.py:44)."""
    tp = str2array(config["tp_sizes_enc"])
    n = len(tp)
    cp = str2array(config["cp_sizes_enc"]) if "cp_sizes_enc" in config else [1] * n
    consec = str2array(config["tp_consecutive_flags"]) if "tp_consecutive_flags" in config else [1] * n
    dp_types = str2array(config["dp_types_enc"])
    use_sp = str2array(config["use_sp"]) if "use_sp" in config else [0] * n
    for name, lst in (("cp_sizes_enc", cp), ("tp_consecutive_flags", consec), ("dp_types_enc", dp_types), ("use_sp", use_sp)):
        if len(lst) != n:
            raise ValueError(f"strategy JSON: {name} has {len(lst)} entries, tp_sizes_enc has {n}")
    return (int(config["pp_deg"]), tp, cp, consec, dp_types, use_sp,
            int(config.get("vtp", 1)), int(config.get("vsp", 0)), int(config.get("vcp", 1)))


def strategy2config(strategy_list):
    """Search-engine strategy list ``[[pp, tp, dp, {flags}], ...]`` -> JSON dict (config_utils.py:46-57)."""
    if len(strategy_list) == 0:
        return {}

    def flag(info, key):
        return bool(info.get(key, False))

    infos = [s[-1] for s in strategy_list]
    return {
        "pp_deg": strategy_list[0][0],
        "tp_sizes_enc": array2str(s[1] for s in strategy_list),
        # an explicit tp:0 is the only way to say "strided TP"
        "tp_consecutive_flags": array2str(0 if ("tp" in i and not i["tp"]) else 1 for i in infos),
        "dp_types_enc": array2str(1 if flag(i, "fsdp") else 0 for i in infos),
        "use_sp": array2str(1 if flag(i, "sp") else 0 for i in infos),
    }
