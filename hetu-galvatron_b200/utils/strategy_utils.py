"""Strategy string codec, e.g. ``"2-2*-2f-c"``  <->  ``[pp, tp, dp, {tp, fsdp, cpt, sp}]``.

Same grammar as ``galvatron/utils/strategy_utils.py:3-63``: ``pp-tp[*]-dp[f][*][-c][-sp]`` where ``*``
marks which of tp/dp is rank-consecutive, ``f`` = sharded DP, ``c`` = checkpoint, ``sp`` = Ulysses.
"""


def form_strategy(strategy):
    if len(strategy) != 4:
        raise ValueError("strategy must be [pp, tp, dp, info]")
    pp, tp, dp, info = strategy
    tp_s, dp_s = str(int(tp)), str(int(dp))
    if info.get("fsdp"):
        dp_s += "f"
    if "tp" in info:
        if info["tp"]:
            tp_s += "*"
        else:
            dp_s += "*"
    if info.get("cpt"):
        dp_s += "-c"
    if info.get("sp"):
        dp_s += "-sp"
    return f"{int(pp)}-{tp_s}-{dp_s}"


def strategy_str2list(text):
    parts = text.split("-")
    tp_consec = None
    if parts[1].endswith("*"):
        tp_consec, parts[1] = 1, parts[1][:-1]
    elif parts[2].endswith("*"):
        tp_consec, parts[2] = 0, parts[2][:-1]
    fsdp = 0
    if parts[2].endswith("f"):
        fsdp, parts[2] = 1, parts[2][:-1]
    tail = parts[3:]
    cpt = 1 if tail[:1] == ["c"] else 0
    sp = 1 if "sp" in tail[:2] else 0
    pp, tp, dp = int(parts[0]), int(parts[1]), int(parts[2])
    info = {}
    if tp > 1 and dp > 1:
        info["tp"] = tp_consec
    if dp > 1:
        info["fsdp"] = fsdp
    if cpt:
        info["cpt"] = 1
    if sp:
        info["sp"] = 1
    return [pp, tp, dp, info]
