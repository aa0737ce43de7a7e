"""Which tensors are alive at the activation peak of one microbatch: 4 real-size Llama-3-8B layers, seq 8192, no checkpointing, one
forward (torch allocator history), dumped as a size/stack table.  One GPU; diagnostic only."""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    opts = bench.parse()
    opts.layers, opts.checkpoint_layers = 4, 0
    torch.cuda.set_device(0)
    _, strategy = bench.strategy_for(1, os.path.join(ROOT, "configs", "ncu_2layers_1microbatch.json"))
    for k in ("tp_sizes_enc", "tp_consecutive_flags", "dp_types_enc", "use_sp", "checkpoint"):
        strategy[k] = ",".join([strategy[k].split(",")[0]] * 4)
    strategy["pp_division"] = "4"
    args, config, model = bench.build_model(opts, strategy)
    from hetu_galvatron_b200.core.runtime.utils import get_optimizer_and_param_scheduler
    opt, _ = get_optimizer_and_param_scheduler(model, args)
    tokens = torch.randint(0, config.vocab_size, (1, config.max_position_embeddings), device="cuda")
    model.forward_backward([tokens], 0, None, loss_func=None, attention_mask=None, labels=tokens.clone())   # warm-up (buffers, workspaces)
    opt.step(); opt.zero_grad()
    torch.cuda.synchronize()
    base = torch.cuda.memory_allocated()
    torch.cuda.memory._record_memory_history(max_entries=200000, stacks="python")
    torch.cuda.reset_peak_memory_stats()
    model.forward_backward([tokens], 1, None, loss_func=None, attention_mask=None, labels=tokens.clone())
    torch.cuda.synchronize()
    snap = torch.cuda.memory._snapshot()
    torch.cuda.memory._record_memory_history(enabled=None)
    # replay the trace: the live set at the moment of peak allocation
    live, cur, peak, peak_live = {}, 0, 0, {}
    for tr in snap["device_traces"][0]:
        if tr["action"] == "alloc":
            frames = [f for f in tr.get("frames", []) if "hetu-galvatron_b200" in f["filename"] or "bench.py" in f["filename"]]
            where = " < ".join("%s:%d %s" % (os.path.basename(f["filename"]), f["line"], f["name"]) for f in frames[:3])
            live[tr["addr"]] = (tr["size"], where)
            cur += tr["size"]
            if cur > peak:
                peak, peak_live = cur, dict(live)
        elif tr["action"] in ("free_requested", "free"):
            if tr["addr"] in live and tr["action"] == "free_requested":
                cur -= live.pop(tr["addr"])[0]
    agg = collections.Counter()
    cnt = collections.Counter()
    for size, where in peak_live.values():
        agg[(size, where)] += size
        cnt[(size, where)] += 1
    rows = [{"MiB_each": round(k[0] / 2 ** 20, 2), "count": cnt[k], "MiB_total": round(v / 2 ** 20, 1), "where": k[1]} for k, v in agg.most_common(40)]
    out = {"layers": 4, "allocated_before_step_GiB": round(base / 2 ** 30, 3), "peak_extra_GiB": round(peak / 2 ** 30, 3),
           "torch_peak_GiB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 3), "live_at_peak": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_memory_snapshot_4layers.json"), "w"), indent=1)
    for r in rows[:30]:
        print("%9.2f MiB x %3d = %8.1f  %s" % (r["MiB_each"], r["count"], r["MiB_total"], r["where"][:150]))
    print("peak extra GiB", out["peak_extra_GiB"])


if __name__ == "__main__":
    sys.argv = [sys.argv[0]]
    main()
