#!/usr/bin/env python
"""bench.py -- tokens/sec of one Llama-3-8B training step under a Galvatron per-layer hybrid strategy on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]                      (N > 1: launched under torchrun)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]     (CPU restatement of the reference path)

One "step" = forward_backward over the global batch (chunks microbatches) + optimizer step, through the public API
(``llama_model_hp`` -> ``GalvatronModel.forward_backward``), on synthetic tokens of the reference's generator
(``DataLoaderForLlama``) and random-init weights.  Workload at every N: the strategy JSON ``configs/galvatron_config_
llama3-8b_<N>gpus.json`` (per-GPU batch fixed => weak scaling).  Prints ONE JSON line on rank 0.

  value      tokens/s with the step's tokens already resident in HBM (CUDA-event timed, max over ranks)
  e2e        the same loop with the tokens/labels copied from pinned host memory every step and the loss read back
  roofline   the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs / CUDA-event launch time vs the measured cuBLAS peak
  cpu_baseline  the oracle CPU restatement (oracle/gloo_backend.py) on a bounded sample, rank 0 at N=1 only
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tokens/sec Llama-3-8B auto-searched hybrid strategy at 1/2/4/8 B200 vs ref CPU"
MODEL = "llama3-8b"
SEQ = 8192
PER_GPU_BATCH = 8


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=4)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--model", default=MODEL)
    p.add_argument("--seq", type=int, default=SEQ)
    p.add_argument("--layers", type=int, default=0, help="debug only: truncate the model (the result is then marked invalid)")
    p.add_argument("--strategy", default=None, help="strategy JSON path (default: configs/ for this N)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                   help="fused: AdamW inside the gradient reduce-scatter kernel; torch: torch.optim.AdamW(fused=True) on fp32 grads")
    p.add_argument("--checkpoint-layers", type=int, default=-1, help="override: checkpoint the first k layers")
    return p.parse_args()


def strategy_for(n_gpus, path=None):
    path = path or os.path.join(ROOT, "configs", "galvatron_config_llama3-8b_%dgpus.json" % n_gpus)
    with open(path) as f:
        return path, json.load(f)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "MEASURED_PEAKS.json bf16_tflops_sustained (measured)"
    return 1400.0, "B200_PROFILING.md fallback (sustained ~1.4 PFLOP/s)"


def ncu_gemm_traffic(flops_per_launch_avg):
    """DRAM bytes per (average) GEMM launch from the committed `ncu --set full` capture: the capture holds three launches of
    known shape; their bytes-per-FLOP ratio is applied to the average launch of the timed region."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_gemm_full_summary.json")
    try:
        with open(path) as f:
            launches = json.load(f)["launches"]
        to_bytes = lambda s: float(s.split()[0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[s.split()[1]]  # noqa: E731
        total = sum(to_bytes(l["dram__bytes_read.sum"]) + to_bytes(l["dram__bytes_write.sum"]) for l in launches)
        # the three captured launches: gate/up dgrad [8192x28672]x[28672x4096], o-proj wgrad and dgrad [8192|4096 x 4096 x 4096|8192]
        flops = 2.0 * 8192 * 4096 * (28672 + 4096 + 4096)
        return round(total / flops * flops_per_launch_avg), ("dram__bytes_read+write of 3 captured launches (profiles/r01_ncu_gemm_full_summary.json) "
                                                            "scaled by FLOPs to the average launch of this run")
    except Exception:  # noqa: BLE001
        return None, "no ncu capture found"


# ---------------------------------------------------------------------------------------------------------------------
def build_model(opts, strategy, backend=None):
    import torch
    from hetu_galvatron_b200.core.runtime.arguments import initialize_galvatron
    from hetu_galvatron_b200.llama_hf import config_from_meta, llama_model_hp, set_model_config
    from hetu_galvatron_b200.llama_hf.meta_configs import _SPECS
    spec = dict(_SPECS[opts.model], n_positions=opts.seq)
    if opts.layers:
        spec["n_layers"] = opts.layers
        n = opts.layers
        for key in ("tp_sizes_enc", "tp_consecutive_flags", "dp_types_enc", "use_sp", "checkpoint", "cp_sizes_enc"):
            if key in strategy:
                strategy[key] = ",".join(strategy[key].split(",")[:n])
        if "pp_division" in strategy:
            pp = strategy["pp_deg"]
            strategy["pp_division"] = ",".join([str(n // pp)] * (pp - 1) + [str(n - n // pp * (pp - 1))])
    if getattr(opts, "checkpoint_layers", -1) >= 0:
        n_l = len(strategy["tp_sizes_enc"].split(","))
        strategy["checkpoint"] = ",".join(["1"] * min(opts.checkpoint_layers, n_l) + ["0"] * max(0, n_l - opts.checkpoint_layers))
    args = initialize_galvatron(galvatron_config_path=strategy, mixed_precision="bf16", fused_optimizer=getattr(opts, "optimizer", "torch") == "fused", sequence_parallel=bool(strategy.get("sequence_parallel", 0)),
                                use_ulysses=False, init_method_std=0.02, seed=1234, local_rank=1, lr=1e-4, adam_weight_decay=0.01,
                                make_vocab_size_divisible_by=128, vocab_tp=strategy.get("vtp", 1), model_size=opts.model,
                                default_dp_type=strategy.get("default_dp_type", "zero2"), chunks=strategy["chunks"],
                                global_train_batch_size=strategy["global_bsz"], pp_deg=strategy["pp_deg"])
    args.vocab_size = spec["vocab_size"]
    config = set_model_config(config_from_meta(spec), args)
    model = llama_model_hp(config, args)
    return args, config, model


def synthetic_batches(args, config, n_steps, dp_idx, dp_size, pin):
    """DataLoaderForLlama semantics (models/llama_hf/dataloader.py:52-80), seed 1234, this rank's data-parallel slice."""
    import numpy as np
    import torch
    rng = np.random.RandomState(1234)
    gbs, seq = args.global_train_batch_size, config.max_position_embeddings
    out = []
    for _ in range(n_steps):
        lengths = rng.randint(1, seq + 1, (gbs,))
        ids = rng.randint(0, config.vocab_size, (gbs, seq + 1))
        ids[np.arange(seq + 1)[None, :] >= lengths[:, None]] = 0
        lo, hi = dp_idx * gbs // dp_size, (dp_idx + 1) * gbs // dp_size
        x = torch.from_numpy(ids[lo:hi]).long()
        tokens, labels = x[:, :-1].contiguous(), x[:, 1:].contiguous()
        if pin:
            tokens, labels = tokens.pin_memory(), labels.pin_memory()
        out.append((tokens, labels))
    return out


def run_ours(opts):
    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")   # 150+ GiB of long-lived state: avoid fragmentation
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert world == opts.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (opts.gpus, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)   # bootstrap only (handles, barriers)
    from hetu_galvatron_b200.core.runtime.backend import get_backend
    from hetu_galvatron_b200.core.runtime.utils import get_optimizer_and_param_scheduler
    spath, strategy = strategy_for(world, opts.strategy)
    args, config, model = build_model(opts, strategy)
    be = get_backend()
    opt, _ = get_optimizer_and_param_scheduler(model, args)
    dp_group = model.vtp_data_group
    dp_idx, dp_size = dp_group.rank_in_group(rank), dp_group.size
    K, W = opts.steps, opts.warmup
    host = synthetic_batches(args, config, 2 * K + W, dp_idx, dp_size, pin=True)
    tokens_per_step = args.global_train_batch_size * config.max_position_embeddings

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step(tokens, labels, it):
        loss = model.forward_backward([tokens], it, None, loss_func=None, attention_mask=None, labels=labels)
        opt.step()
        opt.zero_grad()
        return loss

    it = 0
    for i in range(W):                                         # warm-up (untimed)
        t, l = host[i]
        step(t.to(dev, non_blocking=True), l.to(dev, non_blocking=True), it); it += 1

    def timed(resident):
        nonlocal it
        batches = host[W:W + K] if resident else host[W + K:W + 2 * K]
        if resident:
            batches = [(t.to(dev), l.to(dev)) for t, l in batches]
        sync_all()
        launches0 = be.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for t, l in batches:
            if not resident:
                t, l = t.to(dev, non_blocking=True), l.to(dev, non_blocking=True)
            last = step(t, l, it); it += 1                      # forward_backward returns the loss as a python float (D2H read)
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms[0])
        return ms, be.launch_count() - launches0, last

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    be.gemm_profile = []
    ms_res, launches, loss_res = timed(resident=True)
    prof, be.gemm_profile = be.gemm_profile, None
    ms_e2e, _, loss_e2e = timed(resident=False)
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()
    gemm_ms = sum(rec[0].elapsed_time(rec[1]) for rec in prof)
    gemm_flops = sum(rec[2] for rec in prof)
    gemm_bytes = sum(rec[3] for rec in prof)
    traffic, traffic_note = ncu_gemm_traffic(gemm_flops / max(1, len(prof)))
    peak, peak_src = measured_peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    h2d = 2 * (args.global_train_batch_size // dp_size) * config.max_position_embeddings * 8
    value = tokens_per_step * K / (ms_res * 1e-3)
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_res / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic tokens (DataLoaderForLlama generator, seed 1234), random-init weights",
        "config": {"workload": "%s seq %d, global_bsz %d, strategy %s" % (opts.model, config.max_position_embeddings,
                                                                       args.global_train_batch_size, os.path.basename(spath)),
                   "strategy": {k: strategy[k] for k in ("pp_deg", "chunks", "default_dp_type", "global_bsz") if k in strategy},
                   "tp": sorted(set(strategy["tp_sizes_enc"].split(","))), "checkpointed_layers": strategy.get("checkpoint", "").count("1"),
                   "layers": config.num_hidden_layers, "l2": "inputs (16 GB of bf16 weights + activations per step) far exceed the 126 MB L2",
                   "optimizer": ("AdamW fused into the gradient reduce-scatter kernel (fp32 shards)" if opts.optimizer == "fused"
                                 else "torch.optim.AdamW(fused=True) on fp32 flat shards")},
        "e2e": {"value": round(tokens_per_step * K / (ms_e2e * 1e-3), 1), "unit": "tokens/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4 * max(1, strategy["chunks"]), "ms_per_step": round(ms_e2e / K, 3)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
                     "frac": round(achieved / peak, 4) if peak else None, "traffic": traffic, "traffic_note": traffic_note,
                     "algorithmic_bytes_per_launch_avg": gemm_bytes / max(1, len(prof)), "kernel": "gemm_bf16_kernel (tcgen05/TMEM/TMA)",
                     "launches": len(prof), "kernel_ms_per_step": round(gemm_ms / K, 3), "share_of_step": round(gemm_ms / ms_res, 4),
                     "flops_per_launch_avg": gemm_flops / max(1, len(prof)), "peak_source": peak_src},
        "clocks": clocks, "loss": {"resident": loss_res, "e2e": loss_e2e},
        "memory_gib": {"torch_peak_allocated": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
                       "torch_peak_reserved": round(torch.cuda.max_memory_reserved() / 2 ** 30, 2),
                       "arena": round(be.comm.arena_bytes / 2 ** 30, 2)},
    }
    if opts.layers:
        line["invalid"] = "debug run with --layers %d: not the BASELINE workload" % opts.layers
    if rank == 0 and world == 1 and not opts.no_cpu_baseline:
        from hetu_galvatron_b200.core.runtime.backend import reset_backend
        reset_backend()
        line["cpu_baseline"] = cpu_reference_sample(opts)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(opts, budget_s=25.0):
    """The CPU restatement of the reference path (oracle backend, all host threads) on a bounded sample of the workload:
    Llama-3-8B shapes, embedding + lm_head + ONE and then TWO transformer layers, seq 1024, batch 1.  The two samples separate
    the per-layer cost from the embedding/head cost; ``value`` is the full-depth (32-layer) tokens/s they imply -- an
    extrapolation that favours the CPU (seq 1024 instead of 8192: 8x less attention work per token), labelled as such, never a
    like-for-like 8B measurement."""
    import torch
    from oracle.gloo_backend import OracleBackend
    from hetu_galvatron_b200.core.runtime.backend import reset_backend, set_backend
    from hetu_galvatron_b200.core.runtime.utils import get_optimizer_and_param_scheduler
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)

    def best_step(n_layers, budget):
        set_backend(OracleBackend())
        sample = argparse.Namespace(**vars(opts))
        sample.layers, sample.seq, sample.optimizer, sample.checkpoint_layers = n_layers, 1024, "torch", -1
        one = ",".join(["1"] * n_layers)
        zero = ",".join(["0"] * n_layers)
        strategy = {"pp_deg": 1, "tp_sizes_enc": one, "tp_consecutive_flags": one, "dp_types_enc": zero, "use_sp": zero,
                    "checkpoint": zero, "global_bsz": 1, "chunks": 1, "default_dp_type": "zero2", "vtp": 1}
        args, config, model = build_model(sample, strategy)
        opt, _ = get_optimizer_and_param_scheduler(model, args)
        batches = synthetic_batches(args, config, 3, 0, 1, pin=False)
        times, t_start = [], time.perf_counter()
        for i, (t, l) in enumerate(batches):
            t0 = time.perf_counter()
            model.forward_backward([t], i, None, loss_func=None, attention_mask=None, labels=l)
            opt.step(); opt.zero_grad()
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > budget and i >= 1:
                break
        reset_backend()
        return (min(times[1:]) if len(times) > 1 else times[0]), len(times), config.max_position_embeddings

    t1, n1, tok = best_step(1, budget_s * 0.4)
    t2, n2, _ = best_step(2, budget_s * 0.6)
    layer_s = max(t2 - t1, 0.05 * t1)
    other_s = max(t1 - layer_s, 0.0)
    full_layers = 32
    full_s = other_s + full_layers * layer_s
    return {"value": round(tok / full_s, 3), "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": "oracle CPU restatement (fp32 compute, bf16 storage), Llama-3-8B shapes, seq 1024, batch 1: embedding + lm_head + "
                      "1 layer (%d steps, best %.2f s) and + 2 layers (%d steps, best %.2f s) -> %.2f s per layer, %.2f s for the rest; "
                      "value = 1024 tokens / (rest + 32 layers) = extrapolated full-depth rate, NOT a like-for-like seq-8192 run"
                      % (n1, t1, n2, t2, layer_s, other_s),
            "sample_tokens_per_s_1layer": round(tok / t1, 2)}


def run_reference(opts):
    """--impl reference: the reference's path as restated on CPU (the reference has no CPU runtime: SURVEY 8c / BASELINE.md 3),
    all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    K = max(1, opts.steps)
    from hetu_galvatron_b200.core.runtime import world as _world
    with _world.simulated(0, 1):        # under torchrun the env says world N; the CPU sample is a single-process job
        base = cpu_reference_sample(opts, budget_s=20.0 * K)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "tokens/s", "n_gpus": opts.gpus, "steps": opts.steps,
            "warmup": opts.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 storage / fp32 compute",
            "data": "synthetic tokens (DataLoaderForLlama generator, seed 1234), random-init weights",
            "config": {"workload": "%s seq %d (bounded sample: %s)" % (opts.model, opts.seq, base["sample"])},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ms_per_step": None}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    o = parse()
    if o.impl == "reference":
        run_reference(o)
    else:
        try:
            run_ours(o)
        except Exception:
            # a device-side barrier timeout traps the kernel; its who/where record survives in mapped host memory
            try:
                from hetu_galvatron_b200.core.runtime.backend import get_backend
                sys.stderr.write("rank %s: device error info %s\n" % (os.environ.get("RANK", "0"), get_backend().comm.error_info()))
            except Exception:
                pass
            raise
