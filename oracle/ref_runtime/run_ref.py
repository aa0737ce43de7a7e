"""Run the UNMODIFIED reference runtime (PKU-DAIR/Hetu-Galvatron, installed by ``pip install --target baseline/_ref``; git-ignored,
travels to the GPU box with the snapshot) on GPUs and record what it computes: per-step losses of 3 training steps of the tiny
Llama whose weights are tests/golden/ckpt_llama_tiny (an HF checkpoint converted by the reference's own h2g tool), under a given
parallel strategy.  TEST INFRASTRUCTURE: the output (gpurun_out/ref_runtime_<case>.json) is committed as
tests/golden/ref_runtime/<case>.json and pins this repo's fp parity to the reference RUNTIME (tests/test_ref_runtime_parity.py),
not only to HF.  The harness follows the reference's own tests (tests/core/test_tp.py:18-121: RuntimeArgs, set_args, the family's get_llama_config /
llama_model_hp that tests/utils/model_utils.ModelFactory resolves to).

    torchrun --nproc-per-node N --master-addr 127.0.0.1 oracle/ref_runtime/run_ref.py --case tp2 [--out gpurun_out]

What is NOT the reference here: apex / amp_C / dropout_layer_norm are absent from the image -- oracle/ref_shim supplies the four
import-time symbols and a plain-torch restatement of the norm extension's two entry points; the optimizer is torch.optim.Adam
with weight_decay 0 (the reference's tests use torch Adam too, test_tp.py:78) so that Adam == AdamW.
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "baseline", "_ref")
# the reference first (its own `tests` package must win over this repo's tests/ directory), then the import shims
sys.path[:0] = [REF, os.path.join(REF, "galvatron", "site_package"), os.path.join(ROOT, "oracle", "ref_shim")]
if "" in sys.path:
    sys.path.remove("")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "ckpt_llama_tiny")

# case -> (world, overrides of the reference's RuntimeArgs); the same names / meanings as tests/test_host_runtime.py
CASES = {
    "world1": (1, dict()),
    "dp2_zero2": (2, dict(default_dp_type="zero2")),
    "dp2_zero3_ckpt": (2, dict(sdp=1, embed_sdp=1, global_checkpoint=1)),
    "tp2": (2, dict(global_tp_deg=2, vocab_tp=2)),
    "tp2_megatron_sp": (2, dict(global_tp_deg=2, vocab_tp=2, sequence_parallel=True)),
    "ulysses2": (2, dict(global_tp_deg=2, vocab_tp=2, sequence_parallel=True, use_ulysses=True)),
    "pp2_1f1b": (2, dict(pp_deg=2, chunks=2, pipeline_type="pipedream_flush")),
    "tp2_dp2_sp_zero2": (4, dict(global_tp_deg=2, vocab_tp=2, sequence_parallel=True, default_dp_type="zero2", chunks=2)),
    "pp2_tp2_1f1b": (4, dict(pp_deg=2, global_tp_deg=2, vocab_tp=2, chunks=2, pipeline_type="pipedream_flush")),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True, choices=sorted(CASES))
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--steps", type=int, default=3)
    # --bench: not a fixture but a timing of the reference runtime itself on real-size Llama-3-8B layers (random init on the device,
    # synthetic tokens): the GPU-side yardstick next to bench.py --layers L (same strategy: ZeRO-2, chunks, no checkpointing)
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--seq", type=int, default=8192)
    ap.add_argument("--gbs", type=int, default=8)
    ap.add_argument("--chunks", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=2)
    opts = ap.parse_args()
    world_want, over = CASES[opts.case]
    dist.init_process_group("nccl", init_method="env://")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert opts.bench or world == world_want, "case %s runs on %d ranks" % (opts.case, world_want)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)

    from galvatron.utils.training_utils import set_seed
    from megatron.core.parallel_state import initialize_model_parallel
    from megatron.core.tensor_parallel import random
    from megatron.training.global_vars import set_args
    from galvatron.models.llama_hf.LlamaModel_hybrid_parallel import get_llama_config, llama_model_hp
    from tests.utils.runtime_args import RuntimeArgs

    spec = json.load(open(os.path.join(GOLDEN, "expected.json")))["spec"]
    set_seed(1234)
    initialize_model_parallel(tensor_model_parallel_size=1, pipeline_model_parallel_size=1)
    random.model_parallel_cuda_manual_seed(1234)
    if opts.bench:
        spec = {"hidden_size": 4096, "intermediate_size": 14336, "num_attention_heads": 32, "num_key_value_heads": 8, "num_hidden_layers": opts.layers,
                "rms_norm_eps": 1e-5, "vocab_size": 128256, "max_position_embeddings": opts.seq}
    args = RuntimeArgs(model_type="llama", rank=rank, checkpoint_dir=None if opts.bench else {"converted": GOLDEN}, backend="hf")
    # HEAD's test RuntimeArgs predates context parallelism: every option of the runtime's own parser that it lacks gets the
    # parser's default (galvatron/core/runtime/arguments.py: galvatron_training_args)
    import argparse as _ap
    from galvatron.core.runtime.arguments import galvatron_training_args
    _parser = _ap.ArgumentParser()
    galvatron_training_args(_parser, use_megatron=False)
    for k, v in vars(_parser.parse_args([])).items():
        if not hasattr(args, k):
            setattr(args, k, v)
    # the tiny Llama of the golden checkpoint, as a dict spec (always with ffn_dim: config_utils.py:33-35)
    args.model_size = {"dim": spec["hidden_size"], "ffn_dim": spec["intermediate_size"], "n_heads": spec["num_attention_heads"],
                       "n_kv_heads": spec["num_key_value_heads"], "n_layers": spec["num_hidden_layers"], "norm_eps": spec["rms_norm_eps"],
                       "vocab_size": spec["vocab_size"], "n_positions": spec["max_position_embeddings"], "multiple_of": 32}
    args.global_train_batch_size, args.chunks = (opts.gbs, opts.chunks) if opts.bench else (4, 1)
    args.mixed_precision, args.use_flash_attn = "bf16", True
    args.default_dp_type, args.pipeline_type = "zero2", "pipedream_flush"
    # reduce_in_fp32 / entropy_in_fp32 stay True, as the reference's own tests run (tests/utils/runtime_args.py:60-62): the loss is
    # then resolved to fp32 instead of one bf16 ulp (0.03 at 6.25)
    args.make_vocab_size_divisible_by = 128
    args.untie_embeddings_and_output_weights = True
    args.lr, args.adam_weight_decay = 1e-3, 0.0
    args.init_method_std = 0.02          # random init on the device (--bench: no checkpoint; tensor_parallel/reset.py:13)
    for k, v in over.items():
        setattr(args, k, v)
    args.vocab_sp = 1 if getattr(args, "use_ulysses", False) else 0
    args.micro_batch_size = args.global_train_batch_size
    args.tp_deg = args.global_tp_deg
    set_args(args)

    config = get_llama_config(args, True)            # what tests/utils/model_utils.ModelFactory.create_config / create_model call
    model = llama_model_hp(config, args)
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=0.0)
    dp_group = model.dp_groups_whole[0]
    dp_ranks = list(dp_group.ranks)
    dp_idx, dp = dp_ranks.index(rank), len(dp_ranks)
    gbs, seq = args.global_train_batch_size, spec["max_position_embeddings"]
    g = torch.Generator().manual_seed(11)
    if opts.bench:
        batches = [torch.randint(0, spec["vocab_size"], (gbs, seq + 1), generator=g) for _ in range(opts.warmup + opts.steps)]
        lo, hi = dp_idx * gbs // dp, (dp_idx + 1) * gbs // dp

        def one(x, it):
            tokens, labels = x[lo:hi, :-1].contiguous().to(device), x[lo:hi, 1:].contiguous().to(device)
            loss = model.forward_backward([tokens], it, None, loss_func=None, attention_mask=None, labels=labels)
            optimizer.step()
            optimizer.zero_grad()
            return loss
        for i in range(opts.warmup):
            one(batches[i], i)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(opts.warmup, opts.warmup + opts.steps):
            last = one(batches[i], i)
        e1.record()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rec = {"impl": "reference runtime on GPU (baseline/_ref: FSDP + NCCL + flash-attn 2 + torch ops; torch.optim.Adam on fp32 flat params)",
               "model": "Llama-3-8B shapes, %d layers" % opts.layers, "seq": seq, "global_bsz": gbs, "chunks": args.chunks, "world": world,
               "default_dp_type": args.default_dp_type, "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": round(float(ms[0]) / opts.steps, 3),
               "tokens_per_s": round(gbs * seq * opts.steps / (float(ms[0]) * 1e-3), 1), "last_loss": last,
               "peak_allocated_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
        if rank == 0:
            os.makedirs(opts.out, exist_ok=True)
            with open(os.path.join(opts.out, "ref_runtime_bench_%dlayers_n%d.json" % (opts.layers, world)), "w") as f:
                json.dump(rec, f, indent=1)
            print("REF_BENCH " + json.dumps(rec), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return
    losses, grad_norms = [], []
    for it in range(opts.steps):
        x = torch.randint(0, spec["vocab_size"], (gbs, seq + 1), generator=g)
        tokens, labels = x[:, :-1].contiguous(), x[:, 1:].contiguous()
        lo, hi = dp_idx * gbs // dp, (dp_idx + 1) * gbs // dp
        loss = model.forward_backward([tokens[lo:hi].to(device)], it, None, loss_func=None, attention_mask=None, labels=labels[lo:hi].to(device))
        # every rank's local gradient tensors as the optimizer sees them (FSDP: shards under ZeRO-2/3, full copies under DDP; the
        # tensor-parallel-replicated norm weights once per tp rank), squared and summed over the job
        sq = torch.zeros((), dtype=torch.float64, device=device)
        for p in model.parameters():
            if p.grad is not None:
                sq += p.grad.detach().double().pow(2).sum()
        dist.all_reduce(sq)
        grad_norms.append(float(sq.sqrt()))
        optimizer.step()
        optimizer.zero_grad()
        lt = torch.tensor([loss if loss is not None else 0.0, 1.0 if loss is not None else 0.0], dtype=torch.float64, device=device)
        dist.all_reduce(lt)
        losses.append(float(lt[0] / lt[1]))
    rec = {"case": opts.case, "world": world, "overrides": {k: (v if not isinstance(v, bool) else int(v)) for k, v in over.items()},
           "losses": losses, "grad_norms_all_ranks": grad_norms, "steps": opts.steps, "reduce_in_fp32": int(bool(args.reduce_in_fp32)),
           "entropy_in_fp32": int(bool(args.entropy_in_fp32)), "optimizer": "torch.optim.Adam lr 1e-3 wd 0", "global_batch": gbs, "token_seed": 11,
           "weights": "tests/golden/ckpt_llama_tiny", "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
           "producer": "oracle/ref_runtime/run_ref.py on the unmodified reference runtime (baseline/_ref)"}
    if rank == 0:
        os.makedirs(opts.out, exist_ok=True)
        with open(os.path.join(opts.out, "ref_runtime_%s.json" % opts.case), "w") as f:
            json.dump(rec, f, indent=1)
        print("REF_RUNTIME " + json.dumps(rec), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
