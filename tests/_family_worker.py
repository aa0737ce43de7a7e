"""Worker for tests/test_families.py: one rank of a job running the GPT or BERT family of the product on the CPU oracle backend
(gloo) or on GPUs (HOST_TEST_BACKEND=cuda), checked against the single-process oracle (oracle/gpt_bert_ref.py, pinned to HF) on the
GLOBAL batch: loss within 5e-3 rel (tests/models/test_model_correctness.py:111-117), every parameter's gradient within 3e-2 rel-L2."""
import json
import os
import re
import sys
import traceback

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TINY = {"gpt": dict(n_layer=2, n_embd=128, n_head=4, vocab_size=512, n_positions=64),
        "bert": dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=4, vocab_size=512, max_position_embeddings=64,
                     layer_norm_eps=1e-5)}

# parameter name inside a unit (after stripping relocation / "layer." prefixes) -> (oracle key, tensor-parallel concat dim or None)
LAYER = {"attention.LayerNorm.weight": ("ln1", None), "attention.LayerNorm.bias": ("ln1_b", None),
         "attention.attention.query_key_value.weight": ("qkv", 0), "attention.attention.query_key_value.bias": ("qkv_b", 0),
         "attention.attention.dense.weight": ("dense", 1), "attention.attention.dense.bias": ("dense_b", None),
         "mlp.LayerNorm.weight": ("ln2", None), "mlp.LayerNorm.bias": ("ln2_b", None),
         "mlp.mlp.dense_h_to_4h.weight": ("h_to_4h", 0), "mlp.mlp.dense_h_to_4h.bias": ("h_to_4h_b", 0),
         "mlp.mlp.dense_4h_to_h.weight": ("4h_to_h", 1), "mlp.mlp.dense_4h_to_h.bias": ("4h_to_h_b", None)}
OTHER = {"gpt": {"wte.wte.weight": ("wte", 0), "wpe.wpe.weight": ("wpe", 0), "ln_f.weight": ("norm", None), "ln_f.bias": ("norm_b", None),
                 "lm_head.weight": ("lm_head", 0)},
         "bert": {"word_embeddings.word_embeddings.weight": ("word", 0), "position_embeddings.position_embeddings.weight": ("pos", 0),
                  "token_type_embeddings.token_type_embeddings.weight": ("type", None), "LayerNorm.weight": ("emb_ln", None),
                  "LayerNorm.bias": ("emb_ln_b", None), "transform.weight": ("transform", None), "transform.bias": ("transform_b", None),
                  "transform.LayerNorm.weight": ("transform_ln", None), "transform.LayerNorm.bias": ("transform_ln_b", None),
                  "lm_head.weight": ("decoder", 0), "lm_head.bias": ("decoder_b", 0)}}


def assemble(model, family, world, rank, tensor_of):
    """every rank's per-unit named tensors -> the un-parallelised oracle weight dict"""
    per_unit = []
    for u in model.model.units:
        per_unit.append({"name": u.name, "tp": list(u.tp_group.ranks) if u.tp_group is not None else [rank],
                         "slices": {re.sub(r"^(module\.)*(layer\.)?", "", k): v.detach().float().cpu().clone()
                                    for k, v in u.named_slices(tensor_of(u)).items()}})
    gathered = [None] * world
    dist.all_gather_object(gathered, per_unit)
    by_name = {}
    for r, units in enumerate(gathered):
        for rec in units:
            by_name.setdefault(rec["name"], {})[r] = rec
    out, layers = {}, {}
    for name, recs in by_name.items():
        first = recs[sorted(recs)[0]]
        table = LAYER if re.match(r"(gpt_dec|bert_enc)_\d+", name) else OTHER[family]
        target = layers.setdefault(name, {}) if table is LAYER else out
        for pname in first["slices"]:
            key, dim = table[pname]
            parts = [recs[r]["slices"][pname] for r in first["tp"]]
            target[key] = parts[0] if dim is None or len(parts) == 1 else torch.cat(parts, dim=dim)
    out["layers"] = [layers[k] for k in sorted(layers, key=lambda s: int(s.rsplit("_", 1)[1]))]
    return out


def gather_grads(model, world):
    mine = {u.name: (list(u.group.ranks), u.master_grad.detach().float().cpu().clone()) for u in model.model.units}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    return {u.name: (u.master_grad.detach().float().cpu().clone() if u.dp_type == "ddp" or u.group.size == 1
                     else torch.cat([allr[r][u.name][1] for r in u.group.ranks])) for u in model.model.units}


def leaves_of(w):
    return [t for t in w.values() if torch.is_tensor(t)] + [t for lw in w["layers"] for t in lw.values()]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    over = json.loads(os.environ["HOST_TEST_CONFIG"])
    family = over.pop("_family")
    spec = dict(TINY[family], **over.pop("_spec", {}))
    tol = over.pop("_tol", 3e-2)
    golden_ckpt = over.pop("_golden_ckpt", None)
    use_cuda = os.environ.get("HOST_TEST_BACKEND", "oracle") == "cuda"
    from oracle import gpt_bert_ref as ref
    import smoke_model as sm
    from hetu_galvatron_b200.core.runtime.backend import get_backend, reset_backend, set_backend
    from hetu_galvatron_b200.core.runtime.utils import get_optimizer_and_param_scheduler
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", torch.cuda.current_device()))
        os.environ.setdefault("HGB_ARENA_BYTES", str(256 << 20))
        be = get_backend()
        be.bg.set_tunable("timeout_ms", 30000)
        dev = be.device
    else:
        from oracle.gloo_backend import OracleBackend
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1 if world >= 4 else 2)
        be = set_backend(OracleBackend())
        dev = torch.device("cpu")
    args = sm.tiny_args(**over)
    if family == "gpt":
        from hetu_galvatron_b200.gpt_hf import config_from_meta, gpt_model_hp as build, set_model_config
    else:
        from hetu_galvatron_b200.bert_hf import bert_model_hp as build, config_from_meta, set_model_config
    config = set_model_config(config_from_meta(spec), args)
    model = build(config, args)
    opt, _ = get_optimizer_and_param_scheduler(model, args)
    w = assemble(model, family, world, rank, lambda u: u.read_full_params())
    report0 = {}
    tied = family == "gpt" and not getattr(args, "untie_embeddings_and_output_weights", True)
    if tied:        # one matrix, two copies: they must be identical, and the oracle then uses ONE leaf for both (HF's tie_word_embeddings)
        assert torch.equal(w["wte"], w["lm_head"]), "tied embeddings: wte and lm_head differ after construction"
        w["lm_head"] = w["wte"]
        report0["tied"] = True
    if golden_ckpt:   # the weights the model loaded must be HF's, bit for bit, whatever the tensor-parallel degree
        cfg0 = dict(n_heads=config.num_attention_heads, head_dim=config.hidden_size // config.num_attention_heads)
        hf = ref.to_hf_state_dict(w, cfg0, family)
        want = {}
        for fname in sorted(os.listdir(golden_ckpt)):
            if not fname.endswith(".pt"):
                continue
            blob = torch.load(os.path.join(golden_ckpt, fname), map_location="cpu", weights_only=True)
            stem = fname[:-3]
            for k, v in blob.items():
                if stem == "transformer_embedding":
                    want["lm_head.weight" if k == "weight" else "transformer." + k] = v
                elif stem == "transformer_ln_f":
                    want["transformer.ln_f." + k] = v
                elif stem.startswith("transformer_h_"):
                    want["transformer.h.%s.%s" % (stem.rsplit("_", 1)[1], k)] = v
                elif stem == "bert_embeddings":
                    want["bert.embeddings." + k] = v
                elif stem.startswith("bert_encoder_layer_"):
                    want["bert.encoder.layer.%s.%s" % (stem.rsplit("_", 1)[1], k)] = v
                elif stem == "cls_predictions":
                    want["cls.predictions." + k] = v
        assert set(want) == set(hf), sorted(set(want) ^ set(hf))
        bad = [k for k in want if not torch.equal(want[k].float(), hf[k].float())]
        assert not bad, "loaded weights differ from the HF checkpoint: %s" % bad
        report0["ckpt_tensors_bit_exact"] = len(want)
    cfg = dict(hidden=config.hidden_size, ffn=config.intermediate_size, n_heads=config.num_attention_heads,
               head_dim=config.hidden_size // config.num_attention_heads, n_layers=config.num_hidden_layers, vocab=args.padded_vocab_size,
               seq=config.max_position_embeddings, eps=args.norm_epsilon, gelu_tanh=True)
    gbs, seq = args.global_train_batch_size, config.max_position_embeddings
    dp_group = model.vtp_data_group
    dp_idx, dp = dp_group.rank_in_group(rank), dp_group.size
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, config.vocab_size, (gbs, seq + 1), generator=g)
    tokens, labels = x[:, :-1].contiguous(), x[:, 1:].contiguous()
    kwargs = {}
    mask = tt = None
    if family == "bert":
        lengths = torch.randint(seq // 2, seq + 1, (gbs,), generator=g)
        mask = torch.arange(seq)[None, :] < lengths[:, None]
        tt = (torch.arange(seq)[None, :] >= (lengths[:, None] // 2)).long() * mask.long()
        labels = torch.where(torch.rand(gbs, seq, generator=g) < 0.15, tokens, torch.full_like(tokens, -100))   # MLM: -100 elsewhere
        labels = torch.where(mask, labels, torch.full_like(labels, -100))
    lo, hi = dp_idx * gbs // dp, (dp_idx + 1) * gbs // dp
    if family == "bert":
        kwargs = dict(attention_mask=mask[lo:hi].to(dev), token_type_ids=tt[lo:hi].to(dev))
    else:
        kwargs = dict(attention_mask=None)
    loss = model.forward_backward([tokens[lo:hi].to(dev)], 0, None, loss_func=None, labels=labels[lo:hi].to(dev), **kwargs)
    if use_cuda:
        torch.cuda.synchronize()
        assert be.comm.error_flag() == 0
    for t in leaves_of(w):
        t.requires_grad_(True)
    if family == "gpt":
        _, ref_loss = ref.gpt_forward_loss(w, tokens, labels, cfg, dtype=torch.bfloat16)
    else:
        _, ref_loss = ref.bert_forward_loss(w, tokens, labels, cfg, dtype=torch.bfloat16, attention_mask=mask, token_type_ids=tt)
    ref_loss.backward()
    grads = gather_grads(model, world)
    got = assemble(model, family, world, rank, lambda u: grads[u.name])
    rel = lambda a, b: float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))  # noqa: E731
    # reference semantics (see tests/_host_worker.py): a unit's gradient is averaged over ITS sharded-data-parallel group
    dp_cls = model.hp_configs_whole["dp_sizes_whole"][-1] * model.hp_configs_whole["cp_sizes_whole"][-1]
    sizes = [None] * world
    dist.all_gather_object(sizes, {u.name: u.group.size for u in model.model.units})
    scale = {k: dp_cls / v for d in sizes for k, v in d.items()}
    unit_of = {}
    for u in model.model.units:
        kind = re.sub(r"_\d+$", "", u.name)
        unit_of[kind if kind not in ("gpt_dec", "bert_enc") else u.name] = u.name
    names = [None] * world
    dist.all_gather_object(names, unit_of)
    unit_of = {k: v for d in names for k, v in d.items()}
    errs = {}
    for key, t in got.items():
        if key == "layers":
            continue
        kind = {"wte": "embed", "wpe": "embed", "word": "embed", "pos": "embed", "type": "embed", "emb_ln": "embed", "emb_ln_b": "embed",
                "norm": "norm", "norm_b": "norm", "lm_head": "cls"}.get(key, "mlm_head")
        errs[key] = rel(t, w[key].grad * scale[unit_of[kind]])
    layer_units = sorted([n for n in scale if re.match(r"(gpt_dec|bert_enc)_\d+", n)], key=lambda s: int(s.rsplit("_", 1)[1]))
    for i, (gl, wl) in enumerate(zip(got["layers"], w["layers"])):
        for k in gl:
            errs["%s%d" % (k, i)] = rel(gl[k], wl[k].grad * scale[layer_units[i]])
    lt = torch.tensor([loss if loss is not None else 0.0, 1.0 if loss is not None else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(lt)
    mean_loss = float(lt[0] / lt[1])
    report = dict(report0, loss=mean_loss, ref_loss=float(ref_loss), max_grad_err=max(errs.values()), worst=max(errs, key=errs.get))
    if use_cuda:
        report["launches"] = be.launch_count()
        report["fused_calls"] = dict(getattr(be, "n_fused", {}))
    assert abs(mean_loss - float(ref_loss)) <= 5e-3 * abs(float(ref_loss)), report
    assert report["max_grad_err"] < tol, (report, {k: round(v, 4) for k, v in errs.items() if v > tol / 3})
    opt.step()
    opt.zero_grad()
    loss2 = model.forward_backward([tokens[lo:hi].to(dev)], 1, None, loss_func=None, labels=labels[lo:hi].to(dev), **kwargs)
    lt = torch.tensor([loss2 if loss2 is not None else 0.0, 1.0 if loss2 is not None else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(lt)
    report["loss_step1"] = float(lt[0] / lt[1])
    # the same AdamW step on the oracle's weights (its gradients scaled the way the runtime averages them, see above), then the oracle
    # forward again: the optimizer + the re-gather of the updated parameters must land on the same loss.  (A plain "the loss went
    # down" is not a property of one sign-like Adam step on ~60 MLM labels: it fails for one seed in six on the oracle itself.)
    leaf_scale = {}
    for key, t in w.items():
        if key != "layers":
            kind = {"wte": "embed", "wpe": "embed", "word": "embed", "pos": "embed", "type": "embed", "emb_ln": "embed", "emb_ln_b": "embed",
                    "norm": "norm", "norm_b": "norm", "lm_head": "cls"}.get(key, "mlm_head")
            leaf_scale[id(t)] = scale[unit_of[kind]]
    for i, wl in enumerate(w["layers"]):
        for t in wl.values():
            leaf_scale[id(t)] = scale[layer_units[i]]
    leaves = list({id(t): t for t in leaves_of(w) if t.grad is not None}.values())      # (tied embeddings: one leaf, listed twice)
    with torch.no_grad():
        for t in leaves:
            t.grad.mul_(leaf_scale[id(t)])
    if tied:        # ... and after the step the two copies of the product must still be one matrix
        w_after = assemble(model, family, world, rank, lambda u: u.read_full_params())
        assert torch.equal(w_after["wte"], w_after["lm_head"]), "tied embeddings drifted apart after one optimizer step"
    ref_opt = torch.optim.AdamW(leaves, lr=args.lr, weight_decay=args.adam_weight_decay,
                                betas=(getattr(args, "adam_beta1", 0.9), getattr(args, "adam_beta2", 0.999)), eps=getattr(args, "adam_eps", 1e-8))
    ref_opt.step()
    with torch.no_grad():
        if family == "gpt":
            _, ref_loss1 = ref.gpt_forward_loss(w, tokens, labels, cfg, dtype=torch.bfloat16)
        else:
            _, ref_loss1 = ref.bert_forward_loss(w, tokens, labels, cfg, dtype=torch.bfloat16, attention_mask=mask, token_type_ids=tt)
    report["ref_loss_step1"] = float(ref_loss1)
    assert abs(report["loss_step1"] - report["ref_loss_step1"]) <= 5e-3 * abs(report["ref_loss_step1"]), report
    if rank == 0:
        print("HOST_TEST_REPORT " + json.dumps(report), flush=True)
    dist.barrier()
    if use_cuda:
        reset_backend()
    dist.destroy_process_group()
    return report


if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
