"""Tiny end-to-end training steps of the Llama family through the public API, checked against the oracle
(test infrastructure: used by ``__graft_entry__.smoke()``, tests/_host_worker.py and bench.py's leg parity checks; it lives in
tests/ so that the product package never imports ``oracle``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import types

import torch


def tiny_args(**over):
    from hetu_galvatron_b200.core.runtime.arguments import initialize_galvatron
    kw = dict(pp_deg=1, global_tp_deg=1, global_cp_deg=1, sdp=0, default_dp_type="zero2", chunks=1, global_train_batch_size=4,
              mixed_precision="bf16", pipeline_type="pipedream_flush", sequence_parallel=False, use_ulysses=False,
              vocab_tp=1, vocab_cp=1, global_checkpoint=0, make_vocab_size_divisible_by=128, init_method_std=0.05, seed=1234,
              local_rank=1, lr=1e-3, adam_weight_decay=0.01)
    kw.update(over)
    return initialize_galvatron(**kw)


TINY = dict(dim=128, ffn_dim=352, n_heads=4, n_kv_heads=2, n_layers=2, norm_eps=1e-5, vocab_size=512, n_positions=64)


def build(args, spec=None):
    from hetu_galvatron_b200.llama_hf import config_from_meta, llama_model_hp, set_model_config
    config = set_model_config(config_from_meta(dict(spec or TINY)), args)
    return config, llama_model_hp(config, args)


def oracle_weights(model, config, tensor_of=None):
    """fp32 master weights (or, with ``tensor_of=lambda unit: unit.master_grad``, gradients) of an un-sharded model in the
    oracle's dict layout."""
    units = model.model.units
    get = (lambda u: u.local_master_slices()) if tensor_of is None else (lambda u: u.local_master_slices(tensor_of(u)))
    layers = []
    for u in units[1:-2]:
        s = get(u)
        layers.append({"ln1": s["layer.attention.LayerNorm.weight"], "qkv": s["layer.attention.attention.query_key_value.weight"],
                       "dense": s["layer.attention.attention.dense.weight"], "ln2": s["layer.mlp.LayerNorm.weight"],
                       "h_to_4h": s["layer.mlp.mlp.dense_h_to_4h.weight"], "4h_to_h": s["layer.mlp.mlp.dense_4h_to_h.weight"]})
    return {"embed": get(units[0])["embed_tokens.weight"], "layers": layers, "norm": get(units[-2])["norm.weight"],
            "lm_head": get(units[-1])["lm_head.weight"]}


def oracle_cfg(config, args):
    return dict(hidden=config.hidden_size, ffn=config.intermediate_size, n_heads=config.num_attention_heads,
                n_kv_heads=config.num_key_value_heads, head_dim=config.hidden_size // config.num_attention_heads,
                n_layers=config.num_hidden_layers, vocab=args.padded_vocab_size, eps=config.rms_norm_eps,
                rope_base=getattr(args, "rotary_base", 10000.0))


def run(steps=2, verbose=True, **over):
    """Train the tiny model for ``steps`` iterations on cuda:0 and compare loss + gradients of the first step with the oracle."""
    from oracle import llama_ref
    from hetu_galvatron_b200.core.runtime.backend import reset_backend
    from hetu_galvatron_b200.core.runtime.utils import get_optimizer_and_param_scheduler
    reset_backend()
    args = tiny_args(**over)
    torch.manual_seed(args.seed)
    config, model = build(args)
    opt, _ = get_optimizer_and_param_scheduler(model, args)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(7)
    seq, bsz = config.max_position_embeddings, args.global_train_batch_size
    losses = []
    for it in range(steps):
        x = torch.randint(0, config.vocab_size, (bsz, seq + 1), generator=g)
        tokens, labels = x[:, :-1].contiguous().to(dev), x[:, 1:].contiguous().to(dev)
        if it == 0:
            w = {k: (v.detach().cpu().clone() if torch.is_tensor(v) else [{kk: vv.detach().cpu().clone() for kk, vv in lw.items()} for lw in v])
                 for k, v in oracle_weights(model, config).items()}
        loss = model.forward_backward([tokens], it, None, loss_func=None, attention_mask=None, labels=labels)
        if it == 0:
            torch.cuda.synchronize()
            cfg = oracle_cfg(config, args)
            leaves = [w["embed"], w["norm"], w["lm_head"]] + [t for lw in w["layers"] for t in lw.values()]
            for t in leaves:
                t.requires_grad_(True)
            _, ref_loss = llama_ref.forward_loss(w, tokens.cpu(), labels.cpu(), cfg, dtype=torch.bfloat16)
            ref_loss.backward()
            got = oracle_weights(model, config, tensor_of=lambda u: u.master_grad)
            rel = lambda a, b: float((a.cpu().float() - b.float()).norm() / (b.float().norm() + 1e-12))  # noqa: E731
            errs = {"lm_head": rel(got["lm_head"], w["lm_head"].grad), "embed": rel(got["embed"], w["embed"].grad),
                    "qkv0": rel(got["layers"][0]["qkv"], w["layers"][0]["qkv"].grad),
                    "h_to_4h0": rel(got["layers"][0]["h_to_4h"], w["layers"][0]["h_to_4h"].grad),
                    "ln1_0": rel(got["layers"][0]["ln1"], w["layers"][0]["ln1"].grad),
                    "dense1": rel(got["layers"][-1]["dense"], w["layers"][-1]["dense"].grad)}
            if verbose:
                print("step0 loss ours %.6f oracle(bf16) %.6f  grad rel-L2 errs %s" % (loss, float(ref_loss), errs))
            assert abs(loss - float(ref_loss)) <= 5e-3 * abs(float(ref_loss)) + 1e-3, (loss, float(ref_loss))
            assert max(errs.values()) < 3e-2, errs
        opt.step()
        opt.zero_grad()
        losses.append(loss)
    if verbose:
        print("losses:", losses)
    assert losses[-1] < losses[0] + 0.5
    reset_backend()
    return losses
