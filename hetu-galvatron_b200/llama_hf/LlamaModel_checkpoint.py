"""Layer-wise checkpoint I/O for the Llama family in the reference's two on-disk formats
(``galvatron/models/llama_hf/LlamaModel_checkpoint.py``), so checkpoints move between the two runtimes unchanged:

* **HF-layered** (what ``galvatron/tools/checkpoint_convert_h2g.py:47-87`` writes from a HuggingFace checkpoint):
  ``model_embed_tokens.pt`` {``embed_tokens.weight``}, ``model_layers_<i>.pt`` with the HF keys of one decoder layer,
  ``model_norm.pt`` {``weight``}, ``lm_head.pt`` {``weight``}.  ``load_hf_checkpoint`` (:48-143) slices every tensor for the
  caller's tensor-parallel rank: QKV rows in the per-group interleaved Megatron order, gate/up stacked per rank, row-parallel
  weights by input columns, vocab rows by range.
* **distributed** (what ``save_llama_module`` :156-216 writes and ``load_distributed_checkpoint`` :27-45 reads):
  ``<dir>/hybrid_parallel_configs.json``, ``<dir>/iter_<n>/opt_param_scheduler.json``,
  ``<dir>/iter_<n>/<model_embed_tokens|model_layers_<i>|model_norm|lm_head>/<tp_rank>.pt`` holding the block's full-precision
  state dict (keys relative to the wrapped block: ``attention.attention.query_key_value.weight`` ...), written by the first
  rank of every layer's sharded-data-parallel group, and ``<dir>/iter_<n>/optimizer/<rank>.pt``.

The loader is the ``load_module_func`` callback of ``construct_hybrid_parallel_model_api`` with the reference's signature
``(load, tp_groups, name, submodule, module, distributed_checkpoint)``; it runs once per parameter-owning submodule while a
sharded unit materialises its layer (the reference's ``param_init_fn``, ``parallel.py:79-89``).

Divergence, on purpose: the reference pads a vocabulary smaller than ``padded_vocab_size`` at the FRONT of the embedding /
head matrices (``F.pad(w, (0, 0, padding_size, 0))`` :57,78), which shifts every token's row; here the padding rows go to
the end (Megatron's convention).  Identical whenever no padding is needed (Llama-3: 128256 = 1002 x 128).
"""
import json
import os

import torch

from ..core.runtime.arguments import get_args
from ..core.runtime.backend import get_backend

embedding_name = "model_embed_tokens.pt"
layer_name = "model_layers_%d.pt"
ln_f_name = "model_norm.pt"
cls_name = "lm_head.pt"


def _tp(tp_groups, rank):
    """(index of this rank in the tensor-parallel group, its size); ``tp_groups`` is a CommGroup (or None)."""
    if tp_groups is None or tp_groups.size == 1:
        return 0, 1
    return tp_groups.rank_in_group(rank), tp_groups.size


def _range(total, index, size):
    per = total // size
    return index * per, (index + 1) * per


def _read(path):
    return torch.load(path, mmap=True, map_location="cpu", weights_only=True)


def _put(param, tensor):
    if tuple(param.shape) != tuple(tensor.shape):
        raise ValueError("checkpoint tensor of shape %s does not fit parameter of shape %s" % (tuple(tensor.shape), tuple(param.shape)))
    param.data.copy_(tensor.to(device=param.device, dtype=param.dtype))


def _pad_vocab(weight, padded):
    pad = padded - weight.shape[0]
    if pad < 0:
        raise ValueError("checkpoint vocabulary (%d) exceeds padded_vocab_size (%d)" % (weight.shape[0], padded))
    return weight if pad == 0 else torch.cat([weight, weight.new_zeros(pad, weight.shape[1])], 0)


def _dir_of(name, module):
    if name.endswith("embed_tokens"):
        return embedding_name[:-3]
    if name.endswith("lm_head"):
        return cls_name[:-3]
    if name.endswith("norm") and not hasattr(module, "idx"):
        return ln_f_name[:-3]
    return (layer_name % module.idx)[:-3]


def load_distributed_checkpoint(load, tp_groups, name, submodule, module):
    """``LlamaModel_checkpoint.py:27-45``: one file per (block, tp rank) under ``iter_<load_iteration>``."""
    args = get_args()
    index, _ = _tp(tp_groups, get_backend().rank)
    path = os.path.join(load, "iter_%d" % int(getattr(args, "load_iteration", 0)), _dir_of(name, module), "%d.pt" % index)
    _put(submodule.weight, _read(path)["%s.weight" % name].float())


def load_hf_checkpoint(load, tp_groups, name, submodule, module):
    """``LlamaModel_checkpoint.py:48-143``: slice the HF tensors of one block for this tensor-parallel rank."""
    args = get_args()
    index, size = _tp(tp_groups, get_backend().rank)
    if name.endswith("embed_tokens"):
        w = _pad_vocab(_read(os.path.join(load, embedding_name))["embed_tokens.weight"].float(), args.padded_vocab_size)
        lo, hi = _range(args.padded_vocab_size, index, size)
        return _put(submodule.weight, w[lo:hi])
    if name.endswith("lm_head"):
        w = _pad_vocab(_read(os.path.join(load, cls_name))["weight"].float(), args.padded_vocab_size)
        lo, hi = _range(args.padded_vocab_size, index, size)
        return _put(submodule.weight, w[lo:hi])
    if name.endswith("norm") and not hasattr(module, "idx"):
        return _put(submodule.weight, _read(os.path.join(load, ln_f_name))["weight"].float())
    ckpt = _read(os.path.join(load, layer_name % module.idx))
    if name.startswith("attention"):
        if name.endswith("LayerNorm"):
            return _put(submodule.weight, ckpt["input_layernorm.weight"].float())
        if name.endswith("query_key_value"):
            # HF keeps q [nh*dim, h], k and v [ng*dim, h]; the fused weight is per query group [q heads of the group | k | v]
            nh = args.num_attention_heads
            ng = args.num_query_groups if getattr(args, "group_query_attention", False) and args.num_query_groups else nh
            dim = getattr(args, "kv_channels", None) or args.hidden_size // nh
            fused = torch.cat([ckpt["self_attn.q_proj.weight"].float().reshape(ng, dim * nh // ng, -1),
                               ckpt["self_attn.k_proj.weight"].float().reshape(ng, dim, -1),
                               ckpt["self_attn.v_proj.weight"].float().reshape(ng, dim, -1)], 1).reshape(-1, args.hidden_size)
            lo, hi = _range(fused.shape[0], index, size)
            return _put(submodule.weight, fused[lo:hi])
        if name.endswith("dense"):
            w = ckpt["self_attn.o_proj.weight"].float()
            lo, hi = _range(w.shape[1], index, size)
            return _put(submodule.weight, w[:, lo:hi])
    elif name.startswith("mlp"):
        if name.endswith("LayerNorm"):
            return _put(submodule.weight, ckpt["post_attention_layernorm.weight"].float())
        if name.endswith("dense_h_to_4h"):
            lo, hi = _range(ckpt["mlp.gate_proj.weight"].shape[0], index, size)
            return _put(submodule.weight, torch.cat([ckpt["mlp.gate_proj.weight"][lo:hi].float(), ckpt["mlp.up_proj.weight"][lo:hi].float()], 0))
        if name.endswith("dense_4h_to_h"):
            w = ckpt["mlp.down_proj.weight"].float()
            lo, hi = _range(w.shape[1], index, size)
            return _put(submodule.weight, w[:, lo:hi])
    raise KeyError("no checkpoint rule for submodule %r of %s" % (name, type(module).__name__))


@torch.no_grad()
def load_llama_module(load, tp_groups, name, submodule, module, distributed_checkpoint):
    """``LlamaModel_checkpoint.py:146-151``."""
    if distributed_checkpoint:
        load_distributed_checkpoint(load, tp_groups, name, submodule, module)
    else:
        load_hf_checkpoint(load, tp_groups, name, submodule, module)


@torch.no_grad()
def save_llama_module(save_path, model, optimizer, opt_param_scheduler, iter_num, args):
    """``LlamaModel_checkpoint.py:154-216``: per layer, the first rank of its sharded-data-parallel group writes the block's
    full-precision state for its tensor-parallel rank.  The gather of the fp32 shards is C1 with an fp32 destination."""
    be = get_backend()
    rank = be.rank
    if rank == 0:
        os.makedirs(os.path.join(save_path, "iter_%d" % iter_num), exist_ok=True)
        with open(os.path.join(save_path, "hybrid_parallel_configs.json"), "w") as f:
            json.dump(model.hybrid_parallel_configs, f)
        with open(os.path.join(save_path, "iter_%d" % iter_num, "opt_param_scheduler.json"), "w") as f:
            json.dump(opt_param_scheduler.state_dict() if opt_param_scheduler is not None else {}, f)
    if args.default_dp_type == "ddp":
        raise ValueError("Save / Load distributed checkpoint is not supported for DDP")     # the reference's restriction (:173)
    root = os.path.join(save_path, "iter_%d" % iter_num)
    for block in model.model.model_cur_stage:
        unit = block.unit
        full = be.gather_master(unit)                      # collective over the unit's group: every member calls it
        if unit.group.size > 1 and unit.group.rank_in_group(rank) != 0:
            continue
        inner, prefix = _wrapped_block(block.module)
        state = {}
        for pname, tensor in unit.named_slices(full).items():
            key = pname[len(prefix):] if prefix and pname.startswith(prefix) else pname
            state[key] = tensor.detach().to("cpu", copy=True)
        first = next(iter(state))
        sub = first.rsplit(".", 1)[0]
        target = os.path.join(root, _dir_of(sub, inner))
        os.makedirs(target, exist_ok=True)
        _atomic_save(state, os.path.join(target, "%d.pt" % _tp(unit.tp_group, rank)[0]))
    os.makedirs(os.path.join(root, "optimizer"), exist_ok=True)
    _atomic_save(optimizer.state_dict() if optimizer is not None else {}, os.path.join(root, "optimizer", "%d.pt" % rank))
    be.barrier_all()


def _atomic_save(obj, path):
    """Write next to the target and rename: a failure never leaves a truncated file under the final name."""
    tmp = "%s.tmp.%d" % (path, os.getpid())
    torch.save(obj, tmp)
    os.replace(tmp, path)


def _wrapped_block(module):
    """-> (block, key prefix).  The block the reference wraps (and names checkpoint keys relative to): the decoder layer inside
    ``LlamaLayers_``, the module itself for embedding / final norm / head.  A row whose (tp|sp, cp) differs from its
    predecessor's is wrapped in ``Module_with_relocation`` (one ``module.`` level per wrapper, parallel.py:279-313): the
    reference names keys relative to the FSDP-wrapped block, so the wrappers are transparent."""
    prefix = ""
    while hasattr(module, "groups") and hasattr(module, "module"):      # Module_with_relocation
        module, prefix = module.module, prefix + "module."
    inner = getattr(module, "layer", None)
    if inner is not None and hasattr(inner, "idx"):
        return inner, prefix + "layer."
    return module, prefix
