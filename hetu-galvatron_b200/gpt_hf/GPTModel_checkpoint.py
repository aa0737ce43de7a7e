"""Layer-wise checkpoint loading for the GPT family (``galvatron/models/gpt_hf/GPTModel_checkpoint.py:17-139``): the HF-layered
files ``galvatron/tools/checkpoint_convert_h2g.py:6-41`` writes from a HuggingFace GPT-2 checkpoint --
``transformer_embedding.pt`` {``wte.weight``, ``wpe.weight``[, ``weight`` = lm_head]}, ``transformer_h_<i>.pt`` with the HF keys of one
block, ``transformer_ln_f.pt`` -- sliced for the caller's tensor-parallel rank.  GPT-2's ``Conv1D`` weights are [in, out]: every
projection is transposed; ``c_attn`` goes from HF's (three, heads, head_dim) to Megatron's per-head (heads, three, head_dim) order.
The head is loaded from ``wte.weight`` as the reference does (HF ties them; this runtime then trains them untied).

The loader is the ``load_module_func`` callback ``(load, tp_groups, name, submodule, module, distributed_checkpoint)``.

Divergences, on purpose: vocabulary padding rows go to the END of wte / lm_head (the reference's ``F.pad(w, (0, 0, pad, 0))`` :29-34
prepends them, which shifts every token's row); the distributed format is not written for this family (the reference raises too)."""
import os

import torch

from ..core.runtime.arguments import get_args
from ..core.runtime.backend import get_backend
from ..llama_hf.LlamaModel_checkpoint import _pad_vocab, _put, _range, _read, _tp

embedding_name = "transformer_embedding.pt"
layer_name = "transformer_h_%d.pt"
ln_f_name = "transformer_ln_f.pt"
cls_name = "transformer_embedding.pt"


def _per_head(t, n_heads):
    """HF c_attn output order (three, heads, head_dim) -> Megatron's (heads, three, head_dim), along dim 0"""
    hn = t.shape[0] // (3 * n_heads)
    return t.reshape(3, n_heads, hn, *t.shape[1:]).transpose(0, 1).reshape(t.shape)


def load_hf_checkpoint(load, tp_groups, name, submodule, module):
    args = get_args()
    index, size = _tp(tp_groups, get_backend().rank)
    if name.endswith("wte") or name.endswith("lm_head"):
        w = _pad_vocab(_read(os.path.join(load, embedding_name))["wte.weight"].float(), args.padded_vocab_size)
        lo, hi = _range(args.padded_vocab_size, index, size)
        return _put(submodule.weight, w[lo:hi])
    if name.endswith("wpe"):
        w = _read(os.path.join(load, embedding_name))["wpe.weight"].float()
        if w.shape[0] < args.seq_length:
            raise ValueError("checkpoint has %d positions, the run needs %d" % (w.shape[0], args.seq_length))
        lo, hi = _range(args.seq_length, index, size)
        return _put(submodule.weight, w[lo:hi])
    if name.endswith("ln_f"):
        ckpt = _read(os.path.join(load, ln_f_name))
        _put(submodule.weight, ckpt["weight"].float())
        return _put(submodule.bias, ckpt["bias"].float())
    ckpt = _read(os.path.join(load, layer_name % module.idx))
    part = "ln_1" if name.startswith("attention") else "ln_2"
    if name.endswith("LayerNorm"):
        _put(submodule.weight, ckpt[part + ".weight"].float())
        return _put(submodule.bias, ckpt[part + ".bias"].float())
    if name.endswith("query_key_value"):
        w = _per_head(ckpt["attn.c_attn.weight"].float().t().contiguous(), args.num_attention_heads)
        b = _per_head(ckpt["attn.c_attn.bias"].float(), args.num_attention_heads)
        lo, hi = _range(w.shape[0], index, size)
        _put(submodule.weight, w[lo:hi])
        return _put(submodule.bias, b[lo:hi])
    key = {"dense": "attn.c_proj", "dense_h_to_4h": "mlp.c_fc", "dense_4h_to_h": "mlp.c_proj"}.get(name.rsplit(".", 1)[-1])
    if key is None:
        raise KeyError("no checkpoint rule for submodule %r of %s" % (name, type(module).__name__))
    w, b = ckpt[key + ".weight"].float().t().contiguous(), ckpt[key + ".bias"].float()          # [out, in]
    if key == "mlp.c_fc":                     # column-parallel: output rows and their bias are sliced
        lo, hi = _range(w.shape[0], index, size)
        _put(submodule.weight, w[lo:hi])
        return _put(submodule.bias, b[lo:hi])
    lo, hi = _range(w.shape[1], index, size)  # row-parallel: input columns are sliced, the bias is whole
    _put(submodule.weight, w[:, lo:hi])
    return _put(submodule.bias, b)


@torch.no_grad()
def load_gpt_module(load, tp_groups, name, submodule, module, distributed_checkpoint):
    """``GPTModel_checkpoint.py:134-139``."""
    if distributed_checkpoint:
        raise NotImplementedError("Distributed checkpoint is not supported for GPT")
    load_hf_checkpoint(load, tp_groups, name, submodule, module)
