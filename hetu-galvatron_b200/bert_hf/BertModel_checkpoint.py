"""Layer-wise checkpoint loading for the BERT family (``galvatron/models/bert_hf/BertModel_checkpoint.py``): the HF-layered files
``galvatron/tools/checkpoint_convert_h2g.py:84-130`` writes from a HuggingFace ``BertForMaskedLM`` checkpoint --
``bert_embeddings.pt``, ``bert_encoder_layer_<i>.pt`` with the HF keys of one layer, ``cls_predictions.pt`` -- sliced for the caller's
tensor-parallel rank.  HF keeps q / k / v as three [h, h] ``nn.Linear`` weights ([out, in]); the fused weight is Megatron's per-head
(heads, three, head_dim) order.

The loader is the ``load_module_func`` callback ``(load, tp_groups, name, submodule, module, distributed_checkpoint)``.

Divergences, on purpose (the reference's loader at HEAD cannot reproduce HF's model: it concatenates q | k | v without the per-head
interleave the attention kernel assumes, transposes the ``nn.Linear`` weights as if they were GPT-2 ``Conv1D``s, :147-178, and lacks the
callback's sixth argument): weights are placed so that the loaded model IS the HF model (bit-exact tensors, HF's loss within 5e-3,
tests/test_checkpoint_families.py); vocabulary padding rows go to the end."""
import os

import torch

from ..core.runtime.arguments import get_args
from ..core.runtime.backend import get_backend
from ..llama_hf.LlamaModel_checkpoint import _pad_vocab, _put, _range, _read, _tp

embedding_name = "bert_embeddings.pt"
layer_name = "bert_encoder_layer_%d.pt"
cls_name = "cls_predictions.pt"


def _fuse_heads(q, k, v, n_heads):
    """three [n*hn, ...] tensors -> [(n, three, hn), ...]"""
    hn = q.shape[0] // n_heads
    parts = [t.reshape(n_heads, hn, *t.shape[1:]) for t in (q, k, v)]
    return torch.stack(parts, 1).reshape(3 * q.shape[0], *q.shape[1:])


def load_hf_checkpoint(load, tp_groups, name, submodule, module):
    args = get_args()
    index, size = _tp(tp_groups, get_backend().rank)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "word_embeddings":
        w = _pad_vocab(_read(os.path.join(load, embedding_name))["word_embeddings.weight"].float(), args.padded_vocab_size)
        lo, hi = _range(args.padded_vocab_size, index, size)
        return _put(submodule.weight, w[lo:hi])
    if leaf == "position_embeddings":
        w = _read(os.path.join(load, embedding_name))["position_embeddings.weight"].float()
        if w.shape[0] < args.seq_length:
            raise ValueError("checkpoint has %d positions, the run needs %d" % (w.shape[0], args.seq_length))
        lo, hi = _range(args.seq_length, index, size)
        return _put(submodule.weight, w[lo:hi])
    if leaf == "token_type_embeddings":
        return _put(submodule.weight, _read(os.path.join(load, embedding_name))["token_type_embeddings.weight"].float())
    if leaf == "transform":                   # MLM head: dense + GeLU (the LayerNorm below is its own submodule)
        ckpt = _read(os.path.join(load, cls_name))
        _put(submodule.weight, ckpt["transform.dense.weight"].float())
        return _put(submodule.bias, ckpt["transform.dense.bias"].float())
    if leaf == "lm_head":
        ckpt = _read(os.path.join(load, cls_name))
        w = _pad_vocab(ckpt["decoder.weight"].float(), args.padded_vocab_size)
        b = _pad_vocab(ckpt["decoder.bias"].float()[:, None], args.padded_vocab_size)[:, 0]
        lo, hi = _range(args.padded_vocab_size, index, size)
        _put(submodule.weight, w[lo:hi])
        return _put(submodule.bias, b[lo:hi])
    if leaf == "LayerNorm":
        if not hasattr(module, "idx"):        # the embedding's LayerNorm or the MLM transform's
            if "transform" in name:
                ckpt, key = _read(os.path.join(load, cls_name)), "transform.LayerNorm"
            else:
                ckpt, key = _read(os.path.join(load, embedding_name)), "LayerNorm"
        else:                                 # post-LN blocks: after the attention residual / after the MLP residual
            ckpt = _read(os.path.join(load, layer_name % module.idx))
            key = "attention.output.LayerNorm" if name.startswith("attention") else "output.LayerNorm"
        _put(submodule.weight, ckpt[key + ".weight"].float())
        return _put(submodule.bias, ckpt[key + ".bias"].float())
    ckpt = _read(os.path.join(load, layer_name % module.idx))
    if leaf == "query_key_value":
        n = args.num_attention_heads
        w = _fuse_heads(*[ckpt["attention.self.%s.weight" % k].float() for k in ("query", "key", "value")], n)
        b = _fuse_heads(*[ckpt["attention.self.%s.bias" % k].float() for k in ("query", "key", "value")], n)
        lo, hi = _range(w.shape[0], index, size)
        _put(submodule.weight, w[lo:hi])
        return _put(submodule.bias, b[lo:hi])
    key = {"dense": "attention.output.dense", "dense_h_to_4h": "intermediate.dense", "dense_4h_to_h": "output.dense"}.get(leaf)
    if key is None:
        raise KeyError("no checkpoint rule for submodule %r of %s" % (name, type(module).__name__))
    w, b = ckpt[key + ".weight"].float(), ckpt[key + ".bias"].float()            # nn.Linear: [out, in]
    if leaf == "dense_h_to_4h":
        lo, hi = _range(w.shape[0], index, size)
        _put(submodule.weight, w[lo:hi])
        return _put(submodule.bias, b[lo:hi])
    lo, hi = _range(w.shape[1], index, size)
    _put(submodule.weight, w[:, lo:hi])
    return _put(submodule.bias, b)


@torch.no_grad()
def load_bert_module(load, tp_groups, name, submodule, module, distributed_checkpoint=False):
    if distributed_checkpoint:
        raise NotImplementedError("Distributed checkpoint is not supported for BERT")
    load_hf_checkpoint(load, tp_groups, name, submodule, module)
