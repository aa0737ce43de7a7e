"""CPU restatement of the per-layer collectives' semantics (TEST INFRASTRUCTURE ONLY -- never imported by the
product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may use it).

Each function takes the list of per-rank inputs of ONE group and returns the list of per-rank outputs the
reference runtime would produce, following the cited reference lines.  Plain torch on CPU tensors.

Parity status: the reference's own tests never pin these per-collective numerics (loss-vs-HF only),
and NCCL's internal summation order is unspecified, so for reductions two answers are produced:
  * ``order="reference"``: rounds to the reduce dtype at every point the reference rounds (pre-divide, the
    reduced result, post-divide), summing ranks in ring order 0..n-1 in that dtype;
  * ``order="exact"``: float64 math, one final rounding -- what an ideal reduction returns.
The CUDA kernels accumulate in fp32, so they must match "exact" to fp32 accuracy and "reference" to within
the reduce dtype's rounding (bf16: 2^-8 relative per element).  "parity unpinned" by golden vectors for
these numerics; the strategy->group MAPPING is pinned bit-exact separately (tests/golden/comm_groups.json).
"""
import torch


def fsdp_divide_factors(world_size):
    """torch/distributed/algorithms/_comm_hooks/default_hooks.py:38-42 (DefaultState._get_gradient_predivide_factor):
    predivide = smallest power of two with world % f == 0 and world / f <= f;  postdivide = world / predivide."""
    factor = 1
    while world_size % factor == 0 and world_size / factor > factor:
        factor *= 2
    return float(factor), world_size / float(factor)


def pad_to_multiple(numel, multiple):
    """FSDP pads the flat parameter so every rank's shard is equal (torch/.../_flat_param.py, _runtime_utils.py:896-901)."""
    return (numel + multiple - 1) // multiple * multiple


def all_gather_cast(shards, dst_dtype=torch.bfloat16):
    """C1: torch/distributed/fsdp/_flat_param.py:1477 all_gather_into_tensor of the MixedPrecision-cast shard
    (galvatron/core/runtime/parallel.py:116-122: fp32 master shard -> param_dtype before the gather)."""
    full = torch.cat([s.to(dst_dtype) for s in shards])
    return [full.clone() for _ in shards]


def reduce_scatter_acc(srcs, dst_prev, reduce_dtype=torch.bfloat16, out_dtype=torch.float32, order="reference",
                       prescale=None, postscale=None, accumulate=True):
    """C2: torch/distributed/fsdp/_runtime_utils.py:831-926 _reduce_grad + _post_reduce_grad_callback.

    srcs: per-rank unsharded (padded) gradient in the reduce dtype; dst_prev: per-rank fp32 shard (or None).
      :852  _div_if_needed(padded_unsharded_grad, predivide)      (in reduce dtype)
      :858  dist.reduce_scatter_tensor(new_sharded_grad, padded_unsharded_grad)   (SUM, reduce dtype)
      :879  _div_if_needed(new_sharded_grad, postdivide)
      :917-924  cast to the param dtype (fp32) and ``_saved_grad_shard += sharded_grad``
    """
    n = len(srcs)
    pre, post = fsdp_divide_factors(n)
    pre = 1.0 / prescale if prescale is not None else pre
    post = 1.0 / postscale if postscale is not None else post
    shard = srcs[0].numel() // n
    outs = []
    for r in range(n):
        sl = slice(r * shard, (r + 1) * shard)
        if order == "exact":
            acc = torch.zeros(shard, dtype=torch.float64)
            for s in srcs:
                acc += s[sl].double() / pre
            red = (acc / post).to(out_dtype if dst_prev is None or not accumulate else torch.float64)
        else:
            acc = None
            for s in srcs:
                term = (s[sl].to(reduce_dtype) / pre).to(reduce_dtype)
                acc = term if acc is None else (acc.float() + term.float()).to(reduce_dtype)
            red = (acc / post).to(reduce_dtype).to(out_dtype)
        if accumulate and dst_prev is not None:
            red = (dst_prev[r].double() + red.double()).to(out_dtype) if order == "exact" else dst_prev[r] + red.to(out_dtype)
        outs.append(red.to(out_dtype))
    return outs


def all_reduce(srcs, op="sum", order="exact", scale=1.0):
    """C3/C5/C6/C13: dist.all_reduce at mappings_group.py:19 (_reduce), _runtime_utils.py:940, cross_entropy.py:22-30
    (MAX), :61-72/:78-89 (SUM).  Every rank receives the same tensor."""
    dtype = srcs[0].dtype
    if op == "max":
        out = srcs[0].clone()
        for s in srcs[1:]:
            out = torch.maximum(out, s)
    elif order == "exact":
        acc = torch.zeros_like(srcs[0], dtype=torch.float64)
        for s in srcs:
            acc += s.double()
        out = (acc * scale).to(dtype)
    else:
        acc = srcs[0].clone()
        for s in srcs[1:]:
            acc = (acc.float() + s.float()).to(dtype)
        out = (acc.float() * scale).to(dtype)
    return [out.clone() for _ in srcs]


def ulysses_all_to_all(inputs, scatter_idx, gather_idx):
    """C10: galvatron/core/runtime/tensor_parallel/transformer.py:1928-1987 single_all_to_all with
    batch_dim_idx == 0 ([b, s, n, d] tensors) + post_all2all :1904-1925, restated for the whole group at once.

    scatter_idx=2, gather_idx=1: [b, s/p, n, d] -> [b, s, n/p, d]   (q, k, v before attention)
    scatter_idx=1, gather_idx=2: [b, s, n/p, d] -> [b, s/p, n, d]   (context after attention)
    """
    p = len(inputs)
    sends = []  # sends[r][q]: block rank r sends to rank q (the reference's permuted, contiguous input_t[q])
    for x in inputs:
        if scatter_idx < 2:
            b, s_glob, n_loc, d = x.shape
            t = x.reshape(b, p, s_glob // p, n_loc, d).permute(1, 0, 2, 3, 4).contiguous()
        else:
            b, s_loc, n_tot, d = x.shape
            t = x.reshape(b, s_loc, p, n_tot // p, d).permute(2, 0, 1, 3, 4).contiguous()
        sends.append(t)
    outs = []
    for r in range(p):
        recv = torch.stack([sends[q][r] for q in range(p)])  # all_to_all_single: chunk q comes from rank q
        if scatter_idx < 2:
            b, s_glob, n_loc, d = inputs[r].shape
            o = recv.permute(1, 2, 0, 3, 4).contiguous().reshape(b, s_glob // p, p * n_loc, d)
        else:
            b, s_loc, n_tot, d = inputs[r].shape
            o = recv.permute(1, 0, 2, 3, 4).contiguous().reshape(b, p * s_loc, n_tot // p, d)
        outs.append(o.contiguous())
    return outs


def rope_tables(seq_len, head_dim, base=10000.0, offset=0, dtype=torch.bfloat16):
    """megatron RotaryEmbedding.forward (rotary_pos_embedding.py): freqs = outer(pos, inv_freq); the cos/sin
    applied in apply_rotary_pos_emb are cast to the activation dtype before use."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    pos = torch.arange(seq_len, dtype=torch.float32) + offset
    freqs = torch.outer(pos, inv_freq)
    return torch.cos(freqs).to(dtype).float(), torch.sin(freqs).to(dtype).float()
