"""Microbatch chunking (``galvatron/core/runtime/pipeline/utils.py:12-64``): tensors are split along dim 0 with
``Tensor.chunk`` (so the last microbatch may be smaller and fewer than ``chunks`` may be produced), everything else is
replicated; kwargs named ``*_mask`` with a leading dimension of 1 are broadcast rather than chunked."""
import torch


def listify_model(model):
    return model if isinstance(model, list) else [model]


def _chunk_values(values, chunks, is_chunked):
    pieces = [v.chunk(chunks) if is_chunked(v, i) else None for i, v in enumerate(values)]
    counts = {len(p) for p in pieces if p is not None}
    if len(counts) > 1:
        raise RuntimeError("Found different number of chunks produced for inputs: %s" % sorted(counts))
    n = counts.pop() if counts else chunks
    return [[v if p is None else p[i] for p, v in zip(pieces, values)] for i in range(n)]


def chunk_batch(inputs, chunks):
    if inputs is None:
        return inputs
    return _chunk_values(list(inputs), chunks, lambda v, i: torch.is_tensor(v))


def chunk_dict(kwargs, chunks):
    keys = list(kwargs)
    broadcast_mask = lambda k, v: k.endswith("_mask") and v.shape[0] == 1  # noqa: E731
    rows = _chunk_values([kwargs[k] for k in keys], chunks,
                         lambda v, i: torch.is_tensor(v) and not broadcast_mask(keys[i], v))
    return [dict(zip(keys, row)) for row in rows]
