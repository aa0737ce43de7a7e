"""Tensor-parallel linear / embedding layers with an explicit ``tp_group`` -- the operator API of
``galvatron/site_package/megatron/core/tensor_parallel/layers.py`` (``VocabParallelEmbedding`` :166,
``LinearWithGradAccumulationAndAsyncCommunication`` :375, ``ColumnParallelLinear`` :651, ``RowParallelLinear`` :927).

What differs from the reference is where the bytes go, not the math:
  * the GEMMs are the tcgen05 kernel (``backend.gemm``), and a GEMM whose result is about to be reduced writes it
    straight into the TP group's peer-visible staging buffer, so the all-reduce / reduce-scatter kernel pulls it over
    NVLink with no intermediate copy (reference: cuBLAS, then a separate NCCL kernel on the same stream);
  * the Megatron-SP all-gather lands in that staging buffer too and is consumed in place by the GEMM (reference: a
    global scratch buffer, ``megatron/core/utils.py:62-80``);
  * wgrad accumulates directly into the layer's flat bf16 gradient buffer (``weight._bg_grad``) that the sharded
    data-parallel unit reduce-scatters -- Megatron's ``gradient_accumulation_fusion`` idea without apex.
"""
import math

import torch
import torch.nn as nn

from ..backend import get_backend
from .mappings_group import (_reduce, copy_to_tensor_model_parallel_region_group,
                             gather_from_tensor_model_parallel_region_group,
                             reduce_from_tensor_model_parallel_region_group,
                             reduce_scatter_to_sequence_parallel_region_group,
                             scatter_to_tensor_model_parallel_region_group)


def _size(group):
    return 1 if group is None else group.size


def _rank(group):
    return 0 if group is None or group.size == 1 else group.rank_in_group()


class VocabUtility:
    """``megatron/core/tensor_parallel/utils.py`` VocabUtility: contiguous [first, last) slice per rank."""

    @staticmethod
    def vocab_range_from_per_partition_vocab_size(per_partition, rank, world_size):
        return rank * per_partition, (rank + 1) * per_partition

    @staticmethod
    def vocab_range_from_global_vocab_size(global_size, rank, world_size):
        assert global_size % world_size == 0, "{} is not divisible by {}".format(global_size, world_size)
        return VocabUtility.vocab_range_from_per_partition_vocab_size(global_size // world_size, rank, world_size)


def _write_wgrad(weight, dy2d, x2d):
    """dW = dy^T x, accumulated into the flat gradient buffer when the weight belongs to a sharded unit."""
    be = get_backend()
    sink = getattr(weight, "_bg_grad", None)
    if sink is None:
        return be.gemm(dy2d, x2d, "nt")
    unit = weight._bg_unit
    be.gemm(dy2d, x2d, "nt", out=sink, accumulate=unit.grad_started(weight))
    unit.mark_grad(weight)
    return None


class LinearWithGradAccumulationAndAsyncCommunication(torch.autograd.Function):
    """y = x W^T with the tensor/sequence-parallel communication of layers.py:375-547.

    sequence_parallel: all-gather x along dim 0 before the GEMM (:399-413), re-gather in backward (:449-455) and
        reduce-scatter dgrad (:488-494).
    allreduce_dgrad:   all-reduce dgrad over the TP group (what ``copy_to_tensor_model_parallel_region`` does in
        backward, mappings_group.py:139) -- done here so the dgrad GEMM can write into the staging buffer.
    out_staged:        write y into the staging buffer (it is about to be all-reduced / reduce-scattered).
    recompute:         (opt-in, ``--recompute_activations``) instead of saving ``input`` for the wgrad GEMM, save what it was made
        from -- ("swiglu", gate_up) or ("rmsnorm", x, norm_weight, eps), tensors the producing op keeps anyway -- and redo that
        elementwise pass in backward: one layer then holds 352 MiB less at Llama-3-8B / seq 8192.
    """

    @staticmethod
    def forward(ctx, input, weight, sequence_parallel, allreduce_dgrad, out_staged, tp_group, reduce_scatter_out=False,
                recompute_kind=None, recompute_eps=0.0, *recipe):
        be = get_backend()
        ctx.recompute = (recompute_kind, recompute_eps, len(recipe)) if recompute_kind else None
        if ctx.recompute:
            ctx.save_for_backward(weight, *recipe)
        else:
            ctx.save_for_backward(input, weight)
        ctx.reduce_scatter_out = reduce_scatter_out and _size(tp_group) > 1
        if ctx.reduce_scatter_out:
            # row-parallel forward under Megatron-SP (layers.py:1061-1109): GEMM + reduce-scatter along the sequence
            ctx.sequence_parallel, ctx.allreduce_dgrad, ctx.tp_group = False, False, tp_group
            x2d = input.reshape(-1, input.shape[-1])
            n_out = weight.shape[0]
            if getattr(be, "can_fuse_gemm_rs", None) and be.can_fuse_gemm_rs(x2d.shape[0], n_out, tp_group, x2d.shape[1]):
                out = be.gemm_reduce_scatter(x2d, weight, "tn", tp_group)       # one fused operation
            else:
                staged, _ = be.staging_tensor(tp_group, (x2d.shape[0], n_out), input.dtype)
                be.gemm(x2d, weight, "tn", out=staged)
                out = be.reduce_scatter_first_dim(staged, tp_group)
            return out.view(input.shape[0] // tp_group.size, *input.shape[1:-1], n_out)
        ctx.sequence_parallel = sequence_parallel and _size(tp_group) > 1
        ctx.allreduce_dgrad = allreduce_dgrad and _size(tp_group) > 1
        ctx.tp_group = tp_group
        total = be.all_gather_into_staging(input, tp_group) if ctx.sequence_parallel else input
        x2d = total.reshape(-1, total.shape[-1])
        n_out = weight.shape[0]
        if out_staged and _size(tp_group) > 1 and not ctx.sequence_parallel:
            out, _ = be.staging_tensor(tp_group, (x2d.shape[0], n_out), input.dtype)
            be.gemm(x2d, weight, "tn", out=out)
        else:
            out = be.gemm(x2d, weight, "tn")
        return out.view(*total.shape[:-1], n_out)

    @staticmethod
    def backward(ctx, grad_output):
        be = get_backend()
        if ctx.recompute:
            kind, eps, n_recipe = ctx.recompute
            weight, recipe = ctx.saved_tensors[0], ctx.saved_tensors[1:]
            if kind == "swiglu":
                input = be.swiglu_fwd(recipe[0])
            elif kind == "rmsnorm":
                input, _ = be.rmsnorm_fwd(recipe[0], recipe[1], eps)
            else:
                raise ValueError("unknown recompute recipe %r" % (kind,))
        else:
            n_recipe = 0
            input, weight = ctx.saved_tensors
        group = ctx.tp_group
        if ctx.reduce_scatter_out:   # backward of the reduce-scatter is an all-gather along the sequence (mappings_group.py:243-258)
            grad_output = be.all_gather_first_dim(grad_output.contiguous(), group)
        dy2d = grad_output.reshape(-1, grad_output.shape[-1])
        if not dy2d.is_contiguous():
            dy2d = dy2d.contiguous()
        grad_weight = None
        if weight.requires_grad:
            total = be.all_gather_into_staging(input, group) if ctx.sequence_parallel else input
            grad_weight = _write_wgrad(weight, dy2d, total.reshape(-1, total.shape[-1]))
        grad_input = None
        if ctx.needs_input_grad[0]:
            m, k = dy2d.shape[0], weight.shape[1]
            if ctx.sequence_parallel and getattr(be, "can_fuse_gemm_rs", None) and be.can_fuse_gemm_rs(m, k, group, dy2d.shape[1]):
                # dgrad GEMM + reduce-scatter along the sequence (layers.py:462,488-494) as one fused operation
                out = be.gemm_reduce_scatter(dy2d, weight, "nn", group)
                grad_input = out.view(grad_output.shape[0] // group.size, *grad_output.shape[1:-1], k)
            elif ctx.sequence_parallel or ctx.allreduce_dgrad:
                staged, _ = be.staging_tensor(group, (m, k), dy2d.dtype)  # overwrites the gathered input: wgrad is done
                be.gemm(dy2d, weight, "nn", out=staged)
                full_shape = grad_output.shape[:-1] + (k,)
                if ctx.sequence_parallel:
                    grad_input = be.reduce_scatter_first_dim(staged.view(*full_shape), group)
                else:
                    grad_input = be.all_reduce(staged.view(*full_shape), group)
            else:
                grad_input = be.gemm(dy2d, weight, "nn").view(*grad_output.shape[:-1], k)
        return (grad_input, grad_weight, None, None, None, None, None, None, None) + (None,) * n_recipe


def linear_with_grad_accumulation_and_async_allreduce(input, weight, bias=None, gradient_accumulation_fusion=False,
                                                      async_grad_allreduce=False, sequence_parallel=False, tp_group=None,
                                                      out_staged=False, reduce_scatter_out=False, recompute=None):
    """Same call shape as layers.py:550-648 (``async_grad_allreduce`` here means "all-reduce dgrad over tp_group").
    ``recompute``: None, ("swiglu", gate_up) or ("rmsnorm", x, norm_weight, eps) -- see the Function's docstring."""
    if recompute is None:
        out = LinearWithGradAccumulationAndAsyncCommunication.apply(input, weight, sequence_parallel, async_grad_allreduce,
                                                                    out_staged, tp_group, reduce_scatter_out)
    else:
        kind = recompute[0]
        eps = float(recompute[3]) if kind == "rmsnorm" else 0.0
        recipe = recompute[1:3] if kind == "rmsnorm" else recompute[1:2]
        out = LinearWithGradAccumulationAndAsyncCommunication.apply(input, weight, sequence_parallel, async_grad_allreduce,
                                                                    out_staged, tp_group, reduce_scatter_out, kind, eps, *recipe)
    return out if bias is None else out + bias


class _ParallelLinearBase(nn.Module):
    def _make_weight(self, rows, cols, init_std, params_dtype, device):
        self.weight = nn.Parameter(torch.empty(rows, cols, dtype=params_dtype, device=device))
        self.init_std = init_std
        if self.weight.device.type != "meta":
            self.reset_parameters()

    def reset_parameters(self):
        """``colummn_row_reset_parameters`` (tensor_parallel/reset.py:10-17): N(0, init_method_std), zero bias."""
        nn.init.normal_(self.weight, mean=0.0, std=self.init_std)
        if getattr(self, "bias", None) is not None:
            nn.init.zeros_(self.bias)
            if isinstance(self, RowParallelLinear):
                setattr(self.bias, "sequence_parallel", self.sequence_parallel)  # layers.py:1045 (survives meta materialisation)


class ColumnParallelLinear(_ParallelLinearBase):
    """Y = XA with A split along its output dimension over ``tp_group`` (layers.py:651-910)."""

    def __init__(self, input_size, output_size, *, config=None, init_method=None, bias=False, gather_output=False,
                 skip_bias_add=False, tp_group=None, sp_group=None, cp_group=None, sequence_parallel=None,
                 init_std=0.02, params_dtype=torch.float32, device=None):
        super().__init__()
        self.input_size, self.output_size, self.gather_output = input_size, output_size, gather_output
        self.tp_group, self.sp_group, self.cp_group = tp_group, sp_group, cp_group
        world = _size(tp_group)
        assert output_size % world == 0, "{} is not divisible by {}".format(output_size, world)
        self.output_size_per_partition = output_size // world
        self.skip_bias_add = skip_bias_add
        sp = getattr(config, "sequence_parallel", False) if sequence_parallel is None else sequence_parallel
        self.sequence_parallel = bool(sp) and world > 1
        if config is not None:
            init_std = getattr(config, "init_method_std", init_std)
        self._make_weight(self.output_size_per_partition, input_size, init_std, params_dtype, device)
        if bias:
            self.bias = nn.Parameter(torch.zeros(self.output_size_per_partition, dtype=params_dtype, device=device))
        else:
            self.register_parameter("bias", None)

    def forward(self, input_, recompute=None):
        bias = self.bias if not self.skip_bias_add else None
        # without SP the input is replicated: dgrad must be all-reduced (copy_to_tensor_model_parallel_region, :875)
        out = linear_with_grad_accumulation_and_async_allreduce(
            input_, self.weight, bias, async_grad_allreduce=not self.sequence_parallel,
            sequence_parallel=self.sequence_parallel, tp_group=self.tp_group, recompute=recompute)
        if self.gather_output:
            assert not self.sequence_parallel
            out = gather_from_tensor_model_parallel_region_group(out, self.tp_group)
        return out, (self.bias if self.skip_bias_add else None)


class RowParallelLinear(_ParallelLinearBase):
    """Y = XA with A split along its input dimension; output all-reduced (or reduce-scattered under SP)
    (layers.py:927-1121)."""

    def __init__(self, input_size, output_size, *, config=None, init_method=None, bias=False, input_is_parallel=True,
                 skip_bias_add=False, tp_group=None, sp_group=None, cp_group=None, sequence_parallel=None,
                 init_std=0.02, params_dtype=torch.float32, device=None):
        super().__init__()
        self.input_size, self.output_size, self.input_is_parallel = input_size, output_size, input_is_parallel
        self.tp_group, self.sp_group, self.cp_group = tp_group, sp_group, cp_group
        world = _size(tp_group)
        assert input_size % world == 0, "{} is not divisible by {}".format(input_size, world)
        self.input_size_per_partition = input_size // world
        self.skip_bias_add = skip_bias_add
        sp = getattr(config, "sequence_parallel", False) if sequence_parallel is None else sequence_parallel
        self.sequence_parallel = bool(sp) and world > 1
        if self.sequence_parallel and not input_is_parallel:
            raise RuntimeError("To enable `sequence_parallel`, `input_is_parallel` must be `True`")
        if config is not None:
            init_std = getattr(config, "output_layer_init_std", getattr(config, "init_method_std", init_std))
        self._make_weight(output_size, self.input_size_per_partition, init_std, params_dtype, device)
        if bias:
            self.bias = nn.Parameter(torch.zeros(output_size, dtype=params_dtype, device=device))
            setattr(self.bias, "sequence_parallel", self.sequence_parallel)  # layers.py:1045
        else:
            self.register_parameter("bias", None)

    def forward(self, input_, recompute=None):
        if not self.input_is_parallel:
            input_ = scatter_to_tensor_model_parallel_region_group(input_, self.tp_group)
            recompute = None
        if self.sequence_parallel:
            # GEMM and the sequence reduce-scatter (:1109, C8) are one operation: partial tiles go straight to their owner
            out = linear_with_grad_accumulation_and_async_allreduce(
                input_, self.weight, None, async_grad_allreduce=False, sequence_parallel=False, tp_group=self.tp_group,
                reduce_scatter_out=True, recompute=recompute)
        else:
            out_parallel = linear_with_grad_accumulation_and_async_allreduce(
                input_, self.weight, None, async_grad_allreduce=False, sequence_parallel=False, tp_group=self.tp_group,
                out_staged=True, recompute=recompute)
            out = reduce_from_tensor_model_parallel_region_group(out_parallel, self.tp_group)     # :1114 (C5)
        if not self.skip_bias_add and self.bias is not None:
            out = out + self.bias
        return out, (self.bias if self.skip_bias_add else None)


class _VocabEmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, weight, vocab_start, vocab_end, masked):
        if masked:
            oob = (tokens < vocab_start) | (tokens >= vocab_end)
            idx = (tokens - vocab_start).masked_fill(oob, 0)
        else:
            oob, idx = None, tokens
        out = weight.index_select(0, idx.reshape(-1)).view(*tokens.shape, weight.shape[1])
        if masked:
            out.masked_fill_(oob.unsqueeze(-1), 0.0)
        ctx.save_for_backward(idx, oob if masked else torch.empty(0, device=tokens.device), weight)
        ctx.masked = masked
        return out

    @staticmethod
    def backward(ctx, grad_output):
        idx, oob, weight = ctx.saved_tensors
        g = grad_output.reshape(-1, grad_output.shape[-1])
        if ctx.masked:
            g = g.masked_fill(oob.reshape(-1, 1), 0.0)
        sink = getattr(weight, "_bg_grad", None)
        if sink is not None:
            unit = weight._bg_unit
            if not unit.grad_started(weight):
                sink.zero_()
            sink.index_add_(0, idx.reshape(-1), g.to(sink.dtype))
            unit.mark_grad(weight)
            return None, None, None, None, None
        dw = torch.zeros_like(weight)
        dw.index_add_(0, idx.reshape(-1), g.to(dw.dtype))
        return None, dw, None, None, None


class VocabParallelEmbedding(nn.Module):
    """Embedding split along the vocabulary over ``tp_group``; masked lookup + all-reduce (layers.py:166-262)."""

    def __init__(self, num_embeddings, embedding_dim, *, config=None, init_method=None, tp_group=None, sp_group=None,
                 cp_group=None, init_std=0.02, params_dtype=torch.float32, device=None):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.tp_group, self.sp_group, self.cp_group = tp_group, sp_group, cp_group
        world = _size(tp_group)
        self.vocab_start_index, self.vocab_end_index = VocabUtility.vocab_range_from_global_vocab_size(
            num_embeddings, _rank(tp_group), world)
        self.num_embeddings_per_partition = self.vocab_end_index - self.vocab_start_index
        if config is not None:
            init_std = getattr(config, "init_method_std", init_std)
        self.init_std = init_std
        self.weight = nn.Parameter(torch.empty(self.num_embeddings_per_partition, embedding_dim, dtype=params_dtype, device=device))
        if self.weight.device.type != "meta":
            self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.weight, mean=0.0, std=self.init_std)

    def forward(self, input_):
        masked = _size(self.tp_group) > 1
        out_parallel = _VocabEmbeddingFn.apply(input_, self.weight, self.vocab_start_index, self.vocab_end_index, masked)
        return reduce_from_tensor_model_parallel_region_group(out_parallel, self.tp_group)   # :261
