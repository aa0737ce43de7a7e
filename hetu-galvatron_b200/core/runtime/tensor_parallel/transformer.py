"""Transformer-layer math around the per-layer collectives: ``ParallelMLP`` / ``ParallelAttention`` with explicit
``tp_group / sp_group / cp_group`` (``galvatron/core/runtime/tensor_parallel/transformer.py``: ParallelMLP :82-166,
ParallelAttention :512-900, Ulysses ``_SeqAllToAll`` / ``DistributedAttention`` :1990-2177) and the fused RMSNorm the
Llama family uses (``flash_attn.ops.rms_norm.RMSNorm``, LlamaModel_tensor_parallel.py:2,48).

Data flow of one attention block (SBH activations, flash layout inside):
    hidden [s,b,h] -> ColumnParallelLinear (tcgen05 GEMM, TP/SP comm in staging)           -> mixed [s,b,ng*(r+2)*hn]
    -> ONE kernel: QKV split + RoPE + SBH->BSND relayout (K/V stay un-expanded for GQA)        -> q,k,v
    -> [Ulysses: ONE pull all-to-all for q,k,v with the head/seq transpose folded in]
    -> attention LIBRARY call (cuDNN SDPA; the reference calls flash-attn) -> [Ulysses: inverse all-to-all] -> context [s,b,np*hn]
    -> RowParallelLinear (GEMM into staging -> all-reduce | reduce-scatter over NVLink)      -> out [s,b,h]
The reference runs split, repeat_interleave, 2x rope, 3x rearrange().contiguous() and, per Ulysses tensor, a permute
copy + NCCL all_to_all + a second permute copy (transformer.py:731-767,842-867,1934-1962).
"""
import enum
import math

import torch
import torch.nn as nn

from ..backend import get_backend
from .layers import ColumnParallelLinear, RowParallelLinear


class AttnType(enum.Enum):
    self_attn = 1
    cross_attn = 2


class AttnMaskType(enum.Enum):
    padding = 1
    causal = 2


def _size(group):
    return 1 if group is None else group.size


# ---------------------------------------------------------------------------------------------------------------
# fused elementwise autograd ops
# ---------------------------------------------------------------------------------------------------------------
class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, eps):
        y, rstd = get_backend().rmsnorm_fwd(x.contiguous(), weight, eps)
        ctx.save_for_backward(x, weight, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, rstd = ctx.saved_tensors
        dx, dw = get_backend().rmsnorm_bwd(dy.contiguous(), x.contiguous(), weight, rstd)
        return dx, dw, None


class RMSNorm(nn.Module):
    """y = x * rsqrt(mean(x^2) + eps) * weight, fp32 math, one rounding (flash_attn.ops.rms_norm semantics)."""

    def __init__(self, hidden_size, eps=1e-5, params_dtype=torch.float32, device=None, sequence_parallel=False):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.empty(hidden_size, dtype=params_dtype, device=device))
        # under Megatron-SP the norm sees only the local sequence slice: its grad is summed over the TP group on the
        # last microbatch (sp_grad_reduce.py:104-123)
        self._sequence_parallel = bool(sequence_parallel)
        setattr(self.weight, "sequence_parallel", self._sequence_parallel)
        if self.weight.device.type != "meta":
            self.reset_parameters()

    def reset_parameters(self):
        nn.init.ones_(self.weight)
        # (re)tag: materialising from the meta device replaces the Parameter object and drops custom attributes
        setattr(self.weight, "sequence_parallel", self._sequence_parallel)

    def forward(self, x):
        return _RMSNormFn.apply(x, self.weight, self.eps)


class _SwigluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate_up):
        gate_up = gate_up.contiguous()
        ctx.save_for_backward(gate_up)
        return get_backend().swiglu_fwd(gate_up)

    @staticmethod
    def backward(ctx, dy):
        (gate_up,) = ctx.saved_tensors
        return get_backend().swiglu_bwd(dy.contiguous(), gate_up)


class _QkvRopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mixed, cos, sin, ng, r, hn, stage_group):
        ctx.save_for_backward(cos, sin)
        ctx.dims = (ng, r, hn)
        q, k, v = get_backend().qkv_rope_fwd(mixed.contiguous(), cos, sin, ng, r, hn, stage_group)
        return q, k, v

    @staticmethod
    def backward(ctx, dq, dk, dv):
        cos, sin = ctx.saved_tensors
        ng, r, hn = ctx.dims
        return get_backend().qkv_rope_bwd(dq, dk, dv, cos, sin, ng, r, hn), None, None, None, None, None, None


class _UlyssesFn(torch.autograd.Function):
    """All tensors of one exchange in one launch; backward is the inverse exchange (transformer.py:2040-2062)."""

    @staticmethod
    def forward(ctx, group, to_heads, *tensors):
        ctx.group, ctx.to_heads = group, to_heads
        return tuple(get_backend().ulysses_all_to_all(list(tensors), group, to_heads))

    @staticmethod
    def backward(ctx, *grads):
        back = get_backend().ulysses_all_to_all([g.contiguous() for g in grads], ctx.group, not ctx.to_heads)
        return (None, None) + tuple(back)


class _FlashAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal, softmax_scale, key_mask=None):
        if key_mask is None:
            out, lse, rng = get_backend().attention_fwd(q, k, v, causal, softmax_scale)
        else:
            out, lse, rng = get_backend().attention_fwd(q, k, v, causal, softmax_scale, key_mask)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.causal, ctx.scale, ctx.rng = causal, softmax_scale, rng
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = get_backend().attention_bwd(dout, q, k, v, out, lse, ctx.causal, ctx.scale, ctx.rng)
        return dq, dk, dv, None, None, None


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        y, mean, rstd = get_backend().layernorm_fwd(x.contiguous(), weight, bias, eps)
        ctx.save_for_backward(x, weight, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd = ctx.saved_tensors
        dx, dw, db = get_backend().layernorm_bwd(dy.contiguous(), x.contiguous(), weight, mean, rstd)
        return dx, dw, db, None


class LayerNorm(nn.Module):
    """``torch.nn.LayerNorm`` of the GPT / BERT families (gpt_hf/GPTModel_tensor_parallel.py:34, bert_hf/BertModel_tensor_parallel.py)
    as one fused row kernel forward and one backward: y = (x - mean) * rstd * weight + bias, fp32 math, one rounding."""

    def __init__(self, hidden_size, eps=1e-5, params_dtype=torch.float32, device=None, sequence_parallel=False):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.empty(hidden_size, dtype=params_dtype, device=device))
        self.bias = nn.Parameter(torch.empty(hidden_size, dtype=params_dtype, device=device))
        self._sequence_parallel = bool(sequence_parallel)
        if self.weight.device.type != "meta":
            self.reset_parameters()

    def reset_parameters(self):
        nn.init.ones_(self.weight)
        nn.init.zeros_(self.bias)
        # under Megatron-SP the norm sees only the local sequence slice: its grads are summed over the TP group on the last
        # microbatch (sp_grad_reduce.py:104-123); (re)tag here, materialising from meta makes new Parameter objects
        setattr(self.weight, "sequence_parallel", self._sequence_parallel)
        setattr(self.bias, "sequence_parallel", self._sequence_parallel)

    def forward(self, x):
        return _LayerNormFn.apply(x, self.weight, self.bias, self.eps)


class _BiasGeluFn(torch.autograd.Function):
    """gelu(x + bias) of the GPT / BERT MLP in one pass (transformer.py:150-160 ``bias_gelu_impl``); dbias = column sums of dx."""

    @staticmethod
    def forward(ctx, x, bias, tanh_form):
        x = x.contiguous()
        ctx.save_for_backward(x, bias)
        ctx.tanh_form = tanh_form
        return get_backend().bias_gelu_fwd(x, bias, tanh_form)

    @staticmethod
    def backward(ctx, dy):
        x, bias = ctx.saved_tensors
        dx = get_backend().bias_gelu_bwd(dy.contiguous(), x, bias, ctx.tanh_form)
        db = None if bias is None else dx.reshape(-1, dx.shape[-1]).float().sum(0).to(bias.dtype)
        return dx, db, None


class _CpGatherKvFn(torch.autograd.Function):
    """Context parallelism, K or V: this rank's two zigzag chunks [b, s/c, ng, d] -> the whole sequence in natural token order
    [b, s, ng, d] (all-gather over the cp group + un-zigzag); backward re-applies the zigzag order and reduce-scatters the
    gradient (the reference's ring passes the same bytes around hop by hop, transformer.py:2252-2680)."""

    @staticmethod
    def forward(ctx, x, group):
        from ..redistribute import _reverse_zigzag_transformation
        b, s_loc, ng, d = x.shape
        ctx.group, ctx.dims = group, (b, s_loc, ng, d)
        rows = x.transpose(0, 1).contiguous().reshape(s_loc, b * ng * d)
        full = get_backend().all_gather_first_dim(rows, group)                       # rank-major chunk order
        full = _reverse_zigzag_transformation(full, group.size)
        return full.reshape(group.size * s_loc, b, ng, d).transpose(0, 1).contiguous()

    @staticmethod
    def backward(ctx, grad):
        from ..redistribute import _zigzag_transformation
        b, s_loc, ng, d = ctx.dims
        c = ctx.group.size
        rows = grad.transpose(0, 1).contiguous().reshape(c * s_loc, b * ng * d)
        rows = _zigzag_transformation(rows, c)
        out = get_backend().reduce_scatter_first_dim(rows, ctx.group)
        return out.reshape(s_loc, b, ng, d).transpose(0, 1).contiguous(), None


def _cp_attention(q, k, v, group, scale):
    """Causal self-attention under zigzag context parallelism: the rank holds token chunks (r, 2c-1-r); each chunk attends
    the gathered keys/values up to its own end."""
    be = get_backend()
    c, r = group.size, group.rank_in_group()
    k_full, v_full = _CpGatherKvFn.apply(k, group), _CpGatherKvFn.apply(v, group)
    half = q.shape[1] // 2
    outs = []
    for q_blk, chunk in ((q[:, :half], r), (q[:, half:], 2 * c - 1 - r)):
        end = (chunk + 1) * half
        outs.append(be.attention_prefix(q_blk, k_full[:, :end], v_full[:, :end], scale))
    return torch.cat(outs, 1)


def _recompute_activations():
    try:
        from ..arguments import get_args
        return bool(getattr(get_args(), "recompute_activations", False))
    except RuntimeError:
        return False


def _attention(q, k, v, causal, scale, key_mask=None):
    be = get_backend()
    fn = getattr(be, "attention", None)
    if fn is None:
        out = None
    elif key_mask is None:
        out = fn(q, k, v, causal, scale)                            # differentiable library call (cuDNN SDPA on B200)
    else:
        out = fn(q, k, v, causal, scale, key_mask)
    return out if out is not None else _FlashAttnFn.apply(q, k, v, causal, scale, key_mask)


# ---------------------------------------------------------------------------------------------------------------
# layer modules
# ---------------------------------------------------------------------------------------------------------------
class ParallelMLP(nn.Module):
    """h -> 4h (column-parallel) -> activation -> h (row-parallel) (transformer.py:82-166).
    Llama: gated (gate|up, swiglu), no biases.  GPT / BERT (``add_bias_linear``, not gated): bias + GeLU in one fused pass
    (``bias_gelu_impl``, :150-160), the output bias is returned for the caller to add (``skip_bias_add``, :162-166)."""

    def __init__(self, config, is_expert=False, tp_group=None, params_dtype=torch.float32, device=None):
        super().__init__()
        self.tp_group = tp_group
        ffn = config.ffn_hidden_size
        self.gated = getattr(config, "gated_linear_unit", True)
        self.add_bias = bool(getattr(config, "add_bias_linear", False))
        self.gelu_tanh = bool(getattr(config, "gelu_tanh", True))
        self.dense_h_to_4h = ColumnParallelLinear(config.hidden_size, ffn * 2 if self.gated else ffn, config=config, bias=self.add_bias,
                                                  gather_output=False, skip_bias_add=True, tp_group=tp_group,
                                                  params_dtype=params_dtype, device=device)
        self.dense_4h_to_h = RowParallelLinear(ffn, config.hidden_size, config=config, bias=self.add_bias, input_is_parallel=True,
                                               skip_bias_add=True, tp_group=tp_group, params_dtype=params_dtype, device=device)

    def forward(self, hidden_states, input_recipe=None, residual=None):
        """``residual`` (optional): the block's residual, added to the output inside the last GEMM's epilogue."""
        gate_up, bias = self.dense_h_to_4h(hidden_states, recompute=input_recipe)
        if self.gated:
            if bias is not None:
                gate_up = gate_up + bias
            inter = _SwigluFn.apply(gate_up)
            # --recompute_activations: the 4h->h GEMM keeps gate_up (which SwiGLU's own backward holds anyway) instead of the
            # SwiGLU output and redoes the elementwise pass in backward
            recipe = ("swiglu", gate_up) if _recompute_activations() else None
            return self.dense_4h_to_h(inter, recompute=recipe, residual=residual)
        inter = _BiasGeluFn.apply(gate_up, bias, self.gelu_tanh)
        return self.dense_4h_to_h(inter, residual=residual)


class ParallelAttention(nn.Module):
    """Self-attention with TP heads or Ulysses sequence parallelism (transformer.py:512-900)."""

    def __init__(self, config, layer_number, attention_type=AttnType.self_attn, attn_mask_type=AttnMaskType.padding,
                 tp_group=None, sp_group=None, cp_group=None, cp_ranks=None, use_ulysses=False, use_zigzag_cp=False,
                 params_dtype=torch.float32, device=None):
        super().__init__()
        if attention_type != AttnType.self_attn:
            raise NotImplementedError("only self attention is on the Galvatron hot path")
        self.use_cp = bool(use_zigzag_cp) or _size(cp_group) > 1
        if self.use_cp and use_ulysses and _size(sp_group) > 1:
            raise NotImplementedError("context parallelism together with Ulysses on the same layer is not supported")
        self.layer_number = max(1, layer_number)
        self.attn_mask_type = attn_mask_type
        self.tp_group, self.sp_group, self.cp_group = tp_group, sp_group, cp_group
        self.use_ulysses = use_ulysses and _size(sp_group) > 1
        world = _size(tp_group)
        self.hn = getattr(config, "kv_channels", None) or config.hidden_size // config.num_attention_heads
        n_heads = config.num_attention_heads
        n_groups = getattr(config, "num_query_groups", None) or n_heads
        assert n_heads % world == 0 and n_groups % world == 0, \
            "num_attention_heads / num_query_groups must be divisible by the tensor parallel size"   # :577-581
        if self.use_ulysses:
            assert n_heads % sp_group.size == 0, "num_attention_heads must be divisible by the Ulysses degree"  # :642
        self.np_local, self.ng_local = n_heads // world, n_groups // world
        self.r = self.np_local // self.ng_local
        add_bias = bool(getattr(config, "add_bias_linear", False))       # GPT / BERT: biases on both projections (:600-640)
        self.query_key_value = ColumnParallelLinear(config.hidden_size, (n_heads + 2 * n_groups) * self.hn, config=config,
                                                    bias=add_bias, gather_output=False, tp_group=tp_group,
                                                    params_dtype=params_dtype, device=device)
        self.dense = RowParallelLinear(n_heads * self.hn, config.hidden_size, config=config, bias=add_bias, skip_bias_add=True,
                                       input_is_parallel=True, tp_group=tp_group, params_dtype=params_dtype, device=device)
        self.softmax_scale = 1.0 / math.sqrt(self.hn)
        self._identity_rope = {}

    def _no_rope(self, seq, device):
        """Families with learned absolute positions (GPT, BERT) run the same split + relayout kernel with cos = 1, sin = 0."""
        key = (seq, str(device))
        if key not in self._identity_rope:
            self._identity_rope[key] = (torch.ones(seq, self.hn // 2, dtype=torch.float32, device=device),
                                        torch.zeros(seq, self.hn // 2, dtype=torch.float32, device=device))
        return self._identity_rope[key]

    def forward(self, hidden_states, attention_mask=None, encoder_output=None, inference_params=None, rotary_pos_emb=None,
                input_recipe=None, residual=None):
        # hidden_states [sq, b, h]; rotary_pos_emb = (cos, sin) fp32 tables [sq_local, hn/2] for this rank's positions
        mixed, _ = self.query_key_value(hidden_states, recompute=input_recipe)   # [s, b, ng*(r+2)*hn]
        cos, sin = rotary_pos_emb if rotary_pos_emb is not None else self._no_rope(mixed.shape[0], mixed.device)
        stage_group = self.sp_group if self.use_ulysses else None
        q, k, v = _QkvRopeFn.apply(mixed, cos, sin, self.ng_local, self.r, self.hn, stage_group)
        causal = self.attn_mask_type == AttnMaskType.causal
        # padding mask (BERT): [b, s] bool over the keys, True = attend (the reference builds the extended [b,1,1,s] additive
        # mask in bert_hf/BertModel_sequential.py and hands it to every layer)
        key_mask = attention_mask if (not causal and attention_mask is not None) else None
        if self.use_ulysses:
            p = self.sp_group.size
            if self.ng_local % p:  # too few KV heads to scatter: expand as the reference does (:842-848)
                rep = self.np_local // self.ng_local
                k, v = k.repeat_interleave(rep, dim=2), v.repeat_interleave(rep, dim=2)
            q, k, v = _UlyssesFn.apply(self.sp_group, True, q, k, v)       # [b, s, n/p, hn]
            ctxt = _attention(q, k, v, causal, self.softmax_scale, key_mask)
            (ctxt,) = _UlyssesFn.apply(self.sp_group, False, ctxt)         # [b, s/p, n, hn]
        elif self.use_cp:
            assert causal, "context parallelism is implemented for causal self-attention"
            ctxt = _cp_attention(q, k, v, self.cp_group, self.softmax_scale)   # [b, s/c, np, hn]
        else:
            ctxt = _attention(q, k, v, causal, self.softmax_scale, key_mask)  # [b, s, np, hn]
        b, s = ctxt.shape[0], ctxt.shape[1]
        ctxt = ctxt.reshape(b, s, -1).transpose(0, 1).contiguous()          # "b s h d -> s b (h d)"
        return self.dense(ctxt, residual=residual)
