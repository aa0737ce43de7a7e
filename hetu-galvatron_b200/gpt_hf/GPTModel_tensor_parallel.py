"""GPT layer classes over the group-explicit parallel ops (``galvatron/models/gpt_hf/GPTModel_tensor_parallel.py``):
pre-LayerNorm blocks with biases on every projection, GeLU MLP, learned absolute positions (no RoPE)."""
import types

from torch import nn

from ..core.runtime.arguments import get_args
from ..core.runtime.tensor_parallel import (AttnMaskType, AttnType, ColumnParallelLinear, LayerNorm, ParallelAttention, ParallelMLP,
                                            VocabParallelEmbedding)


def core_transformer_config_from_args(args):
    """The ``TransformerConfig`` fields the layer code reads (megatron ``core_transformer_config_from_args``) for a GPT / BERT style
    block: biases on, GeLU (tanh form: ``bias_gelu_fusion``), no gating, no rotary embedding."""
    return types.SimpleNamespace(
        hidden_size=args.hidden_size, ffn_hidden_size=args.ffn_hidden_size, num_attention_heads=args.num_attention_heads,
        num_query_groups=args.num_attention_heads, kv_channels=args.hidden_size // args.num_attention_heads,
        layernorm_epsilon=args.norm_epsilon, init_method_std=args.init_method_std, sequence_parallel=args.sequence_parallel,
        gated_linear_unit=False, add_bias_linear=True, gelu_tanh=True)


def _megatron_sp(args, tp_group):
    return bool(args.sequence_parallel) and tp_group is not None and tp_group.size > 1


class GPTAttention_tp(nn.Module):
    def __init__(self, config, layer_number, tp_group=None, sp_group=None):
        super().__init__()
        args = get_args()
        self.use_ulysses = sp_group is not None and sp_group.size > 1
        mconf = core_transformer_config_from_args(args)
        self.tp_group = tp_group.group if tp_group is not None else None
        self.sp_group = sp_group.group if sp_group is not None else None
        self.attention = ParallelAttention(mconf, layer_number, attention_type=AttnType.self_attn, attn_mask_type=AttnMaskType.causal,
                                           tp_group=self.tp_group, sp_group=self.sp_group, use_ulysses=self.use_ulysses, device="meta")
        self.LayerNorm = LayerNorm(config.hidden_size, eps=config.layer_norm_epsilon, device="meta",
                                   sequence_parallel=_megatron_sp(args, tp_group))

    def forward(self, hidden_states, attention_mask):
        residual = hidden_states
        hidden_states = self.LayerNorm(hidden_states)
        # causal: the mask is implied (flash path, :36-41); the residual add (:42) rides in the projection GEMM's epilogue
        hidden_states, bias = self.attention(hidden_states, None, residual=residual)
        return hidden_states if bias is None else hidden_states + bias


class GPTMLP_tp(nn.Module):
    def __init__(self, config, tp_group=None):
        super().__init__()
        args = get_args()
        mconf = core_transformer_config_from_args(args)
        self.tp_group = tp_group.group if tp_group is not None else None
        self.mlp = ParallelMLP(mconf, tp_group=self.tp_group, device="meta")
        self.LayerNorm = LayerNorm(config.hidden_size, eps=config.layer_norm_epsilon, device="meta",
                                   sequence_parallel=_megatron_sp(args, tp_group))

    def forward(self, hidden_states):
        residual = hidden_states
        hidden_states = self.LayerNorm(hidden_states)
        hidden_states, bias = self.mlp(hidden_states, residual=residual)
        return hidden_states if bias is None else hidden_states + bias


class GPTLayer_tp(nn.Module):
    def __init__(self, config, layer_number, tp_group=None, sp_group=None):
        super().__init__()
        self.attention = GPTAttention_tp(config, layer_number, tp_group, sp_group)
        self.mlp = GPTMLP_tp(config, tp_group)
        self.idx = layer_number

    def forward(self, hidden_states, attention_mask=None):
        return self.mlp(self.attention(hidden_states, attention_mask))


class GPTSkeleton(nn.Module):
    """Container with the attribute layout of HF ``GPT2LMHeadModel`` (``.transformer.h/.wte/.wpe/.ln_f``, ``.lm_head``) that the
    reference's callbacks mutate; created empty -- every real layer is built by ``construct_tensor_parallel_model``."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.transformer = nn.Module()
        self.transformer.h = nn.ModuleList()
        self.transformer.wte = self.transformer.wpe = None
        self.transformer.ln_f = LayerNorm(config.hidden_size, eps=config.layer_norm_epsilon, device="meta")
        self.lm_head = None


def construct_tensor_parallel_model(model, config, tp_groups_enc, sp_groups_enc):
    """Whole-model rows: [embed, layer_0..L-1, norm, cls]; the 4-argument callback of the family
    (GPTModel_tensor_parallel.py:84-132 -- HEAD's core calls it with 5; this core accepts both)."""
    args = get_args()
    mconf = core_transformer_config_from_args(args)
    layers = nn.ModuleList([GPTLayer_tp(config, i, tp_group=tp_groups_enc[i + 1], sp_group=sp_groups_enc[i + 1])
                            for i in range(config.num_hidden_layers)])
    setattr(model.transformer, "h", layers)
    for name, rows in (("wte", args.padded_vocab_size), ("wpe", args.seq_length)):
        setattr(model.transformer, name, VocabParallelEmbedding(rows, mconf.hidden_size, config=mconf, tp_group=tp_groups_enc[0].group,
                                                                sp_group=sp_groups_enc[0].group, device="meta"))
    # the final norm sits in the "norm" row, whose degrees are the vocabulary's
    model.transformer.ln_f = LayerNorm(config.hidden_size, eps=config.layer_norm_epsilon, device="meta",
                                       sequence_parallel=bool(args.sequence_parallel) and args.vocab_tp > 1 and not args.vocab_sp)
    setattr(model, "lm_head", ColumnParallelLinear(mconf.hidden_size, args.padded_vocab_size, config=mconf, bias=False,
                                                   tp_group=tp_groups_enc[-1].group, sp_group=sp_groups_enc[-1].group, device="meta"))
    return model
