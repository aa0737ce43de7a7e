"""Operator API consumed by the model families (``galvatron/core/runtime/tensor_parallel/__init__.py:1-11`` and
``megatron.core.tensor_parallel``)."""
from .cross_entropy import vocab_parallel_cross_entropy
from .layers import (ColumnParallelLinear, RowParallelLinear, VocabParallelEmbedding, VocabUtility,
                     linear_with_grad_accumulation_and_async_allreduce)
from .mappings_group import (copy_to_tensor_model_parallel_region_group, gather_from_sequence_parallel_region_group,
                             gather_from_tensor_model_parallel_region_group, get_tensor_model_parallel_rank_group,
                             get_tensor_model_parallel_world_size_group, reduce_from_tensor_model_parallel_region_group,
                             reduce_scatter_to_sequence_parallel_region_group, scatter_to_sequence_parallel_region_group,
                             scatter_to_tensor_model_parallel_region_group)
from .transformer import AttnMaskType, AttnType, LayerNorm, ParallelAttention, ParallelMLP, RMSNorm


def colummn_row_reset_parameters(self):
    """``tensor_parallel/reset.py:10-17``: N(0, init_method_std) weights, zero bias."""
    return self.reset_parameters()
