"""Import shim (test infrastructure only): megatron/core/transformer/custom_layers/transformer_engine.py imports Transformer Engine at
module level and subclasses five of its layers; Galvatron's hot path never instantiates them."""
from . import pytorch  # noqa: F401
