"""The Llama family harness (``galvatron/models/llama_hf``): the three callbacks + ModelInfo the core API asks for."""
from .LlamaModel_checkpoint import load_llama_module, save_llama_module
from .LlamaModel_hybrid_parallel import (construct_hybrid_parallel_model, get_hybrid_parallel_configs, get_llama_config,
                                         llama_model_hp)
from .LlamaModel_sequential import LlamaModelInfo, construct_sequential_model
from .LlamaModel_tensor_parallel import LlamaLayer_tp, construct_tensor_parallel_model
from .meta_configs import config_from_meta, set_model_config
