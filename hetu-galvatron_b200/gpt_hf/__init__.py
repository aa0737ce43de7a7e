"""The GPT family harness (``galvatron/models/gpt_hf``): the three callbacks + ModelInfo the core API asks for."""
from .GPTModel_hybrid_parallel import construct_hybrid_parallel_model, get_gpt_config, get_hybrid_parallel_configs, gpt_model_hp
from .GPTModel_sequential import GPTModelInfo, construct_sequential_model
from .GPTModel_tensor_parallel import GPTLayer_tp, construct_tensor_parallel_model
from .meta_configs import config_from_meta, set_model_config
from .GPTModel_checkpoint import load_gpt_module  # noqa: E402,F401
