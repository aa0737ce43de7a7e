"""Entry points of the GPT family (``galvatron/models/gpt_hf/GPTModel_hybrid_parallel.py``)."""
from ..core.runtime.hybrid_parallel_config import get_hybrid_parallel_configs_api
from ..core.runtime.hybrid_parallel_model import construct_hybrid_parallel_model_api
from ..llama_hf.LlamaModel_hybrid_parallel import estimate_arena_bytes
from .GPTModel_checkpoint import load_gpt_module
from .GPTModel_sequential import (GPTCls_, GPTEmbeddings_, GPTModelInfo, GPTPreNorm_, construct_sequential_model)
from .GPTModel_tensor_parallel import GPTLayer_tp, GPTSkeleton, construct_tensor_parallel_model
from .meta_configs import config_from_meta, set_model_config


def get_hybrid_parallel_configs(model_config, training_args):
    return get_hybrid_parallel_configs_api(model_config, training_args, GPTModelInfo)


def construct_hybrid_parallel_model(model, model_config, training_args, hybrid_parallel_configs):
    # GPTModel_hybrid_parallel.py:42: wte and lm_head are tied unless --untie_embeddings_and_output_weights (this runtime's default: untied)
    tied = None if getattr(training_args, "untie_embeddings_and_output_weights", True) else ["wte", ""]
    return construct_hybrid_parallel_model_api(
        model, model_config, training_args, hybrid_parallel_configs, GPTModelInfo, construct_sequential_model,
        construct_tensor_parallel_model, wrap_block_name=[GPTLayer_tp], wrap_checkpoint_block_name=[GPTLayer_tp],
        wrap_other_block_name=[GPTEmbeddings_, GPTPreNorm_, GPTCls_], tied_wte_attr_names=tied, layernorm_name=["LayerNorm", "ln_f"],
        all_block_name=[GPTEmbeddings_, GPTLayer_tp, GPTPreNorm_, GPTCls_], load_module_func=load_gpt_module)


def get_gpt_config(args, overwrite_args=True):
    return set_model_config(config_from_meta(args.model_size), args, overwrite_args)


def gpt_model_hp(config, args):
    hybrid_parallel_configs = get_hybrid_parallel_configs(model_config=config, training_args=args)
    if not getattr(args, "arena_bytes", 0):
        args.arena_bytes = estimate_arena_bytes(config, args, hybrid_parallel_configs)
    return construct_hybrid_parallel_model(model=GPTSkeleton(config), model_config=config, training_args=args,
                                           hybrid_parallel_configs=hybrid_parallel_configs)
