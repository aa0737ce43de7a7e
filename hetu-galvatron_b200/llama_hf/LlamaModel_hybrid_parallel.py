"""Entry points of the family (``galvatron/models/llama_hf/LlamaModel_hybrid_parallel.py``)."""
from ..core.runtime.hybrid_parallel_config import get_hybrid_parallel_configs_api
from ..core.runtime.hybrid_parallel_model import construct_hybrid_parallel_model_api
from .LlamaModel_sequential import LlamaCls_, LlamaEmbeddings_, LlamaModelInfo, LlamaPreNorm_, construct_sequential_model
from .LlamaModel_tensor_parallel import LlamaLayer_tp, LlamaSkeleton, construct_tensor_parallel_model
from .meta_configs import config_from_meta, set_model_config


def get_hybrid_parallel_configs(model_config, training_args):
    return get_hybrid_parallel_configs_api(model_config, training_args, LlamaModelInfo)


def construct_hybrid_parallel_model(model, model_config, training_args, hybrid_parallel_configs):
    return construct_hybrid_parallel_model_api(
        model, model_config, training_args, hybrid_parallel_configs, LlamaModelInfo, construct_sequential_model,
        construct_tensor_parallel_model, wrap_block_name=[LlamaLayer_tp], wrap_checkpoint_block_name=[LlamaLayer_tp],
        wrap_other_block_name=[LlamaEmbeddings_, LlamaPreNorm_, LlamaCls_], layernorm_name=["LayerNorm", "norm"],
        all_block_name=[LlamaEmbeddings_, LlamaLayer_tp, LlamaPreNorm_, LlamaCls_])


def get_llama_config(args, overwrite_args=True):
    config = config_from_meta(args.model_size)
    return set_model_config(config, args, overwrite_args)


def llama_model_hp(config, args):
    hybrid_parallel_configs = get_hybrid_parallel_configs(model_config=config, training_args=args)
    skeleton = LlamaSkeleton(config)
    return construct_hybrid_parallel_model(model=skeleton, model_config=config, training_args=args,
                                           hybrid_parallel_configs=hybrid_parallel_configs)
