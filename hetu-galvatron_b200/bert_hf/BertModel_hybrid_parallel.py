"""Entry points of the BERT family (``galvatron/models/bert_hf/BertModel_hybrid_parallel.py``)."""
from ..core.runtime.hybrid_parallel_config import get_hybrid_parallel_configs_api
from ..core.runtime.hybrid_parallel_model import construct_hybrid_parallel_model_api
from ..llama_hf.LlamaModel_hybrid_parallel import estimate_arena_bytes
from .BertModel_checkpoint import load_bert_module
from .BertModel_sequential import BertEmbeddings_, BertMLMCls_, BertModelInfo, construct_sequential_model
from .BertModel_tensor_parallel import BertLayer_tp, BertSkeleton, construct_tensor_parallel_model
from .meta_configs import config_from_meta, set_model_config


def get_hybrid_parallel_configs(model_config, training_args):
    return get_hybrid_parallel_configs_api(model_config, training_args, BertModelInfo)


def construct_hybrid_parallel_model(model, model_config, training_args, hybrid_parallel_configs):
    return construct_hybrid_parallel_model_api(
        model, model_config, training_args, hybrid_parallel_configs, BertModelInfo, construct_sequential_model,
        construct_tensor_parallel_model, wrap_block_name=[BertLayer_tp], wrap_checkpoint_block_name=[BertLayer_tp],
        wrap_other_block_name=[BertEmbeddings_, BertMLMCls_], tied_wte_attr_names=None, layernorm_name=["LayerNorm"],
        all_block_name=[BertEmbeddings_, BertLayer_tp, BertMLMCls_], load_module_func=load_bert_module)


def get_bert_config(args, overwrite_args=True):
    return set_model_config(config_from_meta(args.model_size), args, overwrite_args)


def bert_model_hp(config, args):
    hybrid_parallel_configs = get_hybrid_parallel_configs(model_config=config, training_args=args)
    if not getattr(args, "arena_bytes", 0):
        args.arena_bytes = estimate_arena_bytes(config, args, hybrid_parallel_configs)
    return construct_hybrid_parallel_model(model=BertSkeleton(config), model_config=config, training_args=args,
                                           hybrid_parallel_configs=hybrid_parallel_configs)
