"""The BERT family harness (``galvatron/models/bert_hf``): the three callbacks + ModelInfo the core API asks for."""
from .BertModel_hybrid_parallel import bert_model_hp, construct_hybrid_parallel_model, get_bert_config, get_hybrid_parallel_configs
from .BertModel_sequential import BertModelInfo, construct_sequential_model
from .BertModel_tensor_parallel import BertLayer_tp, construct_tensor_parallel_model
from .meta_configs import config_from_meta, set_model_config
from .BertModel_checkpoint import load_bert_module  # noqa: E402,F401
