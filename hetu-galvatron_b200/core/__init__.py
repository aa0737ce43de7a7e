"""Public surface mirroring ``galvatron.core`` for the hot path (``galvatron/core/__init__.py:1-17``)."""
from .runtime.comm_groups import CommGroup, gen_comm_groups
from .runtime.hybrid_parallel_config import (ModelInfo, check_hp_config, get_chunks, get_hybrid_parallel_configs_api,
                                             hp_config_whole_model, layer_shapes_dtypes_whole_model,
                                             mixed_precision_dtype)


def __getattr__(name):
    # heavier pieces (need torch + the CUDA extension) are resolved on first use
    if name in ("construct_hybrid_parallel_model_api", "GalvatronModel"):
        from .runtime import hybrid_parallel_model as _m
        return getattr(_m, name)
    if name in ("initialize_galvatron", "get_args", "set_args"):
        from .runtime import arguments as _a
        return getattr(_a, name)
    if name in ("clip_grad_norm", "get_optimizer_and_param_scheduler"):
        from .runtime import utils as _u
        return getattr(_u, name)
    raise AttributeError(name)
