"""Optimizer / grad-norm helpers of ``galvatron/core/runtime/utils.py`` (:124-167).

The optimizer itself is out of the hot-path scope ("keep torch/apex"): apex ``FusedAdam`` (AdamW mode)
becomes ``torch.optim.AdamW(fused=True)`` over the fp32 flat shards the sharded units expose.
"""
import torch


def get_optimizer_and_param_scheduler(model, args):
    params = list(model.parameters())
    fused = all(p.is_cuda for p in params)
    optimizer = torch.optim.AdamW(params, lr=args.lr, weight_decay=args.adam_weight_decay,
                                  betas=(getattr(args, "adam_beta1", 0.9), getattr(args, "adam_beta2", 0.999)),
                                  eps=getattr(args, "adam_eps", 1e-8), fused=fused)
    scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda step: 1.0)   # constant LR (random-data scripts)
    return optimizer, scheduler


def clip_grad_norm(model, max_norm, norm_type=2):
    """Global L2 norm over the job, then scale (utils.py:124-133 -> megatron clip_grad_norm_fp32).  Every flat shard is
    counted once: a unit's shards partition its parameters over the SDP group (DDP units are replicated -> divided by the
    group size); tensor-parallel shards hold different parameters."""
    from .backend import get_backend
    assert norm_type == 2
    be = get_backend()
    grads, total = [], None
    for u in model.model.units:
        g = u.flat_param.grad
        if g is None:
            continue
        grads.append(g)
        sq = g.float().pow(2).sum()
        if u.dp_type == "ddp":
            sq = sq / u.group.size
        total = sq if total is None else total + sq
    if total is None:
        return 0.0
    buf = torch.zeros(4, dtype=torch.float32, device=total.device)
    buf[0] = total
    if be.world > 1:
        from .comm_groups import CommGroup
        world_group = getattr(be, "_world_group", None)
        if world_group is None:
            raise RuntimeError("clip_grad_norm over world > 1 needs backend.reserve_world_group() before exchange()")
        buf = be.all_reduce(buf, world_group)
    norm = buf[0].sqrt()
    coef = (max_norm / (norm + 1e-6)).clamp(max=1.0)
    for g in grads:
        g.mul_(coef)
    return float(norm)
