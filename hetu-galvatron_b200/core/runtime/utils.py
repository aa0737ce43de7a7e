"""Optimizer / grad-norm helpers of ``galvatron/core/runtime/utils.py`` (:124-167).

The optimizer itself is out of the hot-path scope ("keep torch/apex"): apex ``FusedAdam`` (AdamW mode)
becomes ``torch.optim.AdamW(fused=True)`` over the fp32 flat shards the sharded units expose.
"""
import torch


class FusedShardedAdamW(torch.optim.Optimizer):
    """AdamW whose update runs inside each layer's gradient reduce-scatter kernel: when a layer's last
    backward of the step finishes, ONE kernel pulls the peers' gradient slices, sums them and applies the AdamW step to
    the fp32 shard -- the fp32 gradient buffer is never materialised (30 GiB less at Llama-3-8B on one GPU) and the
    optimizer's own pass over the state disappears.  ``step()`` only advances the step counter and fences the reduce
    stream; hyper-parameters take effect for the NEXT forward_backward.  Same rule as torch.optim.AdamW."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        params = list(model.parameters())
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.units = list(model.model.units)
        self.step_count = 0
        self._fallback = None
        plain = []
        for u in self.units:
            u.fused_opt = self
            if u.uses_fused_optimizer():
                u.exp_avg = torch.zeros_like(u.flat_param.data)
                u.exp_avg_sq = torch.zeros_like(u.flat_param.data)
            else:
                plain.append(u.flat_param)
        if plain:   # replicated DDP layers: ordinary AdamW on their fp32 gradients
            g = self.param_groups[0]
            self._fallback = torch.optim.AdamW(plain, lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"],
                                               fused=all(p.is_cuda for p in plain))

    def hyper(self):
        g = self.param_groups[0]
        return g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.step_count + 1

    @torch.no_grad()
    def step(self, closure=None):
        from .backend import get_backend
        get_backend().finish_reductions()
        if self._fallback is not None:
            for gsrc, gdst in zip(self.param_groups, self._fallback.param_groups):
                gdst["lr"] = gsrc["lr"]
            self._fallback.step()
        self.step_count += 1

    def zero_grad(self, set_to_none=True):
        if self._fallback is not None:
            self._fallback.zero_grad(set_to_none)

    def state_dict(self):
        """Per-rank optimizer state (what ``optimizer/<rank>.pt`` of a distributed checkpoint holds): the moments live on the
        sharded units, next to the fp32 shard the reduce-scatter epilogue updates."""
        return {"fused_sharded_adamw": 1, "step_count": self.step_count,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "units": {u.name: {"exp_avg": u.exp_avg.detach().cpu(), "exp_avg_sq": u.exp_avg_sq.detach().cpu()}
                          for u in self.units if u.uses_fused_optimizer()},
                "fallback": None if self._fallback is None else self._fallback.state_dict()}

    def load_state_dict(self, state):
        if not state.get("fused_sharded_adamw"):
            raise ValueError("not a FusedShardedAdamW state (saved with --optimizer torch?)")
        self.step_count = int(state["step_count"])
        for g, saved in zip(self.param_groups, state["param_groups"]):
            g.update(saved)
        for u in self.units:
            if u.uses_fused_optimizer():
                rec = state["units"][u.name]
                u.exp_avg.copy_(rec["exp_avg"])
                u.exp_avg_sq.copy_(rec["exp_avg_sq"])
        if self._fallback is not None and state.get("fallback") is not None:
            self._fallback.load_state_dict(state["fallback"])


def _constant_lr(step):
    return 1.0


def get_optimizer_and_param_scheduler(model, args):
    """``core/runtime/utils.py:140-167``: AdamW over the fp32 flat shards + the LR scheduler; with ``--distributed_checkpoint``
    both resume from ``<load>/iter_<load_iteration>/{optimizer/<rank>.pt, opt_param_scheduler.json}`` (:152-165)."""
    if getattr(args, "fused_optimizer", False):
        optimizer = FusedShardedAdamW(model, lr=args.lr, betas=(getattr(args, "adam_beta1", 0.9), getattr(args, "adam_beta2", 0.999)),
                                      eps=getattr(args, "adam_eps", 1e-8), weight_decay=args.adam_weight_decay)
    else:
        params = list(model.parameters())
        optimizer = torch.optim.AdamW(params, lr=args.lr, weight_decay=args.adam_weight_decay,
                                      betas=(getattr(args, "adam_beta1", 0.9), getattr(args, "adam_beta2", 0.999)),
                                      eps=getattr(args, "adam_eps", 1e-8), fused=all(p.is_cuda for p in params))
    scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, _constant_lr)   # constant LR (random-data scripts)
    if getattr(args, "distributed_checkpoint", False) and getattr(args, "load", None):
        import json
        import os
        from .backend import get_backend
        root = os.path.join(args.load, "iter_%d" % int(getattr(args, "load_iteration", 0)))
        opt_file = os.path.join(root, "optimizer", "%d.pt" % get_backend().rank)
        if os.path.exists(opt_file):
            optimizer.load_state_dict(torch.load(opt_file, map_location="cpu", weights_only=True))
        sched_file = os.path.join(root, "opt_param_scheduler.json")
        if os.path.exists(sched_file):
            saved = json.load(open(sched_file))
            if saved:
                scheduler.load_state_dict(saved)
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
