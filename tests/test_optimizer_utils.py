"""Optimizer-side helpers of galvatron/core/runtime/utils.py:124-167 that the training scripts call every step:
  * the LR / weight-decay scheduler -- pinned to the reference's own class (tests/golden/opt_param_scheduler.json, produced by
    oracle/gen_golden_sched.py from the unmodified megatron/training/optimizer_param_scheduler.py): same values, same state_dict;
  * clip_grad_norm over >1 rank -- the job-wide norm must equal the un-parallelised oracle's (every parameter counted once:
    shards, replicas, tensor-parallel duplicates), every rank must enter the collective, all shards scaled alike."""
import json
import os
import sys
import types

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_host_runtime import launch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "opt_param_scheduler.json")))


class _Opt:
    def __init__(self):
        self.param_groups = [{"lr": 0.0, "weight_decay": 0.0}, {"lr": 0.0, "weight_decay": 0.0, "lr_mult": 0.5, "wd_mult": 0.0}]


def _row(opt):
    return [g["lr"] for g in opt.param_groups] + [g["weight_decay"] for g in opt.param_groups]


@pytest.mark.parametrize("case", range(len(GOLD["cases"])))
def test_scheduler_matches_reference_class(case):
    from hetu_galvatron_b200.core.runtime.utils import OptimizerParamScheduler
    rec = GOLD["cases"][case]
    opt = _Opt()
    s = OptimizerParamScheduler(opt, **rec["kwargs"])
    trace = [_row(opt)]
    for _ in range(len(rec["trace"]) - 1):
        s.step(rec["increment"])
        trace.append(_row(opt))
    for got, want in zip(trace, rec["trace"]):
        assert got == pytest.approx(want, rel=1e-12, abs=0.0)
    assert s.state_dict() == rec["state_dict"]
    # a reference-written opt_param_scheduler.json resumes here exactly as it does there
    opt2 = _Opt()
    s2 = OptimizerParamScheduler(opt2, **rec["kwargs"])
    s2.load_state_dict(rec["state_dict"])
    s2.step(rec["increment"])
    assert _row(opt2) == pytest.approx(rec["after_resume"], rel=1e-12, abs=0.0)


def test_scheduler_rejects_foreign_state_and_defaults_to_constant():
    from hetu_galvatron_b200.core.runtime.utils import get_optimizer_param_scheduler
    args = types.SimpleNamespace(global_train_batch_size=8, lr=1e-4, adam_weight_decay=0.01)
    opt = _Opt()
    s = get_optimizer_param_scheduler(opt, args)
    for _ in range(5):
        s.step(8)
    assert opt.param_groups[0]["lr"] == 1e-4 and opt.param_groups[0]["weight_decay"] == 0.01
    with pytest.raises(ValueError):
        s.load_state_dict({"lr_lambdas": [None], "_last_lr": [1e-4], "last_epoch": 3})
    args = types.SimpleNamespace(global_train_batch_size=8, lr=1e-4, adam_weight_decay=0.01, train_iters=10, lr_warmup_iters=2)
    s = get_optimizer_param_scheduler(_Opt(), args)
    assert (s.lr_decay_style, s.lr_decay_steps, s.lr_warmup_steps) == ("linear", 80, 16)     # training.py:439-447


CLIP = {
    # norm weights are replicated over the tensor-parallel group (and summed over it under Megatron-SP): counted once
    "tp2_megatron_sp": (2, dict(global_tp_deg=2, vocab_tp=2, sequence_parallel=True)),
    "dp2_ddp": (2, dict(default_dp_type="ddp")),
    "dp2_zero3": (2, dict(sdp=1, embed_sdp=1)),
    "tp2_dp2_zero2": (4, dict(global_tp_deg=2, vocab_tp=2, default_dp_type="zero2", chunks=2)),
    "pp2": (2, dict(pp_deg=2, chunks=2)),
}


@pytest.mark.parametrize("name", sorted(CLIP))
def test_clip_grad_norm_multi_rank(name):
    world, over = CLIP[name]
    rep = launch(world, dict(over, _clip_grad=0.05))
    assert abs(rep["clip_norm"] - rep["clip_norm_oracle"]) <= 2e-2 * rep["clip_norm_oracle"], rep
    assert rep["clip_expected_ratio"] < 1.0
    for r in rep["clip_ratios"]:
        assert abs(r - rep["clip_expected_ratio"]) <= 1e-3 * rep["clip_expected_ratio"], rep


def test_clip_with_fused_optimizer_is_rejected():
    with pytest.raises(AssertionError) as err:
        launch(1, dict(fused_optimizer=True, _clip_grad=1.0, _tol=float("inf")))
    assert "fused_optimizer" in str(err.value)
