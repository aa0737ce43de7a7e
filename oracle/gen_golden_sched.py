"""Golden values of the LR / weight-decay schedule, produced by the UNMODIFIED reference class
(galvatron/site_package/megatron/training/optimizer_param_scheduler.py:9-229) run here: the file is loaded on its own, with
a stand-in for its one relative import (``.utils.print_rank_0``).  Writes tests/golden/opt_param_scheduler.json.

    python oracle/gen_golden_sched.py
"""
import importlib.util
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/galvatron/site_package/megatron/training/optimizer_param_scheduler.py"


def load_reference():
    pkg = types.ModuleType("refsched")
    pkg.__path__ = []
    utils = types.ModuleType("refsched.utils")
    utils.print_rank_0 = lambda *a, **k: None
    sys.modules["refsched"], sys.modules["refsched.utils"] = pkg, utils
    spec = importlib.util.spec_from_file_location("refsched.optimizer_param_scheduler", SRC)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "refsched"
    spec.loader.exec_module(mod)
    return mod.OptimizerParamScheduler


class FakeOptimizer:
    def __init__(self):
        self.param_groups = [{"lr": 0.0, "weight_decay": 0.0}, {"lr": 0.0, "weight_decay": 0.0, "lr_mult": 0.5, "wd_mult": 0.0}]


CASES = [
    dict(init_lr=0.0, max_lr=1e-4, min_lr=0.0, lr_warmup_steps=0, lr_decay_steps=640, lr_decay_style="linear",
         start_wd=0.01, end_wd=0.01, wd_incr_steps=640, wd_incr_style="constant"),
    dict(init_lr=1e-6, max_lr=3e-4, min_lr=3e-5, lr_warmup_steps=64, lr_decay_steps=1024, lr_decay_style="cosine",
         start_wd=0.0, end_wd=0.1, wd_incr_steps=1024, wd_incr_style="linear"),
    dict(init_lr=0.0, max_lr=2e-4, min_lr=1e-5, lr_warmup_steps=32, lr_decay_steps=512, lr_decay_style="inverse-square-root",
         start_wd=0.01, end_wd=0.05, wd_incr_steps=256, wd_incr_style="cosine"),
    dict(init_lr=0.0, max_lr=1e-4, min_lr=0.0, lr_warmup_steps=16, lr_decay_steps=128, lr_decay_style="constant",
         start_wd=0.01, end_wd=0.01, wd_incr_steps=128, wd_incr_style="constant"),
]


def main():
    cls = load_reference()
    out = []
    for kw in CASES:
        opt = FakeOptimizer()
        s = cls(opt, use_checkpoint_opt_param_scheduler=True, override_opt_param_scheduler=False, **kw)
        trace = [[g["lr"] for g in opt.param_groups] + [g["weight_decay"] for g in opt.param_groups]]
        for _ in range(48):
            s.step(32)
            trace.append([g["lr"] for g in opt.param_groups] + [g["weight_decay"] for g in opt.param_groups])
        sd = s.state_dict()
        # resume: a fresh scheduler that loads the state continues identically
        opt2 = FakeOptimizer()
        s2 = cls(opt2, use_checkpoint_opt_param_scheduler=True, override_opt_param_scheduler=False, **kw)
        s2.load_state_dict(sd)
        s.step(32); s2.step(32)
        assert [g["lr"] for g in opt.param_groups] == [g["lr"] for g in opt2.param_groups]
        out.append({"kwargs": kw, "increment": 32, "trace": trace, "state_dict": sd,
                    "after_resume": [g["lr"] for g in opt2.param_groups] + [g["weight_decay"] for g in opt2.param_groups]})
    path = os.path.join(ROOT, "tests", "golden", "opt_param_scheduler.json")
    with open(path, "w") as f:
        json.dump({"source": SRC, "cases": out}, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
