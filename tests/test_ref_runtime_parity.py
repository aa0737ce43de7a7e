"""fp parity pinned to the REFERENCE RUNTIME, not only to HF: tests/golden/ref_runtime/<case>.json holds the per-step losses the
UNMODIFIED reference (PKU-DAIR/Hetu-Galvatron runtime: FSDP + Megatron layers + flash-attn, run on B200s by
oracle/ref_runtime/run_ref.py) computes for 3 Adam steps of the tiny Llama of tests/golden/ckpt_llama_tiny under a parallel
strategy.  This repo's runtime, given the same weights (the same converted checkpoint), the same token stream and the same
optimizer, must reproduce them: step 0 (pure forward) within 2e-3 rel, the later steps -- which fold in every gradient through the
optimizer -- within 5e-3 (the reference's own criterion against HF, tests/core/test_tp.py:121; both sides compute in bf16).
CPU: the host runtime on the oracle backend.  GPU (``-m gpu``): the product path through the C ABI."""
import glob
import json
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_host_runtime import launch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ckpt_llama_tiny")
FIXTURES = {os.path.basename(p)[:-5]: json.load(open(p)) for p in sorted(glob.glob(os.path.join(HERE, "golden", "ref_runtime", "*.json")))}


def _ours(fx, backend):
    over = {k: (bool(v) if k in ("sequence_parallel", "use_ulysses") else v) for k, v in fx["overrides"].items()}
    cfg = dict(over, load=GOLDEN, adam_weight_decay=0.0, lr=1e-3, _iters=fx["steps"], _tol=float("inf") if backend == "cuda" else 3e-2)
    return launch(fx["world"], cfg, backend=backend)


def _check(fx, rep):
    ref, got = fx["losses"], rep["losses"]
    assert len(got) == len(ref)
    assert abs(got[0] - ref[0]) <= 2e-3 * abs(ref[0]), (ref, got)
    for a, b in zip(got[1:], ref[1:]):
        assert abs(a - b) <= 5e-3 * abs(b), (ref, got)
    assert ref[-1] < ref[0]            # the reference run itself trains


@pytest.mark.skipif(not FIXTURES, reason="no reference-runtime fixtures committed")
@pytest.mark.parametrize("case", sorted(FIXTURES) or ["none"])
def test_host_runtime_matches_reference_runtime(case):
    fx = FIXTURES[case]
    if fx["world"] > 4:
        pytest.skip("large world")
    _check(fx, _ours(fx, "oracle"))


@pytest.mark.gpu
@pytest.mark.skipif(not FIXTURES, reason="no reference-runtime fixtures committed")
@pytest.mark.parametrize("case", sorted(FIXTURES) or ["none"])
def test_product_path_matches_reference_runtime(case):
    fx = FIXTURES[case]
    if not torch.cuda.is_available() or torch.cuda.device_count() < fx["world"]:
        pytest.skip("needs %d GPU(s)" % fx["world"])
    _check(fx, _ours(fx, "cuda"))
