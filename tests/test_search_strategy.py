"""The strategies bench.py loads ARE the reference Search Engine's output: re-run the unmodified engine
(galvatron/core/search_engine + csrc/dp_core.cpp, via scripts/search_strategy.py) on the committed B200 profiles and compare with
configs/galvatron_config_llama3-8b_<N>gpus.json.  Needs /root/reference (build container only; skipped on the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("pp_deg", "tp_sizes_enc", "tp_consecutive_flags", "dp_types_enc", "use_sp", "checkpoint", "global_bsz", "chunks", "pp_division",
        "pipeline_type", "default_dp_type", "vtp", "vsp", "embed_sdp")


@pytest.mark.skipif(not os.path.isdir("/root/reference/galvatron/core/search_engine"), reason="needs the reference sources")
@pytest.mark.parametrize("n", [1] + ([2, 4, 8] if os.environ.get("HGB_SLOW_TESTS") else []))   # the 8-GPU search takes minutes
def test_bench_strategy_is_the_search_engines_output(n, tmp_path):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "search_strategy.py"), "--memory-gb", "178", "--gpus", str(n),
                          "--out-root", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    got = json.load(open(tmp_path / "searched" / ("galvatron_config_llama3-8b_%dgpus.json" % n)))
    want = json.load(open(os.path.join(ROOT, "configs", "galvatron_config_llama3-8b_%dgpus.json" % n)))
    for k in KEYS:
        if k in want or k in got:
            assert got.get(k) == want.get(k), (k, got.get(k), want.get(k))
    if n == 1:
        assert got["checkpoint"].split(",").count("1") == 15      # the engine's 20 % allocator reserve forces 15 of 32 layers
