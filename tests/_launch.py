"""Starts the N rank processes of a multi-rank test.  GPU (``backend="cuda"``) jobs are plain subprocesses of the worker script.
CPU (gloo / oracle backend) jobs fork from a multiprocessing *forkserver* that has torch, torch._dynamo (imported by every
``torch.optim`` step: 3 s) and this repo's packages loaded once -- a rank process then starts in milliseconds instead of the ~6 s of
interpreter start-up + imports that used to dominate the several hundred short-lived ranks of the CPU suite.  ``HGB_TEST_SPAWN=subprocess``
(or any failure to bring the forkserver up) selects the plain subprocess path for CPU jobs too."""
import importlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PRELOAD = ["torch", "torch.distributed", "torch.optim", "torch._dynamo", "torch.utils.checkpoint", "hetu_galvatron_b200", "hetu_galvatron_b200.core",
            "hetu_galvatron_b200.llama_hf", "hetu_galvatron_b200.gpt_hf", "hetu_galvatron_b200.bert_hf", "oracle.llama_ref", "oracle.gpt_bert_ref",
            "oracle.gloo_backend", "smoke_model"]
_ctx = [None]


def _rank_entry(worker, env, out_path):
    """body of one forked rank: the worker module's main() with the rank's environment, output into ``out_path``"""
    os.environ.update(env)
    fd = os.open(out_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
    os.dup2(fd, 1)
    os.dup2(fd, 2)
    sys.stdout = os.fdopen(1, "w", buffering=1, closefd=False)
    sys.stderr = sys.stdout
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        importlib.import_module(worker).main()
    except BaseException:  # noqa: BLE001 -- the parent shows the tail of this output
        traceback.print_exc()
        sys.stdout.flush()
        os._exit(1)
    sys.stdout.flush()
    os._exit(0)          # (skip interpreter teardown: gloo's helper threads can outlive it)


def _forkserver():
    if _ctx[0] is None:
        ctx = mp.get_context("forkserver")
        ctx.set_forkserver_preload(_PRELOAD)
        _ctx[0] = ctx
    return _ctx[0]


def _run_forked(worker, envs, timeout):
    ctx = _forkserver()
    tmp = tempfile.mkdtemp(prefix="hgb_ranks_")
    paths = [os.path.join(tmp, "rank%d.log" % r) for r in range(len(envs))]
    procs = [ctx.Process(target=_rank_entry, args=(worker, env, path), daemon=True) for env, path in zip(envs, paths)]
    for p in procs:
        p.start()
    deadline = time.time() + timeout
    for p in procs:
        p.join(max(0.1, deadline - time.time()))
    timed_out = [p for p in procs if p.is_alive()]
    for p in timed_out:
        p.kill()                     # exactly the processes this call started
        p.join(5)
    outs, started = [], True
    for path in paths:
        try:
            with open(path, errors="replace") as f:
                outs.append(f.read())
        except OSError:          # the rank never reached _rank_entry (the forkserver could not prepare the child)
            outs.append("")
            started = False
        try:
            os.remove(path)
        except OSError:
            pass
    try:
        os.rmdir(tmp)
    except OSError:
        pass
    if not started and not timed_out:
        raise RuntimeError("forkserver children did not start")
    codes = [(-9 if p in timed_out else p.exitcode) for p in procs]
    return codes, outs


def _run_subprocess(worker, envs, timeout):
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", worker + ".py")], env=dict(os.environ, **env),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for env in envs]
    outs = [p.communicate(timeout=timeout)[0] for p in procs]
    return [p.returncode for p in procs], outs


def launch_ranks(worker, world, config, port, timeout=600, backend="oracle", extra_env=None):
    """-> the HOST_TEST_REPORT dict rank 0 printed; raises AssertionError with the failing rank's output otherwise"""
    envs = [dict(extra_env or {}, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                 HOST_TEST_CONFIG=json.dumps(config), OMP_NUM_THREADS="1" if world >= 4 else "2", HOST_TEST_BACKEND=backend) for rank in range(world)]
    codes = outs = None
    if backend == "oracle" and os.environ.get("HGB_TEST_SPAWN", "forkserver") != "subprocess":
        try:
            codes, outs = _run_forked(worker, envs, timeout)
        except (OSError, RuntimeError, ImportError, EOFError):       # no forkserver here: fall back
            codes = None
    if codes is None:
        codes, outs = _run_subprocess(worker, envs, timeout)
    for rank, (code, out) in enumerate(zip(codes, outs)):
        assert code == 0, "rank %d failed (exit %s):\n%s" % (rank, code, out[-4000:])
    line = [ln for ln in outs[0].splitlines() if ln.startswith("HOST_TEST_REPORT ")][-1]
    return json.loads(line[len("HOST_TEST_REPORT "):])
