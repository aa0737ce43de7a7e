"""The descriptor channel that carries VMM arena handles and multicast objects between the processes of a node
(hetu-galvatron_b200/_bg.py _FdChannel; cudaIpc handles cannot carry either): 3 processes, every rank passes every peer a real file
descriptor (a temp file it wrote) plus a tagged one from rank 0 only, and every receiver reads the sender's bytes through it."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, tempfile
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from hetu_galvatron_b200._bg import _FdChannel
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ch = _FdChannel(rank, world)
    f = tempfile.TemporaryFile()
    f.write(b"arena-of-%%d" %% rank); f.flush()
    out = [(peer, "arena", f.fileno()) for peer in range(world) if peer != rank]
    got = ch.exchange(out, world - 1)
    assert sorted(src for src, _ in got) == [r for r in range(world) if r != rank]
    for (src, tag), fd in got.items():
        assert tag == "arena"
        assert os.pread(fd, 64, 0) == b"arena-of-%%d" %% src      # a NEW descriptor onto the sender's open file
        os.close(fd)
    # second round on the same channel: only rank 0 sends (a group leader handing out a multicast object)
    g = tempfile.TemporaryFile(); g.write(b"mc"); g.flush()
    out = [(peer, ("mc", (0, 1, 2)), g.fileno()) for peer in range(1, world)] if rank == 0 else []
    got = ch.exchange(out, 0 if rank == 0 else 1)
    if rank:
        assert list(got) == [(0, ("mc", (0, 1, 2)))] and os.pread(got[(0, ("mc", (0, 1, 2)))], 8, 0) == b"mc"
    dist.barrier()
    ch.close()
    assert not os.path.exists(ch.path)
    print("FD_OK", rank)
""") % ROOT


def test_descriptors_cross_processes(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(3):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="3", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29450 + os.getpid() % 400))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "FD_OK %d" % rank in out, out[-2000:]
