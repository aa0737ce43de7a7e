"""Real multi-GPU micro-benchmark + parity check of the peer-memory collectives against NCCL (torchrun, one process per GPU).

    torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_collectives.py [--max-mb 1024]

For each message size: our kernel vs the torch.distributed (NCCL) call the reference issues at the same call site, both
CUDA-event timed (max over ranks), bus bandwidth by the nccl-tests convention (AG/RS/A2A (p-1)/p*N, AR 2(p-1)/p*N), and a
result comparison (bit-exact for data movement, bf16 tolerance for reductions).  JSON lines on rank 0."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hetu_galvatron_b200 import _bg as bg  # noqa: E402
from hetu_galvatron_b200.core.runtime.comm_groups import CommGroup  # noqa: E402

BF = torch.bfloat16


def timed(fn, iters=20, warm=5):
    """Per-launch CUDA events (the kernels wait for their peers, so a launch's span includes rank skew): report the
    MEDIAN launch, max over ranks.  A back-to-back loop time is dominated by host launch skew for sub-ms messages."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ts[len(ts) // 2]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-mb", type=int, default=1024)
    ap.add_argument("--ctas", type=int, default=0)
    opts = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    max_bytes = opts.max_mb << 20
    comm = bg.BgComm(rank, world, local, 3 * max_bytes + (64 << 20))
    comm.connect_ipc()
    if opts.ctas:
        bg.set_tunable("comm_ctas", opts.ctas)
    grp = CommGroup(list(range(world)))
    full = comm.sym_alloc(grp, max_bytes)      # unsharded buffer (AG dst / RS src / AR src / A2A src)
    comm.exchange()
    p = world
    out = []

    def emit(rec):
        if rank == 0:
            print(json.dumps(rec), flush=True)

    sizes = [1 << 20]
    while sizes[-1] * 4 <= max_bytes:
        sizes.append(sizes[-1] * 4)
    if sizes[-1] != max_bytes:
        sizes.append(max_bytes)
    torch.manual_seed(1234 + rank)
    for nbytes in sizes:
        n = nbytes // 2                     # bf16 elements of the full buffer
        shard = n // p
        # ---- all-gather (+cast): fp32 shard -> bf16 full ------------------------------------------------------
        master = torch.randn(shard, device="cuda")
        ours_ms = timed(lambda: comm.all_gather_cast(grp, master, full, shard_elems=shard, lane=0))
        got = full.view(BF, n).clone()
        src16 = master.to(BF)
        ref = torch.empty(n, device="cuda", dtype=BF)
        nccl_ms = timed(lambda: dist.all_gather_into_tensor(ref, master.to(BF)))
        ok = bool(torch.equal(got.view(torch.int16), ref.view(torch.int16)))
        emit({"op": "all_gather_cast", "bytes": nbytes, "p": p, "ms": round(ours_ms, 4), "busGBps": round(nbytes * (p - 1) / p / ours_ms / 1e6, 1),
              "nccl_ms(cast+ag)": round(nccl_ms, 4), "nccl_busGBps": round(nbytes * (p - 1) / p / nccl_ms / 1e6, 1), "bit_exact_vs_nccl": ok})
        # ---- reduce-scatter (+scale, cast, accumulate): bf16 full -> fp32 shard ---------------------------------
        grad = torch.randn(n, device="cuda").to(BF)
        full.view(BF, n).copy_(grad)
        acc = torch.zeros(shard, device="cuda")
        torch.cuda.synchronize(); dist.barrier()
        ours_ms = timed(lambda: comm.reduce_scatter_acc(grp, full, BF, acc, shard_elems=shard, prescale=0.5, postscale=1.0 / p * 2, accumulate=False, lane=1))
        ref_sh = torch.empty(shard, device="cuda", dtype=BF)

        def nccl_rs():
            g = grad / 2
            dist.reduce_scatter_tensor(ref_sh, g)
            return ref_sh.float() * (2.0 / p)
        nccl_ms = timed(nccl_rs)
        want = nccl_rs()
        err = float((acc - want).abs().max() / (want.abs().max() + 1e-6))
        emit({"op": "reduce_scatter_acc", "bytes": nbytes, "p": p, "ms": round(ours_ms, 4), "busGBps": round(nbytes * (p - 1) / p / ours_ms / 1e6, 1),
              "nccl_ms(div+rs+cast)": round(nccl_ms, 4), "nccl_busGBps": round(nbytes * (p - 1) / p / nccl_ms / 1e6, 1), "max_rel_err_vs_nccl": round(err, 5)})
        # ---- all-reduce ------------------------------------------------------------------------------------------
        full.view(BF, n).copy_(grad)
        dst = torch.empty(n, device="cuda", dtype=BF)
        torch.cuda.synchronize(); dist.barrier()
        ours_ms = timed(lambda: (full.view(BF, n).copy_(grad), comm.all_reduce(grp, full, dst, elems=n, lane=2)))
        copy_ms = timed(lambda: full.view(BF, n).copy_(grad))
        ref_ar = grad.clone()
        nccl_ms = timed(lambda: dist.all_reduce(ref_ar.copy_(grad)))
        nccl_copy = timed(lambda: ref_ar.copy_(grad))
        ref_ar.copy_(grad); dist.all_reduce(ref_ar)
        full.view(BF, n).copy_(grad); comm.all_reduce(grp, full, dst, elems=n, lane=2); torch.cuda.synchronize()
        err = float((dst.float() - ref_ar.float()).abs().max() / (ref_ar.float().abs().max() + 1e-6))
        o, r = ours_ms - copy_ms, nccl_ms - nccl_copy
        emit({"op": "all_reduce", "bytes": nbytes, "p": p, "ms": round(o, 4), "busGBps": round(nbytes * 2 * (p - 1) / p / o / 1e6, 1),
              "nccl_ms": round(r, 4), "nccl_busGBps": round(nbytes * 2 * (p - 1) / p / r / 1e6, 1), "max_rel_err_vs_nccl": round(err, 5)})
        # ---- Ulysses all-to-all with fused transpose: [b, s/p, heads, d] -> [b, s, heads/p, d] ----------------------
        d, heads = 128, 32
        if heads % p == 0:
            b = 1
            s_loc = n // (heads * d * b)
            if s_loc >= 1:
                x = torch.randn(b, s_loc, heads, d, device="cuda").to(BF)
                full.view(BF, x.numel()).copy_(x.flatten())
                hp = heads // p
                y = torch.empty(b, s_loc * p, hp, d, device="cuda", dtype=BF)
                desc = [dict(src=full, dst=y, batch=b, rows=s_loc, row_elems=hp * d, src_bs=s_loc * heads * d, src_rs=heads * d, src_me_off=hp * d,
                             dst_bs=s_loc * p * hp * d, dst_rs=hp * d, dst_peer_off=s_loc * hp * d)]
                torch.cuda.synchronize(); dist.barrier()
                ours_ms = timed(lambda: comm.all_to_all_rows(grp, desc, BF, lane=3))

                def nccl_a2a():   # transformer.py:1928-1987: permute copy, all_to_all_single, permute copy
                    t = x.reshape(b, s_loc, p, hp, d).permute(2, 0, 1, 3, 4).contiguous()
                    o_ = torch.empty_like(t)
                    dist.all_to_all_single(o_, t)
                    return o_.permute(1, 0, 2, 3, 4).contiguous().reshape(b, p * s_loc, hp, d)
                nccl_ms = timed(nccl_a2a)
                ok = bool(torch.equal(y.view(torch.int16), nccl_a2a().view(torch.int16)))
                nb = x.numel() * 2
                emit({"op": "ulysses_all_to_all", "bytes": nb, "p": p, "ms": round(ours_ms, 4), "busGBps": round(nb * (p - 1) / p / ours_ms / 1e6, 1),
                      "nccl_ms(permute+a2a+permute)": round(nccl_ms, 4), "nccl_busGBps": round(nb * (p - 1) / p / nccl_ms / 1e6, 1), "bit_exact_vs_nccl": ok})
    assert comm.error_flag() == 0
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
