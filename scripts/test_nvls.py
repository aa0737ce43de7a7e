"""Multi-GPU check of the opt-in VMM arena + NVLS all-reduce (torchrun, one process per GPU):
  1. the arena comes from cuMemCreate, is shared as file descriptors, and the ordinary peer kernels run on it unchanged
  2. a multicast object is bound over a symmetric buffer and bg_all_reduce_nvls matches NCCL's all-reduce (bf16 and fp32)
  3. timing against our peer-to-peer all-reduce and NCCL, 1 MiB .. 1 GiB
JSON lines on rank 0; NVLS_OK / NVLS_UNSUPPORTED / NVLS_FAIL on the last line.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/test_nvls.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hetu_galvatron_b200 import _bg  # noqa: E402
from hetu_galvatron_b200.core.runtime.comm_groups import CommGroup  # noqa: E402


def timed(fn, iters=10, warm=3):
    ts = []
    for i in range(warm + iters):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    max_bytes = int(os.environ.get("NVLS_MAX_MB", "1024")) << 20
    _bg.set_tunable("timeout_ms", 20000)
    comm = _bg.BgComm(rank, world, local, max_bytes * 2 + (256 << 20), vmm=True)
    vmm, mc, gran = comm.arena_mode()
    if rank == 0:
        print(json.dumps({"arena": "vmm" if vmm else "cudaMalloc", "multicast_supported": mc, "multicast_granularity": gran}), flush=True)
    comm.connect_vmm()
    grp = CommGroup(list(range(world)))
    align = gran if mc else 256
    buf = comm.sym_alloc(grp, max_bytes, align=max(256, align))
    comm.exchange()
    ok = True
    # 1. the peer-to-peer kernels on a VMM arena
    n = 1 << 20
    src = buf.view(torch.bfloat16, n)
    src.copy_(torch.full((n,), float(rank + 1), device="cuda", dtype=torch.bfloat16))
    out = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize(); dist.barrier()
    comm.all_reduce(grp, buf, out, elems=n)
    torch.cuda.synchronize()
    ok &= bool((out.float() == world * (world + 1) / 2).all())
    if rank == 0:
        print(json.dumps({"check": "peer all-reduce on the VMM arena", "ok": ok}), flush=True)
    if not mc:
        dist.barrier()
        if rank == 0:
            print("NVLS_UNSUPPORTED" if ok else "NVLS_FAIL", flush=True)
        comm.close(); dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    # 2. + 3. NVLS
    comm.setup_nvls([buf])
    for dtype in (torch.bfloat16, torch.float32):
        esz = 2 if dtype == torch.bfloat16 else 4
        for nbytes in [1 << 20, 16 << 20, 64 << 20, 256 << 20, max_bytes]:
            if nbytes > max_bytes:
                continue
            elems = nbytes // esz
            torch.manual_seed(100 + rank)
            x = torch.randn(elems, device="cuda").to(dtype)
            view = buf.view(dtype, elems)

            def fill():
                view.copy_(x)
            fill()
            out = torch.empty(elems, device="cuda", dtype=dtype)
            torch.cuda.synchronize(); dist.barrier()
            comm.all_reduce_nvls(grp, 0, out, elems, dtype)
            ref = x.clone(); dist.all_reduce(ref)
            torch.cuda.synchronize()
            err = float((out.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-6))
            good = err < (2e-2 if dtype == torch.bfloat16 else 1e-5)
            ok &= good
            fill(); torch.cuda.synchronize(); dist.barrier()
            nvls_ms = timed(lambda: comm.all_reduce_nvls(grp, 0, out, elems, dtype))
            fill(); torch.cuda.synchronize(); dist.barrier()
            inplace_ms = timed(lambda: comm.all_reduce_nvls(grp, 0, None, elems, dtype))   # result left in the buffer (no copy-out)
            fill(); torch.cuda.synchronize(); dist.barrier()
            p2p_ms = timed(lambda: comm.all_reduce(grp, buf, out, elems=elems))
            nccl_ms = timed(lambda: dist.all_reduce(ref))
            bus = lambda ms: round(nbytes * 2 * (world - 1) / world / ms / 1e6, 1)  # noqa: E731
            if rank == 0:
                print(json.dumps({"op": "all_reduce", "dtype": str(dtype), "bytes": nbytes, "p": world, "nvls_ms": round(nvls_ms, 4),
                                  "nvls_busGBps": bus(nvls_ms), "nvls_inplace_ms": round(inplace_ms, 4), "nvls_inplace_busGBps": bus(inplace_ms), "ours_p2p_ms": round(p2p_ms, 4), "ours_p2p_busGBps": bus(p2p_ms),
                                  "nccl_ms": round(nccl_ms, 4), "nccl_busGBps": bus(nccl_ms), "max_rel_err_vs_nccl": round(err, 6),
                                  "ok": good}), flush=True)
    assert comm.error_flag() == 0
    dist.barrier()
    if rank == 0:
        print("NVLS_OK" if ok else "NVLS_FAIL", flush=True)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
