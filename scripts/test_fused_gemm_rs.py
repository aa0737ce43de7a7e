"""Multi-GPU parity + timing of the fused GEMM+reduce-scatter against (our GEMM -> NCCL reduce_scatter) and
(cuBLAS -> NCCL reduce_scatter), torchrun one process per GPU.  JSON lines on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hetu_galvatron_b200.core.runtime.arguments import initialize_galvatron  # noqa: E402
from hetu_galvatron_b200.core.runtime.backend import get_backend  # noqa: E402
from hetu_galvatron_b200.core.runtime.comm_groups import CommGroup  # noqa: E402

BF = torch.bfloat16


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    initialize_galvatron(arena_bytes=1 << 30)
    be = get_backend()
    be.bg.set_tunable("timeout_ms", 20000)
    grp = CommGroup(list(range(world)))
    M, N = 8192, 4096
    be.reserve_staging(grp, M * N * 2 * 2)
    be.exchange()
    ok_all = True
    # library handles (cuBLAS workspace, NCCL channels) are created BEFORE any peer-waiting kernel is in flight
    warm = torch.randn(256, 256, device="cuda").to(BF)
    torch.matmul(warm, warm)
    dist.all_reduce(torch.zeros(8, device="cuda"))
    torch.cuda.synchronize()
    for layout, K in (("tn", 14336 // world), ("nn", 6144 // world), ("tn", 4096 // world)):
        torch.manual_seed(7 + rank)
        a = (torch.randn(M, K, device="cuda") * 0.5).to(BF)
        b = (torch.randn((N, K) if layout == "tn" else (K, N), device="cuda") * 0.5).to(BF)
        out = be.gemm_reduce_scatter(a, b, layout, grp)
        if os.environ.get("HGB_DIAG"):   # snapshot of the arrival counters while the kernels may still be spinning
            import time
            time.sleep(3)
            side = torch.cuda.Stream()
            buf = be.staging(grp, M * N * 2)
            n_flags = (M // world // 128) * ((N + 255) // 256)
            with torch.cuda.stream(side):
                cnt = buf.u8[buf.data_bytes: buf.data_bytes + n_flags * 4].view(torch.int32).to("cpu", non_blocking=False)
            hist = {int(v): int((cnt == v).sum()) for v in cnt.unique()}
            print("rank %d: %s K=%d arrival counters %s, compute stream done=%s" % (
                rank, layout, K, hist, torch.cuda.current_stream().query()), file=sys.stderr, flush=True)
        try:
            torch.cuda.synchronize()
        except RuntimeError as exc:
            print("rank %d: fused GEMM+RS failed (%s), error info %s" % (rank, str(exc).splitlines()[0], be.comm.error_info()), file=sys.stderr, flush=True)
            raise
        full = torch.matmul(a, b.t() if layout == "tn" else b)
        ref = torch.empty(M // world, N, device="cuda", dtype=BF)
        dist.reduce_scatter_tensor(ref, full)
        torch.cuda.synchronize()
        err = float((out.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-6))
        for _ in range(3):   # repeated use: counters must reset, buffers must be reusable
            out2 = be.gemm_reduce_scatter(a, b, layout, grp)
        try:
            torch.cuda.synchronize()
        except RuntimeError:
            print("rank %d: repeated fused GEMM+RS failed, error info %s" % (rank, be.comm.error_info()), file=sys.stderr, flush=True)
            raise
        same = bool(torch.equal(out.view(torch.int16), out2.view(torch.int16)))
        fused_ms = timed(lambda: be.gemm_reduce_scatter(a, b, layout, grp))

        def unfused():
            staged, _ = be.staging_tensor(grp, (M, N), BF)
            be.gemm(a, b, layout, out=staged)
            return be.reduce_scatter_first_dim(staged, grp)
        unfused_ms = timed(unfused)

        def nccl():
            f = torch.matmul(a, b.t() if layout == "tn" else b)
            dist.reduce_scatter_tensor(ref, f)
        nccl_ms = timed(nccl)
        gemm_ms = timed(lambda: be.gemm(a, b, layout))
        ok = err < 2e-2 and same
        ok_all &= ok
        if rank == 0:
            print(json.dumps({"op": "gemm_reduce_scatter", "layout": layout, "M": M, "N": N, "K": K, "p": world, "fused_ms": round(fused_ms, 4),
                              "ours_gemm_then_rs_ms": round(unfused_ms, 4), "cublas_then_nccl_rs_ms": round(nccl_ms, 4),
                              "gemm_only_ms": round(gemm_ms, 4), "max_rel_err_vs_nccl": round(err, 5), "deterministic_repeat": same,
                              "rs_busGBps_if_alone": round(M * N * 2 * (world - 1) / world / max(fused_ms - gemm_ms, 1e-3) / 1e6, 1)}), flush=True)
    assert be.comm.error_flag() == 0
    dist.barrier()
    if rank == 0:
        print("FUSED_OK" if ok_all else "FUSED_FAIL", flush=True)
    from hetu_galvatron_b200.core.runtime.backend import reset_backend
    reset_backend()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
