"""Micro-benchmarks of the C-ABI kernels on one GPU (CUDA events, warm-up, L2-exceeding inputs).
Usage: python scripts/bench_kernels.py [gemm] [cast] [coll]  -> JSON lines on stdout."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hetu_galvatron_b200 import _bg as bg  # noqa: E402

BF = torch.bfloat16


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def gemm():
    shapes = [(0, 8192, 6144, 4096), (0, 8192, 4096, 4096), (0, 8192, 28672, 4096), (0, 8192, 4096, 14336),
              (1, 8192, 4096, 6144), (1, 8192, 4096, 28672), (2, 6144, 4096, 8192), (2, 28672, 4096, 8192),
              (2, 4096, 14336, 8192), (0, 8192, 128256, 4096), (0, 8192, 8192, 8192)]
    for layout, m, n, k in shapes:
        a = torch.randn((k, m) if layout == 2 else (m, k), device="cuda").to(BF)
        b = torch.randn((n, k) if layout == 0 else (k, n), device="cuda").to(BF)
        c = torch.empty(m, n, device="cuda", dtype=BF)
        t_mine = timeit(lambda: bg.gemm_bf16(a, b, c, m, n, k, layout))
        at, bt = (a.t() if layout == 2 else a), (b.t() if layout == 0 else b)
        t_ref = timeit(lambda: torch.matmul(at, bt, out=c))
        fl = 2.0 * m * n * k
        print(json.dumps({"bench": "gemm", "layout": layout, "m": m, "n": n, "k": k, "ms": round(t_mine, 4),
                          "tflops": round(fl / t_mine / 1e9, 1), "cublas_ms": round(t_ref, 4),
                          "cublas_tflops": round(fl / t_ref / 1e9, 1)}), flush=True)


def cast():
    n = 1 << 30
    src = torch.randn(n, device="cuda")
    dst = torch.empty(n, device="cuda", dtype=BF)
    t = timeit(lambda: bg.cast(src, dst))
    print(json.dumps({"bench": "cast_f32_bf16", "elems": n, "ms": round(t, 4), "GBps": round(n * 6 / t / 1e6, 1)}), flush=True)
    t = timeit(lambda: dst.copy_(src))
    print(json.dumps({"bench": "torch_copy_cast", "elems": n, "ms": round(t, 4), "GBps": round(n * 6 / t / 1e6, 1)}), flush=True)
    # the n=1 degenerate all-gather+cast / reduce-scatter+acc (what a 1-GPU step runs per layer)
    from hetu_galvatron_b200.core.runtime.comm_groups import CommGroup
    P = 218_112_000
    comm = bg.BgComm(0, 1, 0, (P * 2 + (1 << 20)) * 2)
    grp = CommGroup([0])
    w, g = comm.sym_alloc(grp, P * 2), comm.sym_alloc(grp, P * 2)
    master, mg = torch.randn(P, device="cuda"), torch.zeros(P, device="cuda")
    t = timeit(lambda: comm.all_gather_cast(grp, master, w))
    print(json.dumps({"bench": "all_gather_cast_n1", "elems": P, "ms": round(t, 4), "GBps": round(P * 6 / t / 1e6, 1)}), flush=True)
    t = timeit(lambda: comm.reduce_scatter_acc(grp, g, BF, mg, accumulate=True))
    print(json.dumps({"bench": "reduce_scatter_acc_n1", "elems": P, "ms": round(t, 4), "GBps": round(P * 10 / t / 1e6, 1)}), flush=True)
    comm.close()


def ops():
    """the row kernels at the Llama-3-8B shapes of one microbatch (seq 8192): achieved HBM GB/s = algorithmic bytes / time"""
    import ctypes
    L = bg.lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    rows, h, ffn = 8192, 4096, 14336
    x, dy, w = torch.randn(rows, h, device="cuda").to(BF), torch.randn(rows, h, device="cuda").to(BF), torch.randn(h, device="cuda").to(BF)
    y, dx, rstd = torch.empty_like(x), torch.empty_like(x), torch.empty(rows, device="cuda")
    dwp = torch.empty(444, h, device="cuda")
    gu, dact = torch.randn(rows, 2 * ffn, device="cuda").to(BF), torch.randn(rows, ffn, device="cuda").to(BF)
    act, dgu = torch.empty(rows, ffn, device="cuda", dtype=BF), torch.empty(rows, 2 * ffn, device="cuda", dtype=BF)
    ng, r, hn = 8, 4, 128
    mixed = torch.randn(rows, 1, ng * (r + 2) * hn, device="cuda").to(BF)
    q, k, v = (torch.empty(1, rows, ng * r, hn, device="cuda", dtype=BF), torch.empty(1, rows, ng, hn, device="cuda", dtype=BF),
               torch.empty(1, rows, ng, hn, device="cuda", dtype=BF))
    cos, sin = torch.rand(rows, hn // 2, device="cuda"), torch.rand(rows, hn // 2, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2: written between timed launches

    def timed(fn, iters=20):
        ts = []
        for _ in range(iters + 3):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return sorted(ts[3:])[len(ts[3:]) // 2]
    cases = [
        ("rmsnorm_fwd", lambda: L.bg_rmsnorm_fwd(P(x), P(w), P(y), P(rstd), rows, h, 1e-5, S()), 2 * x.numel() * 2),
        ("rmsnorm_bwd", lambda: L.bg_rmsnorm_bwd(P(dy), P(x), P(w), P(rstd), P(dx), P(dwp), rows, h, 444, S()), 3 * x.numel() * 2),
        ("swiglu_fwd", lambda: L.bg_swiglu_fwd(P(gu), P(act), rows, ffn, S()), 3 * act.numel() * 2),
        ("swiglu_bwd", lambda: L.bg_swiglu_bwd(P(dact), P(gu), P(dgu), rows, ffn, S()), 5 * act.numel() * 2),
        ("qkv_rope_fwd", lambda: L.bg_qkv_rope(P(mixed), P(q), P(k), P(v), P(cos), P(sin), rows, 1, ng, r, hn, 0, S()), 2 * mixed.numel() * 2),
        ("qkv_rope_bwd", lambda: L.bg_qkv_rope(P(mixed), P(q), P(k), P(v), P(cos), P(sin), rows, 1, ng, r, hn, 1, S()), 2 * mixed.numel() * 2),
    ]
    for name, fn, nbytes in cases:
        bg.check(fn())
        t = timed(fn)
        print(json.dumps({"bench": name, "rows": rows, "ms": round(t, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
                          "GBps": round(nbytes / t / 1e6, 1), "frac_of_hbm_peak_6567": round(nbytes / t / 1e6 / 6566.7, 3),
                          "timing": "median of 20 single launches, L2 flushed (256 MiB memset) before each"}), flush=True)


def attn():
    """K3 is a library call in the reference (flash-attn 2); compare the attention libraries present in the image."""
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel
    b, s, n, ng, d = 1, 8192, 32, 8, 128
    q = torch.randn(b, s, n, d, device="cuda", dtype=BF, requires_grad=True)
    k = torch.randn(b, s, ng, d, device="cuda", dtype=BF, requires_grad=True)
    v = torch.randn(b, s, ng, d, device="cuda", dtype=BF, requires_grad=True)
    do = torch.randn(b, s, n, d, device="cuda", dtype=BF)
    fl_f = 4.0 * b * s * s * n * d / 2
    results = {}

    def run(name, fwd):
        try:
            out = fwd()
            t_f = timeit(fwd)
            def fb():
                o = fwd()
                o.backward(do, retain_graph=False)
                q.grad = k.grad = v.grad = None
            t_fb = timeit(fb)
            results[name] = out.detach()
            print(json.dumps({"bench": "attention", "impl": name, "fwd_ms": round(t_f, 3), "fwd_tflops": round(fl_f / t_f / 1e9, 1),
                              "fwd_bwd_ms": round(t_fb, 3), "fwd_bwd_tflops": round(3.5 * fl_f / t_fb / 1e9, 1)}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"bench": "attention", "impl": name, "error": str(e)[:300]}), flush=True)

    from flash_attn import flash_attn_func
    run("flash_attn2", lambda: flash_attn_func(q, k, v, causal=True))
    for be_name, be in (("sdpa_cudnn", SDPBackend.CUDNN_ATTENTION), ("sdpa_flash", SDPBackend.FLASH_ATTENTION)):
        def f(be=be):
            with sdpa_kernel(be):
                return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True,
                                                      enable_gqa=True).transpose(1, 2)
        run(be_name, f)
    if "flash_attn2" in results:
        for name, o in results.items():
            print(json.dumps({"bench": "attention_parity", "impl": name,
                              "max_abs_diff_vs_flash_attn2": float((o.float() - results["flash_attn2"].float()).abs().max())}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "cast"]
    print(json.dumps({"device": torch.cuda.get_device_name(0)}))
    for w in which:
        globals()[w]()
