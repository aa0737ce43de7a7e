"""GPU parity of the local fused kernels (through the C ABI) against plain-torch fp32 restatements of the reference
ops: RMSNorm (flash_attn.ops.rms_norm semantics, LlamaModel_tensor_parallel.py:2,48), swiglu (transformer.py:122-124),
the QKV split + RoPE + relayout chain (transformer.py:731-767,842-867 + megatron apply_rotary_pos_emb), and the
vocab-parallel cross entropy (cross_entropy.py:14-152).  Tolerance: one bf16 rounding (2^-8) on bf16 outputs."""
import ctypes
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
BF = torch.bfloat16


@pytest.fixture(scope="module")
def bg():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import hetu_galvatron_b200._bg as bg
    bg.lib()
    return bg


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def close_bf16(got, want, extra=0.0):
    got, want = got.float(), want.float()
    tol = want.abs() * 2 ** -7 + 1e-3 + extra
    bad = (got - want).abs() > tol
    assert not bad.any(), f"{int(bad.sum())} / {bad.numel()} off, max err {float((got - want).abs().max())}"


@pytest.mark.parametrize("n,scale,acc", [(8, 1.0, False), (4096 * 8 + 8, 0.5, True), (1 << 22, 1.0, False)])
@pytest.mark.parametrize("sd,dd", [(torch.float32, BF), (BF, torch.float32), (BF, BF), (torch.float32, torch.float32)])
def test_cast(bg, n, scale, acc, sd, dd):
    src = (torch.randn(n, device="cuda") * 2).to(sd)
    dst0 = torch.randn(n, device="cuda").to(dd)
    dst = dst0.clone()
    bg.cast(src, dst, scale=scale, accumulate=acc)
    want = (src.float() * scale + (dst0.float() if acc else 0)).to(dd)
    if dd == BF and not acc and scale == 1.0:
        assert torch.equal(dst.view(torch.int16), want.view(torch.int16))
    else:
        close_bf16(dst, want) if dd == BF else torch.testing.assert_close(dst, want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("rows,cols", [(1, 8), (37, 128), (1000, 4096), (64, 8192), (50, 3072), (301, 5120), (2500, 2048)])
def test_rmsnorm_fwd_bwd(bg, rows, cols):
    x = torch.randn(rows, cols, device="cuda").to(BF)
    w = (1 + 0.1 * torch.randn(cols, device="cuda")).to(BF)
    dy = torch.randn(rows, cols, device="cuda").to(BF)
    y, rstd = torch.empty_like(x), torch.empty(rows, device="cuda")
    bg.check(bg.lib().bg_rmsnorm_fwd(_p(x), _p(w), _p(y), _p(rstd), rows, cols, 1e-5, _s()))
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)
    yf = xf * r * wf
    close_bf16(y, yf.detach())
    torch.testing.assert_close(rstd, r.detach().squeeze(-1), rtol=1e-5, atol=1e-6)
    yf.backward(dy.float())
    npart = 64
    dx, dwp = torch.empty_like(x), torch.zeros(npart, cols, device="cuda")
    bg.check(bg.lib().bg_rmsnorm_bwd(_p(dy), _p(x), _p(w), _p(rstd), _p(dx), _p(dwp), rows, cols, npart, _s()))
    close_bf16(dx, xf.grad, extra=2e-3)
    torch.testing.assert_close(dwp.sum(0), wf.grad, rtol=2e-3, atol=2e-2 * (rows ** 0.5) / 30 + 1e-3)


@pytest.mark.parametrize("rows,ffn", [(1, 8), (33, 256), (2048, 14336), (3, 1792), (4097, 3584)])
def test_swiglu(bg, rows, ffn):
    gu = torch.randn(rows, 2 * ffn, device="cuda").to(BF)
    dy = torch.randn(rows, ffn, device="cuda").to(BF)
    y = torch.empty(rows, ffn, device="cuda", dtype=BF)
    bg.check(bg.lib().bg_swiglu_fwd(_p(gu), _p(y), rows, ffn, _s()))
    guf = gu.float().requires_grad_(True)
    g, u = torch.chunk(guf, 2, dim=-1)
    yf = torch.nn.functional.silu(g) * u
    close_bf16(y, yf.detach())
    yf.backward(dy.float())
    dgu = torch.empty_like(gu)
    bg.check(bg.lib().bg_swiglu_bwd(_p(dy), _p(gu), _p(dgu), rows, ffn, _s()))
    close_bf16(dgu, guf.grad, extra=2e-3)


def _rotate_half(x):
    x1, x2 = torch.chunk(x, 2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def _ref_qkv_rope(mixed, cos, sin, ng, r, hn):
    """transformer.py:731-767 split, :853-854 apply_rotary_pos_emb (t*cos + rotate_half(t)*sin), :864 rearrange."""
    s, b = mixed.shape[:2]
    m = mixed.view(s, b, ng, (r + 2) * hn)
    q, k, v = torch.split(m, [r * hn, hn, hn], dim=3)
    q = q.reshape(s, b, ng * r, hn)
    c = torch.cat([cos, cos], -1)[:, None, None, :]
    sn = torch.cat([sin, sin], -1)[:, None, None, :]
    q = q * c + _rotate_half(q) * sn
    k = k * c + _rotate_half(k) * sn
    return [t.permute(1, 0, 2, 3).contiguous() for t in (q, k, v)]


@pytest.mark.parametrize("s,b,ng,r,hn", [(16, 1, 1, 1, 16), (64, 2, 2, 4, 64), (512, 1, 8, 4, 128), (40, 2, 16, 4, 128), (3000, 1, 1, 4, 128)])
def test_qkv_rope_fwd_bwd(bg, s, b, ng, r, hn):
    from oracle.collectives_ref import rope_tables
    mixed = torch.randn(s, b, ng * (r + 2) * hn, device="cuda").to(BF)
    cos, sin = [t.cuda().contiguous() for t in rope_tables(s, hn, offset=3)]
    q = torch.empty(b, s, ng * r, hn, device="cuda", dtype=BF)
    k = torch.empty(b, s, ng, hn, device="cuda", dtype=BF)
    v = torch.empty_like(k)
    bg.check(bg.lib().bg_qkv_rope(_p(mixed), _p(q), _p(k), _p(v), _p(cos), _p(sin), s, b, ng, r, hn, 0, _s()))
    mf = mixed.float().requires_grad_(True)
    qf, kf, vf = _ref_qkv_rope(mf, cos, sin, ng, r, hn)
    close_bf16(q, qf.detach()); close_bf16(k, kf.detach())
    assert torch.equal(v.view(torch.int16), vf.detach().to(BF).view(torch.int16))
    dq, dk, dv = [torch.randn_like(t).to(BF) for t in (qf, kf, vf)]
    (qf * dq.float()).sum().backward(retain_graph=True)
    gq = mf.grad.clone(); mf.grad = None
    ((qf * dq.float()).sum() + (kf * dk.float()).sum() + (vf * dv.float()).sum()).backward()
    dm = torch.empty_like(mixed)
    bg.check(bg.lib().bg_qkv_rope(_p(dm), _p(dq), _p(dk), _p(dv), _p(cos), _p(sin), s, b, ng, r, hn, 1, _s()))
    close_bf16(dm, mf.grad, extra=2e-3)


@pytest.mark.parametrize("rows,vocab,parts", [(5, 64, 1), (64, 1000 * 8, 4), (512, 128256, 1), (128, 128256, 8)])
@pytest.mark.parametrize("dtype", [BF, torch.float32])
def test_vocab_parallel_cross_entropy(bg, rows, vocab, parts, dtype):
    """`parts` vocab shards handled sequentially on one device: the MAX / SUM all-reduces between the kernels are
    done here with torch, exactly where cross_entropy.py:22-30 / :61-89 puts them."""
    L = bg.lib()
    vl = vocab // parts
    logits = (torch.randn(rows, vocab, device="cuda") * 3).to(dtype)
    target = torch.randint(0, vocab, (rows,), device="cuda")
    shards = [logits[:, i * vl:(i + 1) * vl].contiguous() for i in range(parts)]
    code = bg.dtype_code(dtype)
    maxes = []
    for sh in shards:
        m = torch.empty(rows, device="cuda")
        bg.check(L.bg_ce_rowmax(_p(sh), code, _p(m), rows, vl, _s()))
        maxes.append(m)
    gmax = torch.stack(maxes).max(0).values
    assert torch.equal(gmax, logits.float().max(-1).values)
    outs = []
    for i, sh in enumerate(shards):
        o = torch.empty(rows, 2, device="cuda")
        bg.check(L.bg_ce_sumexp(_p(sh), code, _p(target), _p(gmax), _p(o), rows, vl, i * vl, _s()))
        outs.append(o)
    tot = torch.stack(outs).sum(0).contiguous()
    loss = torch.log(tot[:, 0]) - tot[:, 1]
    lf = logits.float().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(lf, target, reduction="none")
    torch.testing.assert_close(loss, want.detach(), rtol=2e-5, atol=2e-5)
    gl = torch.rand(rows, device="cuda")
    want.backward(gl)
    for i, sh in enumerate(shards):
        bg.check(L.bg_ce_bwd(_p(sh), code, _p(target), _p(gmax), _p(tot), _p(gl), rows, vl, i * vl, _s()))
        wg = lf.grad[:, i * vl:(i + 1) * vl]
        if dtype == BF:
            close_bf16(sh, wg)
        else:
            torch.testing.assert_close(sh, wg, rtol=2e-5, atol=1e-6)
