"""The execution backend the layer/runtime code talks to.

Product = ``CudaBackend``: every method lands in a hand-written sm_100a kernel through the C ABI (``_bg``), or --
for attention only -- in the flash-attn library the reference itself calls (transformer.py:495).
There is no CPU implementation in this package: ``get_backend()`` raises when the extension or a GPU is missing.

Host-logic tests run the same layer/schedule code on CPU by *injecting* a backend with ``set_backend()``; that
backend (``oracle/gloo_backend.py``) lives with the oracle, is never imported from here, and is only ever installed
by tests/ and bench.py's reference arm.
"""
import os

import torch

from . import world as _world

_BACKEND = None


def set_backend(backend):
    global _BACKEND
    _BACKEND = backend
    return backend


def get_backend():
    global _BACKEND
    if _BACKEND is None:
        _BACKEND = CudaBackend()
    return _BACKEND


def reset_backend():
    global _BACKEND
    if _BACKEND is not None and hasattr(_BACKEND, "close"):
        _BACKEND.close()
    _BACKEND = None


class CudaBackend:
    """B200 backend: symmetric-memory communicator + fused kernels."""

    name = "cuda-sm100a"
    is_cuda = True
    FLAG_BYTES = 1 << 16

    def __init__(self, comm=None, arena_bytes=None, device=None):
        from ... import _bg
        if not torch.cuda.is_available():
            raise _bg.BgError("hetu-galvatron_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.bg = _bg
        _bg.lib()
        self.rank, self.world = _world.get_rank(), _world.get_world_size()
        self.device_index = _world.get_local_rank() % torch.cuda.device_count() if device is None else device
        torch.cuda.set_device(self.device_index)
        self.device = torch.device("cuda", self.device_index)
        if comm is None:
            if arena_bytes is None:
                arena_bytes = int(os.environ.get("HGB_ARENA_BYTES", 0))
                if not arena_bytes:
                    try:
                        from .arguments import get_args
                        arena_bytes = int(getattr(get_args(), "arena_bytes", 0))
                    except RuntimeError:
                        arena_bytes = 0
                arena_bytes = arena_bytes or (1 << 30)
            # NVLS (default on a multi-GPU job; HGB_NVLS=0 switches it off): the arena comes from the virtual-memory API so that
            # NVSwitch multicast objects can bind it; collectives on multicast-bound buffers then reduce / replicate inside
            # the switch above 1 MiB.  A driver or fabric without multicast keeps the cudaMalloc + cudaIpc arena.
            self.nvls = os.environ.get("HGB_NVLS", "1") == "1" and self.world > 1
            comm = None
            if self.nvls:
                try:
                    comm = _bg.BgComm(self.rank, self.world, self.device_index, arena_bytes, vmm=True)
                    if not comm.arena_mode()[1]:
                        comm.close()
                        comm = None
                except _bg.BgError:
                    comm = None
                if self.world > 1:      # every rank must take the same decision (one box, one driver: it does; checked anyway)
                    import torch.distributed as dist
                    flags = [None] * self.world
                    dist.all_gather_object(flags, comm is not None)
                    if not all(flags):
                        if comm is not None:
                            comm.close()
                        comm = None
                self.nvls = comm is not None
            if comm is None:
                comm = _bg.BgComm(self.rank, self.world, self.device_index, arena_bytes, vmm=False)
            if self.world > 1:
                comm.connect_vmm() if self.nvls else comm.connect_ipc()
        self.comm = comm
        if os.environ.get("HGB_TIMEOUT_MS"):      # device-side barrier timeout (default 60 s): shorter for debugging runs
            _bg.set_tunable("timeout_ms", int(os.environ["HGB_TIMEOUT_MS"]))
        # communication streams are HIGH priority: the CTA distributor then places a slim collective CTA (which fits beside the
        # persistent GEMM CTAs) ahead of any compute grid that is still waiting for SMs, instead of queueing it behind that grid
        self.unshard_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.reduce_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.p2p_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.comm_stream = torch.cuda.Stream(device=self.device, priority=-1)  # push kernels of the fused all-gather + GEMM; overlapped gathers
        self.fuse_gemm_rs = os.environ.get("HGB_FUSE_GEMM_RS", "1") != "0"
        self.fuse_gemm_ar = os.environ.get("HGB_FUSE_GEMM_AR", "1") != "0"
        self.fuse_ag_gemm = os.environ.get("HGB_FUSE_AG_GEMM", "1") != "0"
        self.n_fused = {"gemm_rs": 0, "gemm_ar": 0, "ag_gemm": 0}
        self.comm_profile = None   # bench.py: {kind: [(start_event, end_event, algorithmic_bus_bytes)]} while timing the collectives
        self.attn_impl = os.environ.get("HGB_ATTN", "cudnn")
        self._staging = {}  # group ranks -> SymBuffer
        self._scratch = {}
        self.gemm_profile = None   # bench.py: list of (start_event, end_event, flops) while timing the dominant kernel

    def close(self):
        if self.comm is not None:
            torch.cuda.synchronize()
            self.comm.close()
            self.comm = None

    # ---- in-step timing of the collectives (bench.py path legs) ----------------------------------------------------
    def _timed(self, kind, bus_bytes, stream=None, flops=0.0):
        """Context manager: when ``comm_profile`` is a dict, bracket the launch with CUDA events on its stream and record the
        algorithmic bus bytes (nccl-tests convention: AG/RS/A2A (p-1)/p*N, AR 2(p-1)/p*N, p2p N) of the call -- and, for a fused
        GEMM + collective, the GEMM's FLOPs (its roofline is the slower of FLOPs / GEMM peak and bytes / NVLink)."""
        return _Timed(self, kind, bus_bytes, stream, flops)

    # ---- memory -------------------------------------------------------------------------------------------
    def sym_alloc(self, group, nbytes):
        return self.comm.sym_alloc(group, nbytes)

    NVLS_MIN_BYTES = 1 << 20     # below this the one-shot / two-shot peer kernels win (latency)

    def exchange(self):
        self.comm.exchange()
        if getattr(self, "nvls", False):
            self.nvls_regions = self.comm.setup_nvls()

    def reserve_staging(self, group, nbytes):
        """Per-group activation staging buffer (peer-visible).  Must be called (identically on all members) before
        ``exchange()``; later requests larger than the reservation raise."""
        if group is None or group.size == 1:
            return None
        key = tuple(group.ranks)
        cur = self._staging.get(key)
        nbytes = (int(nbytes) + 255) // 256 * 256
        if cur is None or cur.data_bytes < nbytes:
            if cur is not None and cur.offsets is not None:
                raise self.bg.BgError("staging buffer for group %s is %d B, need %d B (reserve before exchange())" % (key, cur.data_bytes, nbytes))
            # three regions of `nbytes`: [0] general staging / all-gather landing, [1] partial tiles of the fused GEMM +
            # reduce-scatter / all-reduce, [2] the all-reduce result every member broadcasts into; tail: arrival counters
            # (first half: per-tile counters of the scattering GEMMs, second half: per-block counters of the gathering GEMM)
            buf = self.comm.sym_alloc(group, 3 * nbytes + self.FLAG_BYTES)
            buf.data_bytes = nbytes
            buf.u8[3 * nbytes:].zero_()
            self._staging[key] = buf
        return self._staging[key]

    def staging(self, group, nbytes, byte_offset=0):
        buf = self._staging.get(tuple(group.ranks))
        if buf is None or buf.data_bytes < nbytes + byte_offset:
            raise self.bg.BgError("no staging buffer of %d B reserved for group %s" % (nbytes + byte_offset, group.ranks))
        return buf

    def staging_tensor(self, group, shape, dtype, byte_offset=0):
        numel = 1
        for s in shape:
            numel *= int(s)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        buf = self.staging(group, nbytes, byte_offset)
        return buf.u8[byte_offset: byte_offset + nbytes].view(dtype).view(*shape), buf

    def _is_staging(self, t, buf):
        base = buf.u8.data_ptr()
        return base <= t.data_ptr() < base + buf.data_bytes

    def _stage(self, x, group, byte_offset=0):
        """Make ``x`` peer-visible: no-op when it already lives in the group's staging buffer."""
        buf = self.staging(group, x.numel() * x.element_size(), byte_offset)
        if self._is_staging(x, buf):
            return buf, x.data_ptr() - buf.u8.data_ptr()
        dst = buf.u8[byte_offset: byte_offset + x.numel() * x.element_size()].view(x.dtype)
        dst.copy_(x.reshape(-1))
        return buf, byte_offset

    # ---- sharded-unit collectives (side streams) ------------------------------------------------------------------
    def unit_unshard(self, unit):
        """C1 on the unshard stream: all-gather + fp32->bf16 cast of the unit's flat parameter."""
        with torch.cuda.stream(self.unshard_stream):
            if getattr(unit, "_w_wait_event", None) is not None:    # pooled slot: its previous occupant's last use
                self.unshard_stream.wait_event(unit._w_wait_event)
            if unit.dp_type == "ddp":
                self.cast(unit.flat_param.data, unit.w_flat)
            else:
                d = unit.group.size
                with self._timed("sdp_all_gather", (d - 1) / d * unit.padded * 2, self.unshard_stream):
                    self.comm.all_gather_cast(unit.group, unit.flat_param.data, unit.W, shard_elems=unit.shard_elems,
                                              lane=self.bg.LANE_UNSHARD, dst_dtype=unit.param_dtype)
            unit._unshard_event = torch.cuda.Event()
            unit._unshard_event.record(self.unshard_stream)

    def unit_wait_unshard(self, unit):
        if unit._unshard_event is not None:
            torch.cuda.current_stream().wait_event(unit._unshard_event)
            unit._unshard_event = None

    def begin_step(self):
        """The optimizer (current stream) has rewritten the masters: the unshard stream must see that."""
        self.unshard_stream.wait_stream(torch.cuda.current_stream())

    def unit_reduce(self, unit, accumulate):
        """C2/C3 on the reduce stream, ordered after the unit's backward on the current stream."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.reduce_stream):
            self.reduce_stream.wait_event(ev)
            if unit.dp_type != "ddp":
                d, gsz = unit.group.size, (4 if unit.reduce_dtype == torch.float32 else 2)
                with self._timed("sdp_reduce_scatter", (d - 1) / d * unit.padded * gsz, self.reduce_stream):
                    self.comm.reduce_scatter_acc(unit.group, unit.G, unit.reduce_dtype, unit.master_grad,
                                                 shard_elems=unit.shard_elems, prescale=1.0 / unit.prediv,
                                                 postscale=1.0 / unit.postdiv, accumulate=accumulate, lane=self.bg.LANE_REDUCE)
            elif unit.group.size == 1:
                self.cast(unit.g_flat, unit.master_grad, accumulate=accumulate)
            else:  # DDP layers: all-reduce of the full flat gradient (_runtime_utils.py:932-950), then cast/accumulate
                key = (unit.reduce_dtype, unit.padded)
                tmp = self._scratch.get(key)
                if tmp is None:
                    tmp = self._scratch[key] = torch.empty(unit.padded, dtype=unit.reduce_dtype, device=self.device)
                self.comm.all_reduce(unit.group, unit.G, tmp, elems=unit.padded, scale=1.0 / (unit.prediv * unit.postdiv),
                                     lane=self.bg.LANE_REDUCE)
                self.cast(tmp, unit.master_grad, accumulate=accumulate)

    def unit_reduce_adamw(self, unit, opt):
        """C2 with the AdamW epilogue (reduce stream): gradients are consumed in registers, no fp32 gradient shard."""
        lr, b1, b2, eps, wd, step = opt.hyper()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.reduce_stream):
            self.reduce_stream.wait_event(ev)
            d, gsz = unit.group.size, (4 if unit.reduce_dtype == torch.float32 else 2)
            with self._timed("sdp_reduce_scatter_adamw", (d - 1) / d * unit.padded * gsz, self.reduce_stream):
                self.comm.reduce_scatter_adamw(unit.group, unit.G, unit.reduce_dtype, unit.flat_param.data, unit.exp_avg, unit.exp_avg_sq,
                                               unit.shard_elems, 1.0 / unit.prediv, 1.0 / unit.postdiv, lr, b1, b2, eps, wd, step,
                                               lane=self.bg.LANE_REDUCE)

    def finish_reductions(self):
        torch.cuda.current_stream().wait_stream(self.reduce_stream)

    # ---- checkpoint export: gather the fp32 master shards of one unit (C1 with an fp32 destination) -----------------------
    def reserve_checkpoint_gather(self, group, nbytes):
        """Peer-visible landing buffer for ``gather_master`` (one per group, sized for its largest unit); before ``exchange()``."""
        if group is None or group.size == 1:
            return
        bufs = self.__dict__.setdefault("_ckpt_bufs", {})
        key = tuple(group.ranks)
        cur = bufs.get(key)
        if cur is None or cur.nbytes < nbytes:
            if cur is not None and cur.offsets is not None:
                raise self.bg.BgError("checkpoint gather buffer for group %s is already exchanged" % (key,))
            bufs[key] = self.comm.sym_alloc(group, int(nbytes))

    def gather_master(self, unit):
        """Full fp32 flat parameter of ``unit`` on every member of its group (collective)."""
        if unit.dp_type == "ddp" or unit.group.size == 1:
            return unit.flat_param.data
        buf = self.__dict__.get("_ckpt_bufs", {}).get(tuple(unit.group.ranks))
        if buf is None or buf.nbytes < unit.padded * 4:
            raise self.bg.BgError("no checkpoint gather buffer reserved for group %s: construct the model with args.save set" % (unit.group.ranks,))
        self.comm.all_gather_cast(unit.group, unit.flat_param.data, buf, shard_elems=unit.shard_elems, lane=self.bg.LANE_MISC,
                                  dst_dtype=torch.float32)
        return buf.view(torch.float32, unit.padded).clone()

    def barrier_all(self):
        torch.cuda.synchronize()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()

    # ---- events (ordering between successive occupants of a pooled zero3 buffer) ------------------------------------------
    def record_event(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return ev

    def reduce_done_event(self):
        ev = torch.cuda.Event()
        ev.record(self.reduce_stream)
        return ev

    def wait_event(self, ev):
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def make_stage_link(self, my_rank, peer_rank, max_bytes, send_flag_base, recv_flag_base):
        from .pipeline.pipeline import _StageLink
        return _StageLink(self, my_rank, peer_rank, max_bytes, send_flag_base, recv_flag_base)

    # ---- activation collectives (compute stream, LANE_ACT) ---------------------------------------------------
    def all_reduce(self, x, group, op="sum"):
        if group is None or group.size == 1:
            return x
        x = x.contiguous()
        if x.numel() % _vec(x):
            out = torch.empty_like(x)
            self._all_reduce_padded(x, group, op, out)
            return out
        buf, off = self._stage(x, group)
        out = torch.empty_like(x)
        # (above 1 MiB on a multicast-bound staging buffer the C side reduces and replicates inside the NVSwitch)
        with self._timed("all_reduce", 2.0 * (group.size - 1) / group.size * x.numel() * x.element_size()):
            self.comm.all_reduce(group, buf, out, elems=x.numel(), op=self.bg.MAX if op == "max" else self.bg.SUM,
                                 src_byte_offset=off)
        return out

    def all_reduce_inplace(self, x, group):
        """Sum ``x`` over ``group`` IN PLACE when that is possible without a copy: NVLS is on (and HGB_NVLS_INPLACE=1), ``x`` is a
        contiguous view of the group's multicast-bound staging buffer and large enough.  Returns whether it did."""
        if not (getattr(self, "nvls", False) and os.environ.get("HGB_NVLS_INPLACE", "0") == "1" and self.comm.has_nvls(group)):
            return False
        buf = self._staging.get(tuple(group.ranks))
        nbytes = x.numel() * x.element_size()
        if (buf is None or not x.is_contiguous() or not self._is_staging(x, buf) or x.dtype not in (torch.bfloat16, torch.float32)
                or nbytes < self.NVLS_MIN_BYTES or nbytes % 16
                or not self.comm.has_nvls(group, buf, x.data_ptr() - buf.u8.data_ptr(), nbytes)):
            return False
        region = self.comm._nvls[tuple(group.ranks)]
        self.comm.all_reduce_nvls(group, x.data_ptr() - self.comm.arena_ptr - region[0], None, x.numel(), x.dtype)
        return True

    def _all_reduce_padded(self, x, group, op, out):
        n = x.numel()
        padded = (n + 7) // 8 * 8
        buf = self.staging(group, padded * x.element_size())
        tmp_in = buf.u8[: padded * x.element_size()].view(x.dtype)
        tmp_in[:n].copy_(x.reshape(-1))
        tmp_in[n:].zero_()
        tmp_out = torch.empty(padded, dtype=x.dtype, device=x.device)
        self.comm.all_reduce(group, buf, tmp_out, elems=padded, op=self.bg.MAX if op == "max" else self.bg.SUM)
        out.copy_(tmp_out[:n].view_as(out))

    def all_gather_first_dim(self, x, group):
        """[s, ...] -> [n*s, ...] in rank order (mappings_group.py:84-102 _gather_along_first_dim)."""
        if group is None or group.size == 1:
            return x
        x = x.contiguous()
        buf, off = self._stage(x, group)
        n = group.size
        out = torch.empty((n * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        chunk = x.numel()
        self._check_vec(x, chunk)
        with self._timed("all_gather", (n - 1) / n * out.numel() * out.element_size()):
            self.comm.all_to_all_rows(group, [dict(src=buf, src_byte_offset=off, dst=out, batch=1, rows=1, row_elems=chunk,
                                                   src_bs=0, src_rs=0, src_me_off=0, dst_bs=0, dst_rs=0, dst_peer_off=chunk)], x.dtype)
        return out

    def all_gather_into_staging(self, x, group, overlap=False):
        """Push all-gather whose result stays in the group's staging buffer (operand of the next GEMM only):
        the Megatron-SP gather of layers.py:399-413.  ``overlap``: launch it on the communication stream, ordered after the
        work already in the current stream, and return (tensor, event) -- the caller runs independent work (the dgrad GEMM,
        as layers.py:449-462 overlaps them) and waits for the event before touching the gathered tensor."""
        x = x.contiguous()
        n = group.size
        out, buf = self.staging_tensor(group, (n * x.shape[0],) + tuple(x.shape[1:]), x.dtype)
        self._check_vec(x, x.numel())
        bus = (n - 1) / n * out.numel() * out.element_size()
        if not overlap:
            with self._timed("all_gather", bus):
                self.comm.all_gather_cast(group, x, buf, shard_elems=x.numel(), lane=self.bg.LANE_ACT, dst_dtype=x.dtype)
            return out
        cur = torch.cuda.current_stream()
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            with self._timed("all_gather", bus, self.comm_stream):
                self.comm.all_gather_cast(group, x, buf, shard_elems=x.numel(), lane=self.bg.LANE_PUSH, dst_dtype=x.dtype)
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
        x.record_stream(self.comm_stream)
        return out, ev

    def reduce_scatter_first_dim(self, x, group):
        """[n*s, ...] -> [s, ...] sum (mappings_group.py:105-122 _reduce_scatter_along_first_dim)."""
        if group is None or group.size == 1:
            return x
        x = x.contiguous()
        n = group.size
        assert x.shape[0] % n == 0, "First dimension of the tensor should be divisible by tensor parallel size"
        buf, off = self._stage(x, group)
        out = torch.empty((x.shape[0] // n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        self._check_vec(x, out.numel())
        with self._timed("reduce_scatter", (n - 1) / n * x.numel() * x.element_size()):
            self.comm.reduce_scatter_acc(group, buf, x.dtype, out, shard_elems=out.numel(), lane=self.bg.LANE_ACT, src_byte_offset=off)
        return out

    def all_gather_last_dim(self, x, group):
        """mappings_group.py:63-81 _gather_along_last_dim: gather then concatenate along the last dimension."""
        if group is None or group.size == 1:
            return x
        n = group.size
        g = self.all_gather_first_dim(x.contiguous().unsqueeze(0), group)  # [n, ...]
        return torch.cat([g[i] for i in range(n)], dim=-1).contiguous()

    def _check_vec(self, x, numel):
        if numel % _vec(x):
            raise self.bg.BgError("collective chunk of %d %s elements is not a multiple of 16 bytes" % (numel, x.dtype))

    def ulysses_all_to_all(self, tensors, group, to_heads):
        """Ulysses exchange of [b, s, n, d] tensors, all in ONE launch (transformer.py:1928-1987 + :2132-2145).
        to_heads=True : [b, s/p, n, d] -> [b, s, n/p, d]  (scatter heads, gather sequence; q/k/v before attention)
        to_heads=False: [b, s, n/p, d] -> [b, s/p, n, d]  (inverse; context after attention)"""
        p = group.size
        if p == 1:
            return list(tensors)
        descs, outs, off = [], [], 0
        for t in tensors:
            t = t.contiguous()
            b, s_in, n_in, d = t.shape
            buf, boff = self._stage(t, group, byte_offset=off)
            off = max(off, (boff + t.numel() * t.element_size() + 255) // 256 * 256)
            if to_heads:
                assert n_in % p == 0, "Number of heads (%d) must be divisible by the sequence parallel size (%d)!" % (n_in, p)
                hp = n_in // p
                out = torch.empty(b, s_in * p, hp, d, dtype=t.dtype, device=t.device)
                descs.append(dict(src=buf, src_byte_offset=boff, dst=out, batch=b, rows=s_in, row_elems=hp * d,
                                  src_bs=s_in * n_in * d, src_rs=n_in * d, src_me_off=hp * d,
                                  dst_bs=s_in * p * hp * d, dst_rs=hp * d, dst_peer_off=s_in * hp * d))
            else:
                assert s_in % p == 0
                sl = s_in // p
                out = torch.empty(b, sl, n_in * p, d, dtype=t.dtype, device=t.device)
                descs.append(dict(src=buf, src_byte_offset=boff, dst=out, batch=b, rows=sl, row_elems=n_in * d,
                                  src_bs=s_in * n_in * d, src_rs=n_in * d, src_me_off=sl * n_in * d,
                                  dst_bs=sl * n_in * p * d, dst_rs=n_in * p * d, dst_peer_off=n_in * d))
            outs.append(out)
        with self._timed("all_to_all", (p - 1) / p * sum(t.numel() * t.element_size() for t in tensors)):
            self.comm.all_to_all_rows(group, descs, tensors[0].dtype)
        return outs

    # ---- math ops --------------------------------------------------------------------------------------------
    def gemm(self, a, b, layout, out=None, accumulate=False, m=None, n=None, k=None, addend=None):
        """bf16 GEMM on tcgen05.  layout 'tn': a[M,K] b[N,K]; 'nn': a[M,K] b[K,N]; 'nt': a[K,M] b[K,N].
        ``addend`` [M,N]: out = A op B + addend in the epilogue (the residual add behind a projection, one rounding)."""
        code = {"tn": 0, "nn": 1, "nt": 2}[layout]
        if code == 2:
            k_, m_ = a.shape
        else:
            m_, k_ = a.shape
        n_ = b.shape[0] if code == 0 else b.shape[1]
        if out is None:
            out = torch.empty(m_, n_, dtype=torch.bfloat16, device=a.device)
        if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
            raise self.bg.BgError("the B200 GEMM path is bf16-only (mixed_precision must be bf16)")
        assert a.is_contiguous() and b.is_contiguous() and out.is_contiguous()
        if m_ % 8 or n_ % 8 or k_ % 8:
            raise self.bg.BgError("GEMM dims (%d,%d,%d) must be multiples of 8" % (m_, n_, k_))
        if addend is not None:
            assert not accumulate and addend.dtype == torch.bfloat16 and addend.is_contiguous() and addend.numel() == m_ * n_
            launch = lambda: self.bg.gemm_bf16_add(a, b, out, addend, m_, n_, k_, code)  # noqa: E731
        else:
            launch = lambda: self.bg.gemm_bf16(a, b, out, m_, n_, k_, code, accumulate=accumulate)  # noqa: E731
        if self.gemm_profile is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launch()
            e1.record()
            extra = 2 if (accumulate or addend is not None) else 1
            self.gemm_profile.append((e0, e1, 2.0 * m_ * n_ * k_, 2.0 * (m_ * k_ + k_ * n_ + m_ * n_ * extra)))
        else:
            launch()
        return out

    # measured (profiles/r01_fused_gemm_rs_4gpu_8gpu.jsonl, r02_fused_gemm_collectives_2gpu.jsonl, r02_fused_8gpu.jsonl): fusion pays when the
    # GEMM lasts at least as long as the transfer of its output -- K = 3584 at p = 4: 0.226 vs 0.258 ms; below (K = 1792 at p = 8: 0.213 vs
    # 0.197 unfused) the GEMM outruns NVLink and the fused kernel only adds its reducer tail
    FUSE_MIN_K = 3072

    def can_fuse_gemm_rs(self, m, n, group, k=None):
        p = 1 if group is None else group.size
        if not self.fuse_gemm_rs or p < 2 or m % (p * 128) or n % 8:
            return False
        if k is not None and k < self.FUSE_MIN_K and os.environ.get("HGB_FUSE_GEMM_RS") != "force":
            return False
        buf = self._staging.get(tuple(group.ranks))
        tiles = (m // p // 128) * ((n + 255) // 256)
        return buf is not None and buf.data_bytes >= m * n * 2 and tiles * 4 <= self.FLAG_BYTES // 2

    FUSE_AR_MIN_K = 2048   # as FUSE_MIN_K, for the fused GEMM + all-reduce (wins 9-14 % at K >= 2048, p = 2; loses 6-20 % at K <= 1792, p = 8)

    @staticmethod
    def _mnk(a, b, layout):
        code = {"tn": 0, "nn": 1, "nt": 2}[layout]
        if code == 2:
            k_, m_ = a.shape
        else:
            m_, k_ = a.shape
        return code, m_, (b.shape[0] if code == 0 else b.shape[1]), k_

    def can_fuse_gemm_ar(self, m, n, group, k=None):
        p = 1 if group is None else group.size
        if not self.fuse_gemm_ar or p < 2 or m % (p * 128) or n % 8:
            return False
        if k is not None and k < self.FUSE_AR_MIN_K and os.environ.get("HGB_FUSE_GEMM_AR") != "force":
            return False
        buf = self._staging.get(tuple(group.ranks))
        tiles = (m // p // 128) * ((n + 255) // 256)
        return buf is not None and buf.data_bytes >= m * n * 2 and tiles * 4 <= self.FLAG_BYTES // 2 // 2

    def gemm_all_reduce(self, a, b, layout, group):
        """[M, N] = sum over ``group`` of A op B, on every member (GEMM + C5/C6 in one operation): partial tiles go to their
        owner's HBM, the owner's tile reducer sums them as they land and broadcasts the rows into every member's result."""
        code, m_, n_, k_ = self._mnk(a, b, layout)
        buf = self.staging(group, m_ * n_ * 2)
        r = buf.data_bytes
        with self._timed("gemm_all_reduce", 2.0 * (group.size - 1) / group.size * m_ * n_ * 2, flops=2.0 * m_ * n_ * k_):
            self.comm.gemm_all_reduce(group, a, b, m_, n_, k_, code, buf, r, 3 * r, 2 * r)
        self.n_fused["gemm_ar"] += 1
        # the symmetric result buffer is rewritten by the next fused all-reduce of this group: hand out a private copy
        return buf.u8[2 * r: 2 * r + m_ * n_ * 2].view(torch.bfloat16).view(m_, n_).clone()

    def can_fuse_ag_gemm(self, m, k, group):
        p = 1 if group is None else group.size
        if not self.fuse_ag_gemm or p < 2 or m % (p * 128) or k % 8:
            return False
        buf = self._staging.get(tuple(group.ranks))
        return buf is not None and buf.data_bytes >= m * k * 2 and (m // 128) * 4 <= self.FLAG_BYTES // 2

    def all_gather_gemm(self, a_local, b, layout, group):
        """out[M, N] = gather_rows(a_local[M/p, K]) op B (C7 + GEMM in one operation).  Returns (out, gathered A) -- the gathered
        operand sits complete in the group's staging buffer afterwards (the wgrad GEMM of the same layer reads it there)."""
        p = group.size
        code = {"tn": 0, "nn": 1}[layout]
        ml, k_ = a_local.shape
        m_ = ml * p
        n_ = b.shape[0] if code == 0 else b.shape[1]
        buf = self.staging(group, m_ * k_ * 2)
        out = torch.empty(m_, n_, dtype=torch.bfloat16, device=a_local.device)
        with self._timed("all_gather_gemm", (p - 1) / p * m_ * k_ * 2, flops=2.0 * m_ * n_ * k_):
            self.comm.all_gather_gemm(group, a_local, b, out, m_, n_, k_, code, buf, 0, 3 * buf.data_bytes + self.FLAG_BYTES // 2,
                                      self.comm_stream)
        a_local.record_stream(self.comm_stream)
        self.n_fused["ag_gemm"] += 1
        return out, buf.u8[: m_ * k_ * 2].view(torch.bfloat16).view(m_, k_)

    def gemm_reduce_scatter(self, a, b, layout, group):
        """[M, N] = A op B, reduce-scattered along M over ``group`` -> [M/p, N]: the tcgen05 GEMM's epilogue stores each
        partial tile into the owning rank's HBM and a tile reducer sums them as they land (GEMM + C8 in one operation)."""
        code = {"tn": 0, "nn": 1, "nt": 2}[layout]
        if code == 2:
            k_, m_ = a.shape
        else:
            m_, k_ = a.shape
        n_ = b.shape[0] if code == 0 else b.shape[1]
        buf = self.staging(group, m_ * n_ * 2)
        out = torch.empty(m_ // group.size, n_, dtype=torch.bfloat16, device=a.device)
        with self._timed("gemm_reduce_scatter", (group.size - 1) / group.size * m_ * n_ * 2, flops=2.0 * m_ * n_ * k_):
            self.comm.gemm_reduce_scatter(group, a, b, m_, n_, k_, code, buf, buf.data_bytes, 3 * buf.data_bytes, out)
        self.n_fused_gemm_rs = getattr(self, "n_fused_gemm_rs", 0) + 1
        self.n_fused["gemm_rs"] += 1
        return out

    def rmsnorm_fwd(self, x, weight, eps):
        x2 = x.reshape(-1, x.shape[-1])
        y = torch.empty_like(x2)
        rstd = torch.empty(x2.shape[0], dtype=torch.float32, device=x.device)
        L = self.bg.lib()
        self.bg.check(L.bg_rmsnorm_fwd(_p(x2), _p(weight), _p(y), _p(rstd), x2.shape[0], x2.shape[1], float(eps), _s()))
        return y.view_as(x), rstd

    def rmsnorm_bwd(self, dy, x, weight, rstd):
        x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, x.shape[-1])
        dx = torch.empty_like(x2)
        npart = min(444, max(1, x2.shape[0]))       # 3 CTAs of 256 threads x 76 registers per SM
        dwp = torch.empty(npart, x2.shape[1], dtype=torch.float32, device=x.device)
        L = self.bg.lib()
        self.bg.check(L.bg_rmsnorm_bwd(_p(dy2), _p(x2), _p(weight), _p(rstd), _p(dx), _p(dwp), x2.shape[0], x2.shape[1], npart, _s()))
        return dx.view_as(x), dwp.sum(0).to(weight.dtype)

    def layernorm_fwd(self, x, weight, bias, eps):
        """LayerNorm with bias (GPT / BERT families): -> (y, mean[rows], rstd[rows]) fp32 statistics."""
        x2 = x.reshape(-1, x.shape[-1])
        y = torch.empty_like(x2)
        mean = torch.empty(x2.shape[0], dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        self.bg.check(self.bg.lib().bg_layernorm_fwd(_p(x2), _p(weight), _p(bias), _p(y), _p(mean), _p(rstd), x2.shape[0], x2.shape[1],
                                                     float(eps), _s()))
        return y.view_as(x), mean, rstd

    def layernorm_bwd(self, dy, x, weight, mean, rstd):
        x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, x.shape[-1])
        dx = torch.empty_like(x2)
        npart = min(444, max(1, x2.shape[0]))       # partial-sum rows (one per CTA; summed below)
        dwp = torch.empty(npart, x2.shape[1], dtype=torch.float32, device=x.device)
        dbp = torch.empty_like(dwp)
        self.bg.check(self.bg.lib().bg_layernorm_bwd(_p(dy2), _p(x2), _p(weight), _p(mean), _p(rstd), _p(dx), _p(dwp), _p(dbp),
                                                     x2.shape[0], x2.shape[1], npart, _s()))
        return dx.view_as(x), dwp.sum(0).to(weight.dtype), dbp.sum(0).to(weight.dtype)

    def bias_gelu_fwd(self, x, bias, tanh_form=True):
        x2 = x.reshape(-1, x.shape[-1])
        y = torch.empty_like(x2)
        self.bg.check(self.bg.lib().bg_bias_gelu(_p(x2), _p(bias) if bias is not None else None, None, _p(y), x2.shape[0], x2.shape[1],
                                                 1 if tanh_form else 0, _s()))
        return y.view_as(x)

    def bias_gelu_bwd(self, dy, x, bias, tanh_form=True):
        x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, x.shape[-1])
        dx = torch.empty_like(x2)
        self.bg.check(self.bg.lib().bg_bias_gelu(_p(x2), _p(bias) if bias is not None else None, _p(dy2), _p(dx), x2.shape[0], x2.shape[1],
                                                 1 if tanh_form else 0, _s()))
        return dx.view_as(x)

    def swiglu_fwd(self, gate_up):
        rows, two_f = gate_up.reshape(-1, gate_up.shape[-1]).shape
        y = torch.empty(gate_up.shape[:-1] + (two_f // 2,), dtype=gate_up.dtype, device=gate_up.device)
        self.bg.check(self.bg.lib().bg_swiglu_fwd(_p(gate_up), _p(y), rows, two_f // 2, _s()))
        return y

    def swiglu_bwd(self, dy, gate_up):
        rows, two_f = gate_up.reshape(-1, gate_up.shape[-1]).shape
        dgu = torch.empty_like(gate_up)
        self.bg.check(self.bg.lib().bg_swiglu_bwd(_p(dy), _p(gate_up), _p(dgu), rows, two_f // 2, _s()))
        return dgu

    def qkv_rope_fwd(self, mixed, cos, sin, ng, r, hn, stage_group=None):
        """mixed [s,b,ng*(r+2)*hn] -> q [b,s,ng*r,hn], k,v [b,s,ng,hn] (rotated).  With ``stage_group`` the
        outputs are written straight into that group's staging buffer (Ulysses source, no extra copy)."""
        s, b = mixed.shape[0], mixed.shape[1]
        shapes = [(b, s, ng * r, hn), (b, s, ng, hn), (b, s, ng, hn)]
        if stage_group is not None and stage_group.size > 1:
            outs, off = [], 0
            for sh in shapes:
                t, _ = self.staging_tensor(stage_group, sh, mixed.dtype, byte_offset=off)
                outs.append(t)
                off = (off + t.numel() * t.element_size() + 255) // 256 * 256
            q, k, v = outs
        else:
            q, k, v = [torch.empty(sh, dtype=mixed.dtype, device=mixed.device) for sh in shapes]
        self.bg.check(self.bg.lib().bg_qkv_rope(_p(mixed), _p(q), _p(k), _p(v), _p(cos), _p(sin), s, b, ng, r, hn, 0, _s()))
        return q, k, v

    def qkv_rope_bwd(self, dq, dk, dv, cos, sin, ng, r, hn):
        b, s = dq.shape[0], dq.shape[1]
        dmixed = torch.empty(s, b, ng * (r + 2) * hn, dtype=dq.dtype, device=dq.device)
        self.bg.check(self.bg.lib().bg_qkv_rope(_p(dmixed), _p(dq.contiguous()), _p(dk.contiguous()), _p(dv.contiguous()),
                                                _p(cos), _p(sin), s, b, ng, r, hn, 1, _s()))
        return dmixed

    def rope_tables(self, seq_len, head_dim, base, offset, dtype, device):
        inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=device) / head_dim))
        pos = torch.arange(seq_len, dtype=torch.float32, device=device) + offset
        freqs = torch.outer(pos, inv_freq)
        # the reference casts cos/sin to the activation dtype before applying them (apply_rotary_pos_emb)
        return torch.cos(freqs).to(dtype).float().contiguous(), torch.sin(freqs).to(dtype).float().contiguous()

    def attention(self, q, k, v, causal, softmax_scale, key_mask=None):
        """Attention is a LIBRARY call, as in the reference (transformer.py:495 calls flash-attn; K3 is not a collective and
        is outside the hot-path scope).  On B200 the fastest library in the image is cuDNN's fused SDPA (tcgen05 kernels:
        measured 1459 TFLOP/s fwd vs 370 for flash-attn 2's sm80-class kernels, profiles/r01_attention_libraries.jsonl), reached
        through torch SDPA; HGB_ATTN=flash selects flash-attn 2.  q [b,s,n,d], k/v [b,s,ng,d] (GQA un-expanded).  Differentiable."""
        import torch.nn.functional as F
        from torch.nn.attention import SDPBackend, sdpa_kernel
        if key_mask is not None:
            # BERT's padding mask (bert_hf/BertModel_sequential.py: get_extended_attention_mask): key j of sample b is visible iff
            # key_mask[b, j]; fused SDPA with an additive mask (cuDNN where it accepts the mask, the memory-efficient kernel else)
            assert not causal
            bias = torch.zeros(key_mask.shape[0], 1, 1, key_mask.shape[1], dtype=q.dtype, device=q.device)
            bias.masked_fill_(~key_mask.bool()[:, None, None, :], float("-inf"))
            with sdpa_kernel([SDPBackend.CUDNN_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]):
                o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=bias,
                                                   scale=softmax_scale, enable_gqa=k.shape[2] != q.shape[2])
            return o.transpose(1, 2)
        if self.attn_impl != "cudnn":
            return None
        with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
            o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal,
                                               scale=softmax_scale, enable_gqa=k.shape[2] != q.shape[2])
        return o.transpose(1, 2)

    def attention_prefix(self, q, k, v, softmax_scale):
        """Causal attention of a query block against a LONGER key/value prefix, diagonal aligned to the bottom-right corner
        (query i sees keys j <= i + sk - sq): one zigzag chunk of a context-parallel rank against everything before it.
        Library call, as all attention here: flash-attn 2 (its causal mask has exactly this alignment for sq != sk; torch's
        SDPA aligns top-left).  q [b,sq,n,d], k/v [b,sk,ng,d].  Differentiable."""
        from flash_attn import flash_attn_func
        return flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=softmax_scale, causal=True)

    def attention_fwd(self, q, k, v, causal, softmax_scale):
        """flash-attn 2 library call (the reference's choice); used when HGB_ATTN=flash."""
        from flash_attn.flash_attn_interface import _flash_attn_forward
        out, lse, _, rng = _flash_attn_forward(q, k, v, 0.0, softmax_scale, causal=causal, window_size_left=-1,
                                               window_size_right=-1, softcap=0.0, alibi_slopes=None, return_softmax=False)
        return out, lse, rng

    def attention_bwd(self, dout, q, k, v, out, lse, causal, softmax_scale, rng):
        from flash_attn.flash_attn_interface import _flash_attn_backward
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        _flash_attn_backward(dout.contiguous(), q, k, v, out, lse, dq, dk, dv, 0.0, softmax_scale, causal, -1, -1, 0.0, None,
                             False, rng_state=rng)
        return dq, dk, dv

    def ce_fwd(self, logits2d, target, vocab_start, tp_group):
        """vocab-parallel CE forward on [rows, V/t]: returns (loss[rows] fp32, rowmax, sum2) -- cross_entropy.py:14-100."""
        L, rows, vl = self.bg.lib(), logits2d.shape[0], logits2d.shape[1]
        code = self.bg.dtype_code(logits2d.dtype)
        rowmax = torch.empty(rows, dtype=torch.float32, device=logits2d.device)
        self.bg.check(L.bg_ce_rowmax(_p(logits2d), code, _p(rowmax), rows, vl, _s()))
        rowmax = self.all_reduce(rowmax, tp_group, op="max")
        out2 = torch.empty(rows, 2, dtype=torch.float32, device=logits2d.device)
        self.bg.check(L.bg_ce_sumexp(_p(logits2d), code, _p(target), _p(rowmax), _p(out2), rows, vl, int(vocab_start), _s()))
        out2 = self.all_reduce(out2, tp_group)
        loss = torch.log(out2[:, 0]) - out2[:, 1]
        return loss, rowmax, out2

    def ce_bwd(self, logits2d, target, rowmax, sum2, grad_loss, vocab_start):
        """in place: logits <- dlogits (cross_entropy.py:103-152)."""
        L, rows, vl = self.bg.lib(), logits2d.shape[0], logits2d.shape[1]
        self.bg.check(L.bg_ce_bwd(_p(logits2d), self.bg.dtype_code(logits2d.dtype), _p(target), _p(rowmax), _p(sum2),
                                  _p(grad_loss.contiguous()), rows, vl, int(vocab_start), _s()))
        return logits2d

    def cast(self, src, dst, scale=1.0, accumulate=False):
        self.bg.cast(src, dst, scale=scale, accumulate=accumulate)

    def launch_count(self):
        return self.bg.launch_count()


class _Timed:
    def __init__(self, be, kind, bus_bytes, stream, flops=0.0):
        self.be, self.kind, self.bus_bytes, self.stream, self.flops = be, kind, bus_bytes, stream, flops

    def __enter__(self):
        prof = self.be.comm_profile
        if prof is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record(self.stream or torch.cuda.current_stream())
        return self

    def __exit__(self, *exc):
        prof = self.be.comm_profile
        if prof is not None and exc[0] is None:
            self.e1.record(self.stream or torch.cuda.current_stream())
            prof.setdefault(self.kind, []).append((self.e0, self.e1, float(self.bus_bytes), float(self.flops)))
        return False


def _vec(x):
    return 16 // x.element_size()


def _p(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def _s():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
