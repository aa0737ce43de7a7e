"""The ZeRO-3 rotating buffer pool (hetu-galvatron_b200/core/runtime/parallel.py SlotPool): slot choice must be a pure function of
the acquire / release sequence (every rank of the group walks the same sequence, so peers' pushes land in the matching slot),
a released copy is reused without a gather until its slot is reclaimed, a prefetch never steals, a demand acquire evicts only
layers that are not in use, and exhaustion is a loud error."""
import pytest

from hetu_galvatron_b200.core.runtime.comm_groups import CommGroup
from hetu_galvatron_b200.core.runtime.parallel import SlotPool

import torch


class _Buf:
    def __init__(self, nbytes):
        self.nbytes = nbytes


class _Backend:
    def __init__(self):
        self.allocs = []

    def sym_alloc(self, group, nbytes):
        self.allocs.append(nbytes)
        return _Buf(nbytes)


class _Unit:
    def __init__(self, name, padded):
        self.name, self.padded, self._in_use, self.evicted = name, padded, 0, 0
        self.pool = None

    def evict(self, role):
        self.evicted += 1
        self.pool.release(self, "evict-%s" % self.name)


def _pool(n_slots, sizes=(100, 300, 200)):
    be = _Backend()
    pool = SlotPool(be, CommGroup([0, 1]), torch.bfloat16, n_slots, "param")
    units = [_Unit("u%d" % i, s) for i, s in enumerate(sizes)]
    for u in units:
        u.pool = pool
        pool.register(u)
    return be, pool, units


def test_sized_for_the_largest_layer_and_allocated_once():
    be, pool, units = _pool(3)
    pool.finalize()
    pool.finalize()
    assert be.allocs == [600, 600, 600] and pool.nbytes() == 1800


def test_slot_sequence_is_deterministic_and_fifo():
    picks = []
    for _ in range(2):                      # two "ranks" walking the same program
        _, pool, (a, b, c) = _pool(2)
        seq = []
        s, cached = pool.acquire(a, 1); seq.append((s.index, cached))
        s, cached = pool.acquire(b, 1); seq.append((s.index, cached))
        pool.release(a, "ev-a")
        s, cached = pool.acquire(c, 1); seq.append((s.index, cached, s.free_event))
        pool.release(b, "ev-b")
        pool.release(c, "ev-c")
        s, cached = pool.acquire(a, 1); seq.append((s.index, cached, s.free_event))
        picks.append(seq)
    assert picks[0] == picks[1] == [(0, False), (1, False), (0, False, "ev-a"), (1, False, "ev-b")]


def test_released_copy_is_reused_until_reclaimed():
    _, pool, (a, b, c) = _pool(3)
    sa, _ = pool.acquire(a, 7)
    pool.release(a, "ev")
    s, cached = pool.acquire(a, 7)                    # same version, slot untouched: no gather needed
    assert cached and s is sa and pool.n_gather_skipped == 1
    pool.release(a, "ev")
    s, cached = pool.acquire(a, 8)                    # parameters changed (optimizer step): must gather again
    assert not cached
    pool.release(a, "ev")
    for version, u in enumerate((b, c, b, c)):        # other layers (fresh versions: no reuse) cycle through every slot
        pool.acquire(u, 100 + version)
        pool.release(u, "ev")
    _, cached = pool.acquire(a, 8)
    assert not cached                                 # a's copy was overwritten meanwhile


def test_prefetch_never_steals_and_demand_evicts_idle_prefetches():
    _, pool, (a, b, c) = _pool(2)
    pool.acquire(a, 1); a._in_use = 1                 # running
    pool.acquire(b, 1)                                # prefetched, idle
    assert pool.acquire(c, 1, demand=False) == (None, False)
    s, cached = pool.acquire(c, 1, demand=True)       # c is needed now: the idle prefetch gives way
    assert b.evicted == 1 and not cached and s.free_event == "evict-u1"
    c._in_use = 1
    with pytest.raises(RuntimeError, match="exhausted"):
        pool.acquire(b, 1, demand=True)               # both slots hold layers in use
