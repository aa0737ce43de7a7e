"""Process-wide view of "who am I" for the SPMD job (one process per GPU).

The reference reads ``torch.distributed.get_rank()/get_world_size()`` at every use
(e.g. ``comm_groups.py:73``); here the answer is resolved once, and can be overridden for
host-side tests that enumerate every rank of a strategy inside one process.
"""
import contextlib
import os

_override = None  # (rank, world_size)


def _from_torch():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        return int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    return 0, 1


def get_rank():
    return _override[0] if _override is not None else _from_torch()[0]


def get_world_size():
    return _override[1] if _override is not None else _from_torch()[1]


def get_local_rank():
    return int(os.environ.get("LOCAL_RANK", get_rank()))


@contextlib.contextmanager
def simulated(rank, world_size):
    """Pretend to be ``rank`` of ``world_size`` (pure integer paths only: group mapping, config expansion)."""
    global _override
    prev, _override = _override, (int(rank), int(world_size))
    try:
        yield
    finally:
        _override = prev
