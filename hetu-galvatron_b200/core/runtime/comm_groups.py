"""Comm-group builder: per-layer strategy -> rank lists (the bit-exact part of the drop-in boundary).

Same call signature and 16-tuple result as ``galvatron/core/runtime/comm_groups.py:416 gen_comm_groups``;
the reference enumerates *every* group of every kind with ``torch.distributed.new_group`` (an NCCL
communicator bootstrap per group, twice, on all ranks -- ``comm_groups.py:441-445``) and keeps the one the
caller is in.  Here a group is a pure value: each rule below computes the caller's own rank list in closed
form from the rank layout ``PP -> [DP -> CP -> (TP|SP)]`` (TP/SP innermost-consecutive, CP stride ``t*s``,
DP stride ``t*s*c``; ``comm_groups.py:101-109,126-135``), so construction is O(1) and needs no collective.
The device runtime addresses a group by its rank list (see ``_bg.BgComm.group_id``).

Reference quirks kept because they are visible in the mapping:
  * ``gen_redistributed_group_with_cp`` returns (tp_old, tp_new, cp_NEW, cp_OLD) (:306) and the caller
    unpacks it as (split_tp_sp, allgather_tp_sp, split_cp, allgather_cp) (:483) -- so ``split_cp_groups[i]``
    is the NEW layer's CP group and ``allgather_cp_groups[i]`` the OLD one.
  * the SDP ("seq_data") group ignores sp/cp: whole stage when tp==1, stride-tp otherwise (:382-405).
  * ``tp_consecutive_flags[i]`` is rewritten in place to 1 when tp in {1, world/pp} (:427-428).
Quirk rejected: a strided TP request (flag 0 with 1 < tp < world/pp) makes the reference return ``None``
groups that crash later (:78); here it raises immediately.
"""
from . import world as _world


class CommGroup(object):
    """A communication group = sorted, de-duplicated rank list (``comm_groups.py:7-29``).

    ``.group`` is what layers receive where the reference passes a ``ProcessGroup``; it is the object itself
    (the device runtime resolves it by rank list), so ``tp_group.group.size`` etc. keep working.
    """

    __slots__ = ("ranks", "size", "intra_group_id", "_handles")

    def __init__(self, ranks):
        if not isinstance(ranks, (list, tuple, range)):
            raise TypeError("Rank list or range should be provided to create a CommGroup!")
        self.ranks = sorted(set(int(r) for r in ranks))
        self.size = len(self.ranks)
        self.intra_group_id = None
        self._handles = {}

    @property
    def group(self):
        return self

    def has_rank(self, rank):
        if rank in self.ranks:
            self.intra_group_id = self.ranks.index(rank)
            return True
        return False

    def rank_in_group(self, rank=None):
        rank = _world.get_rank() if rank is None else rank
        return self.ranks.index(rank)

    def print(self):
        print(self.ranks, end=" ")

    def __repr__(self):
        return "CommGroup(%s)" % (self.ranks,)

    def __eq__(self, other):
        return isinstance(other, CommGroup) and other.ranks == self.ranks

    def __hash__(self):
        return hash(tuple(self.ranks))


_interned = {}


def _group(ranks, rank):
    """Intern: equal rank lists share one object, so handles bound by the device runtime are shared."""
    key = tuple(sorted(set(ranks)))
    g = _interned.get(key)
    if g is None:
        g = _interned[key] = CommGroup(list(key))
    g.has_rank(rank)
    return g


def show_groups(groups):
    for g in groups:
        if g is None:
            print("None", end=" ")
        else:
            g.print()
    print()


# ---------------------------------------------------------------------------------------------
# closed-form membership rules (rank, world are explicit so every rank can be enumerated in tests)
# ---------------------------------------------------------------------------------------------
def _block(rank, size):
    """Consecutive block of ``size`` containing ``rank`` (TP: :79-81, SP: :155-157, SEP: :228-230)."""
    lo = rank // size * size
    return range(lo, lo + size)


def _tp_ranks(rank, world, tp_size):
    return _block(rank, tp_size)


def _pp_ranks(rank, world, pp_size):
    per_stage = world // pp_size
    return range(rank % per_stage, world, per_stage)


def _cp_ranks(rank, world, pp_size, mul, cp):
    """CP peers: stride ``mul`` (= tp*sp) inside the rank's DP block of ``mul*cp`` (:103-111)."""
    per_stage = world // pp_size
    base = rank // per_stage * per_stage
    local = rank - base
    blk = local // (mul * cp) * (mul * cp)
    first = base + blk + local % mul
    return range(first, first + mul * cp, mul)


def _dp_ranks(rank, world, pp_size, mul, cp):
    """DP peers: stride ``mul*cp`` across the rank's pipeline stage (:130-135)."""
    per_stage = world // pp_size
    base = rank // per_stage * per_stage
    stride = mul * cp
    return range(base + (rank - base) % stride, base + per_stage, stride)


def _sdp_ranks(rank, world, pp_size, tp_size):
    """Sharded-DP ("seq_data") peers: the whole stage, thinned by tp (:386-403)."""
    per_stage = world // pp_size
    base = rank // per_stage * per_stage
    if tp_size == 1:
        return range(base, base + per_stage)
    return range(base + (rank - base) % tp_size, base + per_stage, tp_size)


def _fused_ranks(rank, big, small):
    """merge_redistributed_group (:315-360): stride-``small`` comb over the block of ``big``."""
    lo = rank // big * big
    return range(lo + (rank - lo) % small, lo + big, small)


def gen_tp_group_dist(tp_size, pp_size, to_print=True, consecutive=True, world_ranks=None, rank=None, world_size=None):
    rank, world = _resolve(rank, world_size, world_ranks)
    return _group(_tp_ranks(rank, world, tp_size), rank) if consecutive else None


def gen_sp_group_dist(sp_size, pp_size, to_print=True, consecutive=True, world_ranks=None, rank=None, world_size=None):
    rank, world = _resolve(rank, world_size, world_ranks)
    return _group(_block(rank, sp_size), rank) if consecutive else None


def gen_cp_group_dist(mul_size, cp_size, pp_size, to_print=True, consecutive=False, world_ranks=None, rank=None, world_size=None):
    rank, world = _resolve(rank, world_size, world_ranks)
    return None if consecutive else _group(_cp_ranks(rank, world, pp_size, mul_size, cp_size), rank)


def gen_dp_group_dist(mul_size, cp_size, pp_size, to_print=True, consecutive=False, world_ranks=None, rank=None, world_size=None):
    rank, world = _resolve(rank, world_size, world_ranks)
    return None if consecutive else _group(_dp_ranks(rank, world, pp_size, mul_size, cp_size), rank)


def gen_pp_group_dist(pp_size, to_print=True, world_ranks=None, rank=None, world_size=None):
    rank, world = _resolve(rank, world_size, world_ranks)
    per_stage = world // pp_size
    all_pp = [CommGroup(range(i, world, per_stage)) for i in range(per_stage)]
    return _group(_pp_ranks(rank, world, pp_size), rank), all_pp


def gen_embedding_group_dist(pp_size, all_pp_groups, to_print=True, rank=None):
    rank = _world.get_rank() if rank is None else rank
    for pp_group in all_pp_groups:
        ends = [pp_group.ranks[0], pp_group.ranks[-1]] if pp_size > 1 else [pp_group.ranks[0]]
        if rank in ends:
            return _group(ends, rank)
    return None


def gen_sep_group_dist(tp_size, cp_size, pp_size, to_print=True, consecutive=True, world_ranks=None, rank=None, world_size=None):
    rank, world = _resolve(rank, world_size, world_ranks)
    return _group(_block(rank, tp_size * cp_size), rank) if consecutive else None


def gen_seq_data_group_dist(pp_size, tp_size, to_print=False, world_ranks=None, rank=None, world_size=None):
    rank, world = _resolve(rank, world_size, world_ranks)
    return _group(_sdp_ranks(rank, world, pp_size, tp_size), rank)


def gen_redistributed_group_with_cp(tp_size_old, tp_size_new, tp_consec_old, tp_consec_new, tp_group_old, tp_group_new,
                                    cp_size_old, cp_size_new, cp_consec_old, cp_consec_new, cp_group_old, cp_group_new):
    """Same return ORDER as the reference (:298-306): (tp_old, tp_new, cp_new, cp_old)."""
    if (tp_size_old == tp_size_new and tp_consec_old == tp_consec_new
            and cp_size_old == cp_size_new and cp_consec_old == cp_consec_new):
        return (None, None, None, None)
    return (None if tp_size_old == 1 else tp_group_old, None if tp_size_new == 1 else tp_group_new,
            None if cp_size_new == 1 else cp_group_new, None if cp_size_old == 1 else cp_group_old)


def merge_redistributed_group(split_group, allgather_group, world_ranks=None, rank=None, world_size=None):
    """(fused_split, fused_allgather) for a layer boundary whose sequence-group size changes (:315-360)."""
    if split_group is None or allgather_group is None:
        return None, None
    rank, _ = _resolve(rank, world_size, world_ranks)
    s, a = split_group.size, allgather_group.size
    if s > a:
        return _group(_fused_ranks(rank, s, a), rank), None
    if s < a:
        return None, _group(_fused_ranks(rank, a, s), rank)
    return None, None


def _resolve(rank, world_size, world_ranks=None):
    if world_ranks is not None:
        # the reference's own world_ranks path is broken (get_world_size(list), :249,265); not supported
        raise NotImplementedError("world_ranks is not supported (reference path is broken, SURVEY 8g)")
    return (_world.get_rank() if rank is None else rank,
            _world.get_world_size() if world_size is None else world_size)


def gen_comm_groups(all_tp_sizes, all_sp_sizes, all_cp_sizes, pp_size, tp_consecutive_flags, show_rank=-1,
                    world_ranks=None, *, rank=None, world_size=None):
    """Per-layer groups for the calling rank.  Returns the reference's 16-tuple (:550-569):

    (pp_group, tp_groups, sp_groups, cp_groups, dp_groups, seq_data_groups, allgather_tp_sp_groups,
     split_tp_sp_groups, allgather_cp_groups, split_cp_groups, allgather_tp_sp_cp_groups,
     split_tp_sp_cp_groups, fused_allgather_groups, fused_split_groups, embedding_group, vtp_data_group)
    """
    rank, world = _resolve(rank, world_size, world_ranks)
    n = len(all_tp_sizes)
    if not (len(all_sp_sizes) == len(all_cp_sizes) == len(tp_consecutive_flags) == n):
        raise ValueError("tp/sp/cp/consec lists must have one entry per whole-model layer")
    if world % pp_size:
        raise ValueError("world size %d not divisible by pp_deg %d" % (world, pp_size))
    per_stage = world // pp_size
    for i in range(n):
        t, s, c = all_tp_sizes[i], all_sp_sizes[i], all_cp_sizes[i]
        assert t == 1 or s == 1, "DeepSpeed Ulysses is not compatible with Megatron Tensor Parallel!"
        assert tp_consecutive_flags[i] in (0, 1)
        if t in (1, per_stage):
            tp_consecutive_flags[i] = 1
        if tp_consecutive_flags[i] == 0:
            raise ValueError("layer %d: strided TP (tp_consecutive_flags=0 with 1 < tp=%d < %d) has no runtime "
                             "path in the reference (comm_groups.py:78 returns None)" % (i, t, per_stage))
        if per_stage % (t * s * c):
            raise ValueError("layer %d: tp*sp*cp=%d does not divide the %d ranks of a stage" % (i, t * s * c, per_stage))

    pp_group, all_pp_groups = gen_pp_group_dist(pp_size, to_print=False, rank=rank, world_size=world)
    embedding_group = gen_embedding_group_dist(pp_size, all_pp_groups, to_print=False, rank=rank)

    tp_groups, sp_groups, cp_groups, dp_groups, seq_data_groups = [], [], [], [], []
    for i in range(n):
        t, s, c = all_tp_sizes[i], all_sp_sizes[i], all_cp_sizes[i]
        tp_groups.append(_group(_tp_ranks(rank, world, t), rank))
        sp_groups.append(_group(_block(rank, s), rank))
        cp_groups.append(_group(_cp_ranks(rank, world, pp_size, t * s, c), rank))
        dp_groups.append(_group(_dp_ranks(rank, world, pp_size, t * s, c), rank))
        seq_data_groups.append(_group(_sdp_ranks(rank, world, pp_size, t), rank))
    vtp_data_group = dp_groups[0]

    # relocation groups for the boundary (layer i-1 -> layer i); entry 0 is always None (:434-437)
    ag_tp_sp, sp_tp_sp = [None], [None]
    ag_cp, sp_cp = [None], [None]
    ag_sep, sp_sep = [None], [None]
    fused_sp, fused_ag = [None], [None]
    for i in range(1, n):
        old_is_tp, new_is_tp = all_tp_sizes[i - 1] != 1, all_tp_sizes[i] != 1
        old_size = all_tp_sizes[i - 1] if old_is_tp else all_sp_sizes[i - 1]
        new_size = all_tp_sizes[i] if new_is_tp else all_sp_sizes[i]
        old_grp = tp_groups[i - 1] if old_is_tp else sp_groups[i - 1]
        new_grp = tp_groups[i] if new_is_tp else sp_groups[i]
        old_cp, new_cp = all_cp_sizes[i - 1], all_cp_sizes[i]

        split_tp_sp, allgather_tp_sp, split_cp, allgather_cp = gen_redistributed_group_with_cp(
            old_size, new_size, tp_consecutive_flags[i - 1], tp_consecutive_flags[i], old_grp, new_grp,
            old_cp, new_cp, False, False, cp_groups[i - 1], cp_groups[i])
        if old_size == new_size and old_cp == new_cp:
            split_sep = allgather_sep = None
        else:
            split_sep = _group(_block(rank, old_size * old_cp), rank)
            allgather_sep = _group(_block(rank, new_size * new_cp), rank)
        f_split, f_allgather = merge_redistributed_group(split_sep, allgather_sep, rank=rank, world_size=world)

        ag_tp_sp.append(allgather_tp_sp)
        sp_tp_sp.append(split_tp_sp)
        ag_cp.append(allgather_cp)
        sp_cp.append(split_cp)
        ag_sep.append(allgather_sep)
        sp_sep.append(split_sep)
        fused_sp.append(f_split)
        fused_ag.append(f_allgather)

    if show_rank >= 0 and rank == show_rank:
        print("====================== Galvatron Communication Group ===========================")
        for title, val in (("Embedding group", [embedding_group]), ("TP groups", tp_groups), ("SP groups", sp_groups),
                           ("CP groups", cp_groups), ("DP groups", dp_groups), ("SDP groups", seq_data_groups),
                           ("Split TP/SP groups", sp_tp_sp), ("AllGather TP/SP groups", ag_tp_sp),
                           ("Split CP groups", sp_cp), ("AllGather CP groups", ag_cp),
                           ("Split TP/SP/CP groups", sp_sep), ("AllGather TP/SP/CP groups", ag_sep),
                           ("Fused split groups", fused_sp), ("Fused allgather groups", fused_ag)):
            print("%s for rank %d:" % (title, show_rank))
            show_groups(val)
        print("================================================================================")

    return (pp_group, tp_groups, sp_groups, cp_groups, dp_groups, seq_data_groups, ag_tp_sp, sp_tp_sp, ag_cp, sp_cp,
            ag_sep, sp_sep, fused_ag, fused_sp, embedding_group, vtp_data_group)
