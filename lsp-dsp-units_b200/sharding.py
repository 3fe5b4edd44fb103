"""Multi-GPU sharding plans (one process per GPU).

* ``channel_shard``    -- independent convolver instances: contiguous ranges, no communication.
* ``partition_shard``  -- one long impulse response split by partition range: rank ``g`` owns the
                          taps ``[p_lo*F, p_hi*F)`` and initialises its instances with
                          ``b200conv_init_range(..., part_offset=p_lo)``; the per-block outputs of
                          all ranks are summed (one fp32 all-reduce of ``channels * F`` floats).
"""


def channel_shard(instances, world, rank):
    """Instance indices [lo, hi) owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    lo = instances * rank // world
    hi = instances * (rank + 1) // world
    return lo, hi


def partition_shard(taps, frame, world, rank):
    """(p_lo, p_hi, tap_lo, tap_hi) for ``rank``: partitions of ``frame`` taps, balanced so that
    the folded-overlap row counts (p_hi - p_lo + 1) differ by at most one.  Ranks beyond the
    number of partitions get an empty range (p_lo == p_hi)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    bins = (taps + frame - 1) // frame
    p_lo = bins * rank // world
    p_hi = bins * (rank + 1) // world
    tap_lo = min(p_lo * frame, taps)
    tap_hi = min(p_hi * frame, taps)
    return p_lo, p_hi, tap_lo, tap_hi
