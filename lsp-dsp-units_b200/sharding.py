"""Multi-GPU sharding of the batched Convolver (one process per GPU).

* ``channel_shard`` / ``ChannelShardedBatch``
      independent convolver instances: contiguous ranges per rank, **no communication** on the
      data path (BASELINE configs 2, 3, 4).
* ``partition_shard`` / ``PartitionShardedConvolver``
      ONE convolver (``channels`` channels) whose long impulse response is split by partition
      range (BASELINE config 5): rank ``g`` owns the taps ``[p_lo*F, p_hi*F)`` and initialises its
      instances with ``b200conv_init_range(..., part_offset=p_lo)``; every rank feeds the same
      input and the partial output blocks are summed
        - ``reduce="fused"``: inside the kernel tails, all-to-all over NVLink peer memory
          (``b200conv_reduce_*``; CUDA IPC handles exchanged once through ``torch.distributed``);
          every rank ends with the same summed block, no collective call per block;
        - ``reduce="nccl"``: one ``all_reduce`` of ``channels * F`` floats per block (baseline).

``torch.distributed`` is plumbing here (handle exchange, the baseline collective); importing this
module does not import torch.
"""

HANDLE_BYTES = 64


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


def exchange_handles(mine, group=None, device=None):
    """All-gathers one fixed-size byte string per rank (the CUDA IPC handle of the exchange
    buffer) over ``torch.distributed``; returns the list in rank order.  ``device`` = torch
    device of the transport tensors (cuda for NCCL, None / cpu for gloo)."""
    import torch
    import torch.distributed as dist
    if len(mine) != HANDLE_BYTES:
        raise ValueError("expected a %d-byte handle" % HANDLE_BYTES)
    world = dist.get_world_size(group)
    t = torch.tensor(list(mine), dtype=torch.uint8, device=device)
    every = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(every, t, group=group)
    return [bytes(e.cpu().tolist()) for e in every]


class ChannelShardedBatch:
    """The local share of ``instances`` independent convolvers: rank ``rank`` of ``world`` owns
    instances ``[lo, hi)``.  Nothing is ever exchanged."""

    def __init__(self, pkg, instances, world, rank, device):
        self.lo, self.hi = channel_shard(instances, world, rank)
        self.batch = pkg.ConvolverBatch(max(1, self.hi - self.lo), device)

    def init(self, global_idx, data, rank_fft, phase=0.0):
        if not (self.lo <= global_idx < self.hi):
            return True                     # another rank's instance
        return self.batch.init(global_idx - self.lo, data, rank_fft, phase)

    def owns(self, global_idx):
        return self.lo <= global_idx < self.hi

    def close(self):
        self.batch.close()


class PartitionShardedConvolver:
    """One ``channels``-channel convolver with a long IR, split by partition range over the ranks
    of a ``torch.distributed`` process group (one process per GPU)."""

    def __init__(self, pkg, channels, rank_fft, device, reduce="fused", group=None):
        import torch.distributed as dist
        if reduce not in ("fused", "nccl"):
            raise ValueError("reduce must be 'fused' or 'nccl'")
        self._dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.channels, self.rank_fft, self.device = channels, rank_fft, device
        self.reduce = reduce if self.world > 1 else "none"
        self.batch = pkg.ConvolverBatch(channels, device)
        self.frame = 1 << (max(8, min(16, rank_fft)) - 1)
        self.connected = False

    def init(self, irs, phase=0.0):
        """``irs``: one full-length impulse response per channel (every rank passes the same)."""
        shards, offsets = [], []
        for ir in irs:
            p_lo, p_hi, t_lo, t_hi = partition_shard(len(ir), self.frame, self.world, self.rank)
            # a rank with an empty range still takes part in the exchange: a 1-tap zero IR
            shards.append(ir[t_lo:t_hi] if t_hi > t_lo else ir[:1] * 0)
            offsets.append(p_lo)
        import time
        t0 = time.perf_counter()
        ok = self.batch.init_many(list(range(len(irs))), shards, self.rank_fft, [phase] * len(irs), offsets)
        self.init_ms = (time.perf_counter() - t0) * 1e3         # IR upload + partition transforms of this rank's share
        self.connect_ms = 0.0
        if ok and self.reduce == "fused":
            import torch
            t0 = time.perf_counter()
            mine = self.batch.reduce_prepare(self.rank, self.world)
            self.batch.reduce_connect(exchange_handles(mine, self.group, torch.device("cuda", self.device)))
            self._dist.barrier(self.group)      # nobody stores into a buffer its owner has not opened yet
            self.connected = True
            self.connect_ms = (time.perf_counter() - t0) * 1e3  # IPC handle exchange, peer mappings, barrier
        return ok

    def process_device(self, dst, src, count, stream=None):
        """``src`` / ``dst``: torch CUDA tensors ``[channels][>= count]`` (row-contiguous); every
        rank passes the same input and receives the summed output."""
        import torch
        st = stream.cuda_stream if stream is not None else None
        self.batch.process_device(dst.data_ptr(), src.data_ptr(), src.stride(0), count, st,
                                  dst_stride=dst.stride(0))
        if self.reduce == "nccl":
            ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(
                torch.cuda.ExternalStream(self.batch.stream()))
            with ctx:
                blk = dst[:, :count].contiguous()
                self._dist.all_reduce(blk, op=self._dist.ReduceOp.SUM, group=self.group)
                dst[:, :count] = blk

    def timed_out(self):
        return self.batch.reduce_timed_out()

    def close(self):
        if self.connected:
            self.batch.sync()
            self._dist.barrier(self.group)      # peers may still be storing into this rank's buffer
            self.batch.reduce_disconnect()
            self.connected = False
        self.batch.close()
