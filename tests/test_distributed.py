"""N > 1 host logic on CPU: torch.distributed (gloo, world size 2) around the sharding plans.

No GPU here, so the per-rank arithmetic is the numpy model of the engine (tests/engine_model.py,
index-for-index what the kernels do, including ``part_offset`` partition-range shards); what is
under test is the plan (who owns what) and the exchange step (sum of partial output blocks /
no exchange at all for independent channels)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist            # noqa: E402
import torch.multiprocessing as mp          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _setup(rank, world, port):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    ge.load()
    import lsp_dsp_units_b200.sharding as sharding
    return sharding


def _partition_worker(rank, world, port, result):
    sharding = _setup(rank, world, port)
    import engine_model as em
    import synth
    R, F, L, blocks = 9, 256, 5000, 12
    ir, x = synth.decaying_ir(2, L), synth.noise(2, blocks * F)
    p_lo, p_hi, t_lo, t_hi = sharding.partition_shard(L, F, world, rank)
    conv = em.ModelConvolver(ir[t_lo:t_hi], R, 0.0, part_offset=p_lo)
    out = np.empty(blocks * F, np.float32)
    for b in range(blocks):
        part = torch.from_numpy(conv.process(x[b * F:(b + 1) * F]).copy())
        dist.all_reduce(part, op=dist.ReduceOp.SUM)         # the one exchange step of config 5
        out[b * F:(b + 1) * F] = part.numpy()
    want = np.convolve(x.astype(np.float64), ir.astype(np.float64))[:blocks * F]
    err = float(np.max(np.abs(out - want)) / np.max(np.abs(want)))
    if rank == 0:
        result.put(("partition", err))
    dist.destroy_process_group()


def _channel_worker(rank, world, port, result):
    sharding = _setup(rank, world, port)
    import engine_model as em
    import synth
    R, F, L, blocks, channels = 8, 128, 700, 6, 5
    lo, hi = sharding.channel_shard(channels, world, rank)
    errs = []
    for c in range(lo, hi):                                  # no collective on the data path
        ir, x = synth.decaying_ir(c, L), synth.noise(c, blocks * F)
        conv = em.ModelConvolver(ir, R, 0.0)
        out = np.concatenate([conv.process(x[b * F:(b + 1) * F]) for b in range(blocks)])
        want = np.convolve(x.astype(np.float64), ir.astype(np.float64))[:blocks * F]
        errs.append(float(np.max(np.abs(out - want)) / np.max(np.abs(want))))
    # only the bookkeeping is gathered: every channel was processed exactly once
    owned = torch.zeros(channels, dtype=torch.int32)
    owned[lo:hi] = 1
    dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    worst = torch.tensor([max(errs) if errs else 0.0], dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        result.put(("channel", bool((owned == 1).all()), float(worst.item())))
    dist.destroy_process_group()


def _equalizer_worker(rank, world, port, result):
    """Row f2: equalizers are independent instances -- contiguous ranges per rank, no collective on
    the data path; one rank hands a kernel over smoothly mid-stream, the others must not notice."""
    sharding = _setup(rank, world, port)
    import synth
    from equalizer_model import ModelEqualizer, band_kernel
    fir_rank, instances, blocks = 7, 7, 6
    F = 1 << fir_rank
    lo, hi = sharding.channel_shard(instances, world, rank)
    errs = []
    for c in range(lo, hi):
        k0, k1 = band_kernel(fir_rank, 0.0, 0.1 * (c + 1)), band_kernel(fir_rank, 0.3, 0.9)
        x = synth.noise(c, blocks * F)
        eq = ModelEqualizer(fir_rank)
        eq.set_kernel(k0)
        first = eq.process(x[:2 * F + 17])
        swapped = (c == 0)
        if swapped:
            eq.set_kernel(k1, smooth=True)
        out = np.concatenate([first, eq.process(x[2 * F + 17:])])
        if not swapped:
            want = np.concatenate([np.zeros(F), np.convolve(x.astype(np.float64), k0.astype(np.float64))])[:blocks * F]
            errs.append(float(np.max(np.abs(out - want)) / np.max(np.abs(want))))
        else:
            # after the cross-fade block the output is the new kernel's delayed convolution
            want = np.concatenate([np.zeros(F), np.convolve(x.astype(np.float64), k1.astype(np.float64))])[:blocks * F]
            tail = slice(5 * F, blocks * F)
            errs.append(float(np.max(np.abs(out[tail] - want[tail])) / np.max(np.abs(want))))
    owned = torch.zeros(instances, dtype=torch.int32)
    owned[lo:hi] = 1
    dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    worst = torch.tensor([max(errs) if errs else 0.0], dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        result.put(("equalizer", bool((owned == 1).all()), float(worst.item())))
    dist.destroy_process_group()


def _handles_worker(rank, world, port, result):
    """The transport of the fused reduce's set-up: every rank contributes a 64-byte handle and gets
    all of them back in rank order (gloo, CPU tensors); and the all-to-all sum itself, modelled on
    the host: every rank adds the partial blocks of all ranks in RANK ORDER, so every rank ends
    with the same bits (what the kernel tails do over NVLink)."""
    sharding = _setup(rank, world, port)
    import engine_model as em
    import synth
    mine = bytes([(17 * rank + i) % 251 for i in range(sharding.HANDLE_BYTES)])
    every = sharding.exchange_handles(mine, device=None)
    ok = (len(every) == world) and all(every[r] == bytes([(17 * r + i) % 251 for i in range(sharding.HANDLE_BYTES)])
                                       for r in range(world))
    try:
        sharding.exchange_handles(b"short")
        ok = False
    except ValueError:
        pass
    R, F, L, blocks = 9, 256, 7000, 10
    ir, x = synth.decaying_ir(4, L), synth.noise(4, blocks * F)
    p_lo, p_hi, t_lo, t_hi = sharding.partition_shard(L, F, world, rank)
    conv = em.ModelConvolver(ir[t_lo:t_hi], R, 0.0, part_offset=p_lo)
    out = np.empty(blocks * F, np.float32)
    for b in range(blocks):
        part = torch.from_numpy(conv.process(x[b * F:(b + 1) * F]).astype(np.float32).copy())
        slots = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(slots, part)                        # every rank's block lands in every rank's slots
        acc = torch.zeros_like(part)
        for r in range(world):                              # rank order: bit-identical everywhere
            acc += slots[r]
        out[b * F:(b + 1) * F] = acc.numpy()
    want = np.convolve(x.astype(np.float64), ir.astype(np.float64))[:blocks * F]
    err = float(np.max(np.abs(out - want)) / np.max(np.abs(want)))
    mine_out = torch.from_numpy(out.copy())
    ref = mine_out.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([int(torch.equal(ref, mine_out))])
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        result.put(("handles", ok, err, bool(same.item())))
    dist.destroy_process_group()


def _run(worker, port):
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, result)) for r in range(2)]
    for p in procs:
        p.start()
    out = result.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_partition_range_shards_allreduce_gloo_world2():
    tag, err = _run(_partition_worker, 29611)
    assert tag == "partition" and err <= 1e-5


def test_channel_shards_no_collective_gloo_world2():
    tag, complete, err = _run(_channel_worker, 29612)
    assert tag == "channel" and complete and err <= 1e-5


def test_equalizer_instances_shard_by_instance_gloo_world2():
    tag, complete, err = _run(_equalizer_worker, 29613)
    assert tag == "equalizer" and complete and err <= 1e-9


def test_handle_exchange_and_all_to_all_sum_gloo_world2():
    tag, ok, err, same = _run(_handles_worker, 29614)
    assert tag == "handles" and ok and same and err <= 1e-5


def test_shard_plans_cover_everything():
    import __graft_entry__ as ge
    ge.load()
    import lsp_dsp_units_b200.sharding as sharding
    for n, world in ((64, 1), (64, 8), (5, 2), (3, 8)):
        spans = [sharding.channel_shard(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    for taps, F, world in ((5760000, 1024, 8), (5000, 256, 2), (100, 128, 4)):
        spans = [sharding.partition_shard(taps, F, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][3] == taps
        assert all(a[1] == b[0] and a[3] == b[2] for a, b in zip(spans, spans[1:]))
    # config 5: 5625 partitions over 8 GPUs -> 703 or 704 each
    sizes = {s[1] - s[0] for s in [sharding.partition_shard(5760000, 1024, 8, r) for r in range(8)]}
    assert sizes == {703, 704}
