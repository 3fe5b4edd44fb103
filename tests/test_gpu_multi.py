"""Multi-GPU paths on real devices (needs >= 2 GPUs; skipped otherwise): one process per GPU,
NCCL.  (a) independent channels sharded by instance, no collective; (b) BASELINE config 5 in
miniature: one long IR split by partition range, partial output blocks summed with one NCCL
all-reduce per block."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist            # noqa: E402
import torch.multiprocessing as mp          # noqa: E402

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, result):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import __graft_entry__ as ge
    import synth
    from oracle.bindings import direct_convolve
    pkg = ge.load()
    import lsp_dsp_units_b200.sharding as sharding

    # (b) partition-range sharding: 4 channels x 300 partitions, rank 11
    R, F, ch, blocks = 11, 1024, 4, 40
    L = 300 * F - 17
    irs = [synth.decaying_ir(c, L) for c in range(ch)]
    x = np.stack([synth.noise(c, blocks * F) for c in range(ch)])
    p_lo, p_hi, t_lo, t_hi = sharding.partition_shard(L, F, world, rank)
    b = pkg.ConvolverBatch(ch, rank)
    for c in range(ch):
        assert b.init(c, irs[c][t_lo:t_hi], R, 0.0, part_offset=p_lo)
    src = torch.from_numpy(x).cuda()
    dst = torch.empty_like(src)
    st = torch.cuda.Stream()            # a real stream: NULL would mean "the batch's own stream"
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for i in range(blocks):
            b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, blocks * F, F, st.cuda_stream)
            blk = dst[:, i * F:(i + 1) * F].contiguous()
            dist.all_reduce(blk, op=dist.ReduceOp.SUM)          # 4 x 1024 floats over NVLink
            dst[:, i * F:(i + 1) * F] = blk
    torch.cuda.synchronize()
    out = dst.cpu().numpy()
    err_b = max(float(np.max(np.abs(out[c] - direct_convolve(x[c], irs[c], blocks * F)))
                      / np.max(np.abs(direct_convolve(x[c], irs[c], blocks * F)))) for c in range(ch))
    b.close()

    # (a) channel sharding: 6 channels, no collective on the data path
    lo, hi = sharding.channel_shard(6, world, rank)
    b = pkg.ConvolverBatch(hi - lo, rank)
    irs = [synth.decaying_ir(10 + c, 20000) for c in range(lo, hi)]
    x = np.stack([synth.noise(10 + c, 16 * F) for c in range(lo, hi)])
    for i, ir in enumerate(irs):
        assert b.init(i, ir, R, 0.0)
    out = np.concatenate([b.process(x[:, i * F:(i + 1) * F].copy()) for i in range(16)], axis=1)
    err_a = max(float(np.max(np.abs(out[i] - direct_convolve(x[i], irs[i], 16 * F)))
                      / np.max(np.abs(direct_convolve(x[i], irs[i], 16 * F)))) for i in range(hi - lo))
    b.close()

    worst = torch.tensor([err_a, err_b], dtype=torch.float64, device="cuda")
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        result.put(tuple(float(v) for v in worst.cpu()))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_channel_and_partition_sharding_nccl():
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29641, result)) for r in range(2)]
    for p in procs:
        p.start()
    err_a, err_b = result.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert err_a <= 1e-5 and err_b <= 1e-5


def _fused_worker(rank, world, port, result):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import __graft_entry__ as ge
    import synth
    from oracle.bindings import direct_convolve
    pkg = ge.load()
    import lsp_dsp_units_b200.sharding as sharding

    R, F, ch, blocks = 11, 1024, 4, 60
    L = 300 * F - 17
    irs = [synth.decaying_ir(c, L) for c in range(ch)]
    x = np.stack([synth.noise(c, blocks * F) for c in range(ch)])
    conv = sharding.PartitionShardedConvolver(pkg, ch, R, rank, reduce="fused")
    assert conv.init(irs)               # slices the IRs, exchanges the IPC handles, connects
    b = conv.batch

    src = torch.from_numpy(x).cuda()
    dst = torch.zeros_like(src)
    torch.cuda.synchronize()
    for i in range(blocks):                 # back to back: launches overlap, tails talk over NVLink
        b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, blocks * F, F)
    b.sync()
    timed_out = b.reduce_timed_out()
    dist.barrier()
    # EVERY rank ends with the summed block (all-to-all), bit-identical across ranks
    out = dst.cpu().numpy()
    err = 0.0
    for c in range(ch):
        want = direct_convolve(x[c], irs[c], blocks * F)
        err = max(err, float(np.max(np.abs(out[c] - want)) / np.max(np.abs(want))))
    ref0 = dst.clone()
    dist.broadcast(ref0, src=0)
    differs = float(not torch.equal(ref0, dst))
    flags = torch.tensor([float(timed_out), differs, err], device="cuda", dtype=torch.float64)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    conv.close()
    if rank == 0:
        result.put((float(flags[2].item()), float(flags[0].item()), float(flags[1].item())))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_fused_nvlink_reduce_of_partition_shards():
    """BASELINE config 5 in miniature without a collective call: the launch tails exchange the partial
    output blocks over NVLink peer memory (b200conv_reduce_*, all-to-all); every rank ends with the
    same summed block."""
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, 29651, result)) for r in range(2)]
    for p in procs:
        p.start()
    err, timed_out, differs = result.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert timed_out == 0.0 and differs == 0.0
    assert err <= 1e-5
