#!/usr/bin/env python
"""BASELINE config 5: ONE 8-channel convolver with a 120 s IR (5.76 M taps, 5625 partitions of
1024 taps) split by partition range across N GPUs, partial output blocks summed with an NCCL
all-reduce.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port P tools/bench_cfg5.py [--blocks 2000] [--check]

Every rank feeds the same input block, streams its own partition range (b200conv_init_range,
part_offset = p_lo) and produces a partial block; block t's all-reduce (8 x 1024 floats, 32 KiB,
latency-bound) runs on a side stream under block t+1's partition stream (`--depth` blocks in
flight).  Prints one JSON line on rank 0: output samples/s (8 channels), per-block time, and with
--check the error against float64 truth on a short prefix.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import __graft_entry__ as ge
import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taps", type=int, default=5760000)
    ap.add_argument("--channels", type=int, default=8)
    ap.add_argument("--blocks", type=int, default=2000)
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--reduce-every", type=int, default=1,
                    help="blocks per all-reduce: 1 = every 1024-sample block gets its own 32 KiB reduce "
                         "(real-time use); K > 1 amortises the per-collective host cost (offline use)")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--fused-reduce", action="store_true",
                    help="sum the partial blocks inside the launch tails over NVLink peer memory "
                         "(b200conv_reduce_*) instead of one NCCL all-reduce per block")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.load()
    import lsp_dsp_units_b200.sharding as sharding

    R, F, C = 11, 1024, args.channels
    p_lo, p_hi, t_lo, t_hi = sharding.partition_shard(args.taps, F, world, rank)
    irs = [synth.decaying_ir(c, args.taps) for c in range(C)]
    b = pkg.ConvolverBatch(C, local)
    for c in range(C):
        assert b.init(c, irs[c][t_lo:t_hi], R, 0.0, part_offset=p_lo)

    nblk = args.blocks
    g = torch.Generator(device="cuda").manual_seed(1234)           # same input on every rank
    src = torch.rand((C, 64 * F), generator=g, device="cuda") * 2 - 1
    K = max(1, args.reduce_every)
    ring = [torch.empty((C, K * F), device="cuda") for _ in range(args.depth)]
    done = [None] * args.depth
    out_keep = torch.empty((C, 64 * F), device="cuda") if args.check else None
    compute, comm = torch.cuda.Stream(), torch.cuda.Stream()

    def run(blocks, keep):
        for g0 in range(0, blocks, K):
            k = (g0 // K) % args.depth
            if done[k] is not None:
                compute.wait_event(done[k])                       # the slot's previous reduce has finished
            with torch.cuda.stream(compute):
                for u in range(K):                                # K consecutive 1024-sample process calls
                    i = (g0 + u) % 64
                    b.process_device(ring[k].data_ptr() + 4 * u * F, src.data_ptr() + 4 * i * F,
                                     64 * F, F, compute.cuda_stream, dst_stride=K * F)
                ready = torch.cuda.Event()
                ready.record(compute)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                if world > 1:
                    dist.all_reduce(ring[k], op=dist.ReduceOp.SUM)
                if keep:
                    for u in range(K):
                        t = g0 + u
                        if t < 64:
                            out_keep[:, t * F:(t + 1) * F].copy_(ring[k][:, u * F:(u + 1) * F])
                ev = torch.cuda.Event()
                ev.record(comm)
                done[k] = ev

    fused = args.fused_reduce and world > 1
    if fused:
        mine = torch.tensor(list(b.reduce_prepare(rank, world)), dtype=torch.uint8, device="cuda")
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        b.reduce_connect([bytes(t.cpu().tolist()) for t in every])
        dist.barrier()
        fdst = torch.zeros((C, 64 * F), device="cuda")

        def run_fused(blocks, keep):
            with torch.cuda.stream(compute):
                for t in range(blocks):
                    i = t % 64
                    b.process_device(fdst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, 64 * F, F,
                                     compute.cuda_stream)
                if keep:
                    compute.synchronize()
                    out_keep.copy_(fdst)
        run_ = run_fused
    else:
        run_ = run

    run_(64, args.check)                                          # warm-up (and the checked prefix)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(compute)
    run_(nblk, False)
    last = compute if fused else comm
    last.synchronize()
    e1.record(last)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    timed_out = b.reduce_timed_out() if fused else False

    err = None
    if args.check and rank == 0:
        from oracle.bindings import direct_convolve
        x = src[0].cpu().numpy()
        want = direct_convolve(x, irs[0], 64 * F)
        got = out_keep[0].cpu().numpy()
        err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
    if rank == 0:
        bins = (args.taps + F - 1) // F
        rate = C * F * nblk / (float(ms.item()) * 1e-3)
        print(json.dumps({
            "config": "cfg5: %d ch x %d-tap IR, rank 11, 1024-sample blocks, partition range split over %d GPU(s), "
                      "%d-byte partial output blocks summed per block" % (C, args.taps, world, C * F * 4),
            "n_gpus": world, "partitions_total": bins, "partitions_per_gpu": p_hi - p_lo,
            "samples_per_s": rate, "us_per_block": float(ms.item()) * 1e3 / nblk,
            "realtime_factor": rate / (C * 48000.0), "blocks_in_flight": args.depth * K, "blocks_per_allreduce": K,
            "max_err_vs_float64_of_peak": err,
            "reduce": ("fused in the launch tails over NVLink peer memory" if fused else
                       ("NCCL all-reduce from the host" if world > 1 else "none")),
            "peer_wait_timed_out": bool(timed_out)}), flush=True)
    if fused:
        b.reduce_disconnect()
    b.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
