#!/usr/bin/env python
"""BASELINE config 5 experiments: ONE 8-channel convolver with a 120 s IR (5.76 M taps, 5625
partitions of 1024 taps) split by partition range across N GPUs (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port P tools/bench_cfg5.py [--blocks 2000] [--reduce fused|nccl] [--stages S]
        [--splits K] [--stream own|caller] [--early 0|1|2] [--taps T]

Prints one JSON line on rank 0: output samples/s, per-block time (max over ranks, CUDA events).
bench.py carries the driver-visible version of this shape (key "cfg5_split"); this tool is for A/B runs.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import __graft_entry__ as ge
import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taps", type=int, default=5760000)
    ap.add_argument("--channels", type=int, default=8)
    ap.add_argument("--blocks", type=int, default=2000)
    ap.add_argument("--reduce", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--splits", type=int, default=0)
    ap.add_argument("--stream", default="own", choices=["own", "caller"])
    ap.add_argument("--early", type=int, default=1)
    ap.add_argument("--pdl", type=int, default=1)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.load()
    from lsp_dsp_units_b200 import sharding

    R, F, C = 11, 1024, args.channels
    irs = [synth.decaying_ir(c, args.taps) for c in range(C)]
    conv = sharding.PartitionShardedConvolver(pkg, C, R, local, reduce=args.reduce)
    assert conv.init(irs)
    b = conv.batch
    b.set_option("mac_stages", args.stages)
    b.set_option("mac_splits", args.splits)
    b.set_option("early_src", args.early)
    b.set_option("pdl", args.pdl)
    p_lo, p_hi, _, _ = sharding.partition_shard(args.taps, F, world, rank)

    g = torch.Generator(device="cuda").manual_seed(1234)           # same input on every rank
    src = torch.rand((C, 64 * F), generator=g, device="cuda") * 2 - 1
    dst = torch.empty_like(src)
    own = args.stream == "own"
    stream = torch.cuda.ExternalStream(b.stream()) if own else torch.cuda.Stream()
    sp, dp = src.data_ptr(), dst.data_ptr()
    torch.cuda.synchronize()

    def call(t):
        o = 4 * (t % 64) * F
        if args.reduce == "fused" or world == 1:
            b.process_device(dp + o, sp + o, 64 * F, F, None if own else stream.cuda_stream)
        else:
            conv.process_device(dst[:, o // 4:o // 4 + F], src[:, o // 4:o // 4 + F], F, None if own else stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for t in range(256):
            call(t)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for t in range(args.blocks):
            call(t)
        e1.record(stream)
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        us = float(ms.item()) * 1e3 / args.blocks
        print(json.dumps({"n_gpus": world, "reduce": args.reduce if world > 1 else "none", "stages": args.stages,
                          "splits": args.splits, "stream": args.stream, "early_src": args.early, "pdl": args.pdl,
                          "partitions_this_gpu": p_hi - p_lo, "us_per_block": us,
                          "samples_per_s": C * F / (us * 1e-6), "timed_out": bool(conv.timed_out())}), flush=True)
    conv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
