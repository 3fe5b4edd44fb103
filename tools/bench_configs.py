#!/usr/bin/env python
"""Secondary measurements: the five BASELINE.json configs on ONE GPU (bench.py stays the headline,
config 3).  Per config: device-resident throughput of back-to-back process calls (CUDA events),
its share of the HBM roofline (16*bins+24 bytes per output sample per instance), and
the host-visible latency of one synchronous b200conv_process_planar call on pinned buffers
(median / p99 wall time) -- the number a real-time caller sees.

    python tools/bench_configs.py [--configs 1,2,3,4,5] > profiles/rN_configs.jsonl
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import __graft_entry__ as ge
import synth

#        name                                   instances taps      rank block phases        note
CONFIGS = {
    1: ("cfg1 mono 65536 taps, 1024 blocks",          1,   65536,   11, 1024, (0.0,),        ""),
    2: ("cfg2 stereo 4 s IR, 256 blocks, rank 9",     2,   192000,  9,  256,  (0.0, 0.5),    "phases 0 / 0.5: second instance takes the partial-frame path"),
    3: ("cfg3 64 ch x 10 s IR, 1024 blocks",          64,  480000,  11, 1024, (0.0,),        ""),
    4: ("cfg4 4096 mono x 1 s IR, 1024 blocks",       4096, 48000,  11, 1024, (0.0,),        "one frame per call"),
    6: ("cfg4 4096 mono x 1 s IR, 8192-sample calls", 4096, 48000,  11, 8192, (0.0,),        "offline render: 8 frames per call share one pass over the IR spectra (k_mac_multi<8>)"),
    7: ("cfg3 64 ch x 10 s IR, 8192-sample calls",    64,  480000,  11, 8192, (0.0,),        "offline render of config 3: 8 frames per call"),
    8: ("64 ch x 10 s IR, rank 9, 256 blocks",        64,  480000,  9,  256,  (0.0,),        "low-latency rank, one k_frame<9> launch per block"),
    9: ("64 ch x 10 s IR, rank 13, 4096 blocks",      64,  480000,  13, 4096, (0.0,),        "ranks 12..16: k_fwd -> k_mac -> k_inv per block (not fused yet)"),
    10: ("64 ch x 10 s IR, rank 16, 32768 blocks",    64,  480000,  16, 32768, (0.0,),       "largest rank"),
    11: ("4096 mono x 1024-tap IR, 8192-sample calls",  4096, 1024,    11, 8192, (0.0,),        "one partition: the transforms alone (k_fwd + k_inv over 32768 frames per call)"),
    12: ("1 x 200000 taps, rank 16, 31-sample calls",   1,   200000,  16, 31,   (0.0,),        "calls inside a frame: store + direct-form head in one launch"),
    13: ("64 ch x 10 s IR, rank 14, 8192 blocks",     64,  480000,  14, 8192, (0.0,),        ""),
    14: ("64 ch x 10 s IR, rank 15, 16384 blocks",    64,  480000,  15, 16384, (0.0,),       ""),
    15: ("8 x 60000 taps, rank 10, 77-sample calls",  8,   60000,   10, 77,   (0.0, 0.37),   "utest-like unaligned calls: one launch per step"),
    5: ("cfg5 8 ch x 120 s IR on ONE GPU",            8,   5760000, 11, 1024, (0.0,),        "all 5625 partitions on one GPU; the 8-way split is 1/8 of this per GPU + a 32 KiB all-reduce"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,5,4,6,7")
    ap.add_argument("--seconds", type=float, default=0.4)
    ap.add_argument("--multi", type=int, default=8)
    ap.add_argument("--no-host", action="store_true")
    ap.add_argument("--eager", type=int, default=1)
    ap.add_argument("--caller-stream", action="store_true", help="enqueue on a torch stream instead of the batch's own")
    args = ap.parse_args()
    pkg = ge.load()
    peak = 6546.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    for cid in [int(c) for c in args.configs.split(",")]:
        name, n, taps, rank, block, phases, note = CONFIGS[cid]
        F = 1 << (rank - 1)
        bins = (taps + F - 1) // F
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("multi_frame", args.multi)
        b.set_option("eager", args.eager)
        irs = [synth.decaying_ir(c, taps) for c in range(min(n, 4))]
        for c in range(n):
            assert b.init(c, irs[c % len(irs)], rank, phases[c % len(phases)])
        bytes_per_sample = 16 * bins + 24           # per-frame model: one pass over the IR per frame

        # device-resident throughput
        frames = max(8, min(512, int(2e8 // (n * block))))
        src = torch.rand((n, frames * block), device="cuda") * 2 - 1
        dst = torch.empty_like(src)
        # the batch's own stream (there the engine may run a block's input transform early)
        st = torch.cuda.ExternalStream(b.stream()) if not args.caller_stream else torch.cuda.Stream()
        torch.cuda.synchronize()
        def run():
            for i in range(frames):
                b.process_device(dst.data_ptr() + 4 * i * block, src.data_ptr() + 4 * i * block,
                                 frames * block, block, st.cuda_stream if args.caller_stream else None)
        with torch.cuda.stream(st):
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps, t0 = 0, time.time()
            e0.record(st)
            while reps < 3 or time.time() - t0 < args.seconds:
                run()
                reps += 1
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        rate = reps * frames * block * n / (ms * 1e-3)
        us_per_call = ms * 1e3 / (reps * frames)

        # host-visible latency of one synchronous call on pinned buffers
        hs = torch.zeros((n, block)).pin_memory()
        hs.copy_(src[:, :block].cpu())
        hd = torch.zeros((n, block)).pin_memory()
        a, o = hs.numpy(), hd.numpy()
        for _ in range(20 if not args.no_host else 1):
            b.process(a, o)
        lat = []
        t_end = time.time() + args.seconds
        while (len(lat) < 200 or time.time() < t_end) and not (args.no_host and len(lat) >= 3):
            t0 = time.perf_counter()
            b.process(a, o)
            lat.append(time.perf_counter() - t0)
            if len(lat) >= 5000:
                break
        lat = np.sort(np.array(lat)) * 1e6
        # the same call when the host comes back once per audio block (paced: 1 ms between calls,
        # the block period is 5-21 ms): with the eager pending MAC only the head is left to do
        paced = []
        if not args.no_host:
            for _ in range(60):
                t_wait = time.perf_counter() + 1e-3
                while time.perf_counter() < t_wait:
                    pass
                t0 = time.perf_counter()
                b.process(a, o)
                paced.append(time.perf_counter() - t0)
        paced = np.sort(np.array(paced if paced else [0.0])) * 1e6
        # ... and with the pending MAC forced to start under the delivering launch (early_pend = 2;
        # the default, 1, does that only for a caller that keeps the GPU busy): it competes with
        # the latency-critical launch
        paced_lat = []
        if not args.no_host:
            b.set_option("early_pend", 2)
            for _ in range(60):
                t_wait = time.perf_counter() + 1e-3
                while time.perf_counter() < t_wait:
                    pass
                t0 = time.perf_counter()
                b.process(a, o)
                paced_lat.append(time.perf_counter() - t0)
            b.set_option("early_pend", 1)
        paced_lat = np.sort(np.array(paced_lat if paced_lat else [0.0])) * 1e6
        line = {
            "config": name, "instances": n, "taps": taps, "rank": rank, "block": block, "partitions": bins,
            "device_samples_per_s": rate, "device_us_per_call": us_per_call,
            "hbm_roofline_samples_per_s": peak * 1e9 / bytes_per_sample,
            "share_of_hbm_roofline": rate * bytes_per_sample / (peak * 1e9),
            "realtime_factor": rate / (n * 48000.0),
            "host_call_us_median": float(lat[len(lat) // 2]), "host_call_us_p99": float(lat[int(len(lat) * 0.99)]),
            "host_samples_per_s": n * block / (float(lat[len(lat) // 2]) * 1e-6),
            "host_call_us_paced_median": float(paced[len(paced) // 2]),
            "host_call_us_paced_median_early_pend_2": float(paced_lat[len(paced_lat) // 2]),
            "block_duration_us_at_48k": block / 48000.0 * 1e6, "note": note,
        }
        print(json.dumps(line), flush=True)
        b.close()
        del src, dst
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
