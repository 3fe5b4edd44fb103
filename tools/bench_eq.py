#!/usr/bin/env python
"""Throughput of the Equalizer FIR / FFT data path (b200conv_eq_*, scope-table row f2) on ONE GPU.

Per (instances, fir_rank): device-resident output samples/s of block-aligned process_device calls
(CUDA events on the batch's stream), the share of the HBM roofline at 32 algorithmic bytes per
sample per instance (4 in + 4 out + 8 kernel spectrum + 8 vInBuffer store / load + 8 vOutBuffer
overlap tail load / store), and the CPU oracle (one thread, one instance) beside it.

    python tools/bench_eq.py > profiles/rN_equalizer.jsonl
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
from equalizer_model import band_kernel

BYTES_PER_SAMPLE = 32.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="4096:10,4096:12,1024:13,256:15,64:10,1:13")
    ap.add_argument("--seconds", type=float, default=0.3)
    ap.add_argument("--blocks-per-call", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    pkg = ge.load()
    peak = 6546.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    for case in args.cases.split(","):
        n, fir_rank = [int(v) for v in case.split(":")]
        F = 1 << fir_rank
        call = F * args.blocks_per_call
        eq = pkg.EqualizerBatch(n, fir_rank, device=0)
        ks = [band_kernel(fir_rank, 0.0, 0.1 * (c + 1)) for c in range(4)]
        for c in range(n):
            eq.set_kernel(c, ks[c % 4])
        calls = max(4, min(256, int(3e8 // (n * call))))
        src = torch.rand((n, calls * call), device="cuda") * 2 - 1
        dst = torch.empty_like(src)
        st = torch.cuda.ExternalStream(eq.stream())

        def run():
            for i in range(calls):
                eq.process_device(dst.data_ptr() + 4 * i * call, calls * call, src.data_ptr() + 4 * i * call,
                                  calls * call, call, None)
        torch.cuda.synchronize()
        run()
        eq.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps, t0 = 0, time.time()
        e0.record(st)
        while reps < 3 or time.time() - t0 < args.seconds:
            run()
            reps += 1
        e1.record(st)
        eq.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rate = reps * calls * call * n / (ms * 1e-3)

        cpu_rate = None
        if not args.no_cpu:
            from oracle.bindings import CpuEqualizer
            ref = CpuEqualizer(fir_rank)
            ref.set_kernel(ks[0])
            xin = synth.noise(1, F * max(4, (1 << 18) // F))
            ref.process(xin[:2 * F])
            t0 = time.perf_counter()
            done = 0
            while time.perf_counter() - t0 < 1.0:
                ref.process(xin)
                done += xin.size
            cpu_rate = done / (time.perf_counter() - t0)

        line = {
            "path": "Equalizer FIR/FFT data path (b200conv_eq_process_device)", "instances": n, "fir_rank": fir_rank,
            "fir_size": F, "samples_per_call": call, "device_samples_per_s": rate,
            "device_us_per_block": ms * 1e3 / (reps * calls * args.blocks_per_call),
            "hbm_roofline_samples_per_s": peak * 1e9 / BYTES_PER_SAMPLE,
            "share_of_hbm_roofline": rate * BYTES_PER_SAMPLE / (peak * 1e9),
            "realtime_factor_at_48k": rate / (n * 48000.0),
            "cpu_oracle_samples_per_s_one_thread": cpu_rate,
        }
        print(json.dumps(line), flush=True)
        eq.close()
        del src, dst
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
