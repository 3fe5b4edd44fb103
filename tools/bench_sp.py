#!/usr/bin/env python
"""Throughput of the batched SpectralProcessor (row f4): `instances` processors with a gain table
bound, device-resident calls of `call` samples.  One JSON line per rank."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import __graft_entry__ as ge

pkg = ge.load()
peak = 6546.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for rank, n, call in ((8, 4096, 4096), (10, 4096, 8192), (12, 2048, 16384), (14, 512, 65536)):
    N = 1 << rank
    sp = pkg.SpectralProcessorBatch(n, rank, device=0)
    gain = np.linspace(0.5, 1.0, N).astype(np.float32)
    for c in range(0, n, 7):
        sp.bind_gain(c, gain)                   # a 7th of the instances unbound: both branches
    for c in range(n):
        if c % 7:
            sp.bind_gain(c, gain)
    src = torch.rand((n, call), device="cuda") * 2 - 1
    dst = torch.empty_like(src)
    st = torch.cuda.ExternalStream(sp.stream())
    torch.cuda.synchronize()
    for _ in range(3):
        sp.process_device(dst.data_ptr(), call, src.data_ptr(), call, call)
    sp.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(reps):
            sp.process_device(dst.data_ptr(), call, src.data_ptr(), call, call)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rate = n * call / (ms * 1e-3)
    # per sample: src 4 + dst 4 + input buffer 4 w + 8 r (two windows read it) + 4 shift, output buffer 8 r + 8 w ~ 40 B
    print(json.dumps({"what": "SpectralProcessor batch", "rank": rank, "instances": n, "call": call,
                      "ms_per_call": ms, "samples_per_s": rate, "frames_per_s": rate / (N / 2),
                      "approx_share_of_hbm_at_40B_per_sample": rate * 40 / (peak * 1e9)}), flush=True)
    sp.close()

# the batched SpectralSplitter as a three-way FFTCrossover: input samples per second (every input
# sample yields one output sample per band)
for rank, n, call in ((10, 2048, 8192), (12, 1024, 16384)):
    N = 1 << rank
    ss = pkg.SpectralSplitterBatch(n, rank, 3, device=0)
    k = np.minimum(np.arange(N), N - np.arange(N)) / (N / 2)
    lo = (1.0 / (1.0 + (k / 0.05) ** 4)).astype(np.float32)
    mid = ((1.0 - lo) / (1.0 + (k / 0.4) ** 4)).astype(np.float32)
    hi = (1.0 - lo - mid).astype(np.float32)
    for c in range(n):
        for h, g in enumerate((lo, mid, hi)):
            ss.bind_gain(c, h, g)
    src = torch.rand((n, call), device="cuda") * 2 - 1
    dst = torch.empty((3, n, call), device="cuda")
    st = torch.cuda.ExternalStream(ss.stream())
    torch.cuda.synchronize()
    for _ in range(3):
        ss.process_device(dst.data_ptr(), n * call, call, src.data_ptr(), call, call)
    ss.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(reps):
            ss.process_device(dst.data_ptr(), n * call, call, src.data_ptr(), call, call)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"what": "SpectralSplitter batch, 3 bands with real gain curves (FFTCrossover)", "rank": rank, "instances": n,
                      "call": call, "ms_per_call": ms, "input_samples_per_s": n * call / (ms * 1e-3),
                      "band_samples_per_s": 3 * n * call / (ms * 1e-3)}), flush=True)
    ss.close()
