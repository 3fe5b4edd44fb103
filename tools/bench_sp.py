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
