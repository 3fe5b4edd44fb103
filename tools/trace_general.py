#!/usr/bin/env python
"""A few dozen process calls of the latency-bound shapes (BASELINE config 1, config 2, unaligned
utest-like calls, rank 16 with 31-sample calls), for an ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file out.csv \\
        python tools/trace_general.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import __graft_entry__ as ge
import synth

pkg = ge.load()
SHAPES = [(1, 65536, 11, 1024, (0.0,)), (2, 192000, 9, 256, (0.0, 0.5)), (8, 60000, 10, 77, (0.0, 0.37)),
          (1, 200000, 16, 31, (0.0,))]
for n, taps, rank, block, phases in SHAPES:
    b = pkg.ConvolverBatch(n, 0)
    for c in range(n):
        assert b.init(c, synth.decaying_ir(c, taps), rank, phases[c % len(phases)])
    calls = 40
    src = torch.rand((n, calls * block), device="cuda") * 2 - 1
    dst = torch.empty_like(src)
    torch.cuda.synchronize()
    for i in range(calls):
        b.process_device(dst.data_ptr() + 4 * i * block, src.data_ptr() + 4 * i * block, calls * block, block)
    b.sync()
    b.close()
