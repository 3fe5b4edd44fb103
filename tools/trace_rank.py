#!/usr/bin/env python
"""A few blocks of the 64-channel x 10 s geometry at one rank (default 16), for ncu captures of
the big-rank transforms:  python tools/trace_rank.py [rank] [blocks]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import __graft_entry__ as ge
import synth

rank = int(sys.argv[1]) if len(sys.argv) > 1 else 16
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pkg = ge.load()
n, taps, F = 64, 480000, 1 << (rank - 1)
b = pkg.ConvolverBatch(n, 0)
irs = [synth.decaying_ir(c, taps) for c in range(2)]
assert b.init_many(list(range(n)), [irs[c % 2] for c in range(n)], rank)
src = torch.rand((n, blocks * F), device="cuda") * 2 - 1
dst = torch.empty_like(src)
torch.cuda.synchronize()
for i in range(blocks):
    b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, blocks * F, F)
b.sync()
b.close()
