#!/usr/bin/env python
"""Developer tool: where does ONE k_frame_gen launch (the job-list form: any call size / phase) spend
its time?  Uses a library built with -DB200CONV_TIMING (built under tools/ab/ when missing; that
directory is git-ignored but travels to the GPU box).  Stamps (%globaltimer, thread 0 of each CTA):
0 start, 4 after P1, 5 end of the main partition stream, 6 after the input transform, 1 after the
last stage, 2 ticket, 7 after the inverse transform, 3 end."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
lib_path = os.path.join(ROOT, "tools", "ab", "libb200conv_timing.so")
if "--build" in sys.argv or not os.path.exists(lib_path):
    os.makedirs(os.path.dirname(lib_path), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                           "-shared", "-DB200CONV_TIMING", "-I", os.path.join(ROOT, "include"), "-o", lib_path,
                           os.path.join(ROOT, "lsp-dsp-units_b200", "csrc", "engine.cu"), "-lcuda"])
    if "--build" in sys.argv:
        sys.exit(0)
import torch
import __graft_entry__ as ge
import synth
pkg = ge.load()
pkg.LIB_PATH = lib_path
pkg._lib = None
lib = pkg.lib()
lib.b200conv_debug_frame_times.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
SHAPES = [("cfg2", 2, 192000, 9, 256, (0.0, 0.5)), ("8 x 60000 rank 10, 77-sample calls", 8, 60000, 10, 77, (0.0, 0.37))]
for name, n, taps, rank, block, phases in SHAPES:
    b = pkg.ConvolverBatch(n, 0)
    for c in range(n):
        assert b.init(c, synth.decaying_ir(c, taps), rank, phases[c % len(phases)])
    calls = 64
    src = torch.rand((n, calls * block), device="cuda") * 2 - 1
    dst = torch.empty_like(src)
    torch.cuda.synchronize()
    for i in range(calls):
        b.process_device(dst.data_ptr() + 4 * i * block, src.data_ptr() + 4 * i * block, calls * block, block)
        b.sync()
        if i < calls - 4:
            continue
        ncta = 1024
        buf = (ctypes.c_ulonglong * (ncta * 8))()
        assert lib.b200conv_debug_frame_times(buf, ncta * 8) == 0
        t = np.array(buf[:], dtype=np.float64).reshape(ncta, 8)
        live = t[:, 0] > t[:, 0].max() - 1e6            # stamps of this launch
        t0 = t[live, 0].min()
        r = (t - t0) / 1e3
        print("%s call %d: %d CTAs, start spread %.1f us" % (name, i, live.sum(), r[live, 0].max()))
        for c in np.nonzero(live)[0]:
            row = r[c]
            fresh = t[c] >= t0
            def g(k): return ("%6.1f" % row[k]) if fresh[k] else "     -"
            if fresh[6] or fresh[7] or fresh[3]:
                print("   cta %3d  start %s  P1 %s  stream %s  fwd %s  last-stage %s  ticket %s  inverse %s  end %s" % (
                    c, g(0), g(4), g(5), g(6), g(1), g(2), g(7), g(3)))
        others = live & ~((t[:, 6] >= t0) | (t[:, 3] >= t0))
        if others.any():
            print("   other CTAs: stream end median %.1f max %.1f; ticket median %.1f max %.1f us" % (
                np.median(r[others, 5]), r[others, 5].max(), np.median(r[others, 2]), r[others, 2].max()))
    b.close()
