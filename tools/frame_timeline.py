#!/usr/bin/env python
"""Developer tool: where does an ISOLATED k_frame launch spend its time?  Needs a library built with
-DB200CONV_TIMING (tools/frame_timeline.py builds its own copy under /tmp).  Prints, for one
isolated launch of the config-3 grid: start spread, end of streaming per CTA, tail."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as ge
import synth

lib_path = "/tmp/libb200conv_timing.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                       "-shared", "-DB200CONV_TIMING", "-I", os.path.join(ROOT, "include"), "-o", lib_path,
                       os.path.join(ROOT, "lsp-dsp-units_b200", "csrc", "engine.cu")])
pkg = ge.load()
pkg.LIB_PATH = lib_path
pkg._lib = None
lib = pkg.lib()
n, taps, F = 64, 480000, 1024
b = pkg.ConvolverBatch(n, 0)
irs = [synth.decaying_ir(c, taps) for c in range(4)]
for c in range(n):
    b.init(c, irs[c % 4], 11, 0.0)
for bias in (6, 12):
    b.set_option("fft_bias", bias)
    src = torch.rand((n, F), device="cuda"); dst = torch.empty_like(src)
    for _ in range(20):
        b.process_device(dst.data_ptr(), src.data_ptr(), F, F); b.sync()
    buf = (ctypes.c_ulonglong * (512 * 8))()
    lib.b200conv_debug_frame_times.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    assert lib.b200conv_debug_frame_times(buf, 512 * 8) == 0
    t = np.array(buf[:], dtype=np.float64).reshape(512, 8)[:, :4]
    t0 = t[:, 0].min()
    start, stream_end, ticket, tail_end = [(t[:, k] - t0) / 1e3 for k in range(4)]
    fft = np.arange(512) % 8 == 0
    last = t[:, 3] > t0
    print("bias %d: CTA start   max %.1f us" % (bias, start.max()))
    print("  stream end (non-FFT CTAs) min %.1f  median %.1f  p90 %.1f  max %.1f us" % (
        stream_end[~fft].min(), np.median(stream_end[~fft]), np.percentile(stream_end[~fft], 90), stream_end[~fft].max()))
    print("  stream end (FFT CTAs)     min %.1f  median %.1f  max %.1f us" % (stream_end[fft].min(), np.median(stream_end[fft]), stream_end[fft].max()))
    print("  ticket     max %.1f us;  tail end (last CTAs) median %.1f  max %.1f us;  tail duration median %.1f us" % (
        ticket.max(), np.median(tail_end[last]), tail_end[last].max(), np.median((tail_end - ticket)[last])))
