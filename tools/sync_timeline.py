#!/usr/bin/env python
"""Developer tool: one synchronous host call (page-locked buffers, eager pending MAC) of config 1 and of
config 3 -- where does the head-only k_frame launch spend its time, and what does the host see?
Needs tools/ab/libb200conv_timing.so (tools/gen_timeline.py --build)."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as ge
import synth
pkg = ge.load()
pkg.LIB_PATH = os.path.join(ROOT, "tools", "ab", "libb200conv_timing.so")
pkg._lib = None
lib = pkg.lib()
lib.b200conv_debug_frame_times.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
for name, n, taps in (("cfg1", 1, 65536), ("cfg3", 64, 480000)):
    F = 1024
    b = pkg.ConvolverBatch(n, 0)
    ir = synth.decaying_ir(0, taps)
    b.init_many(list(range(n)), [ir] * n, 11, [0.0] * n)
    hs = torch.rand((n, F)).pin_memory(); hd = torch.empty((n, F)).pin_memory()
    hsn, hdn = hs.numpy(), hd.numpy()
    lat = []
    for i in range(300):
        t0 = time.perf_counter()
        b.process(hsn, hdn)
        lat.append((time.perf_counter() - t0) * 1e6)
        time.sleep(0.0005)                      # let the pending MAC finish: the paced case
    b.sync()
    buf = (ctypes.c_ulonglong * (64 * 8))()
    assert lib.b200conv_debug_frame_times(buf, 64 * 8) == 0
    t = np.array(buf[:], dtype=np.float64).reshape(64, 8)[:n]
    t0 = t[:, 0].min()
    r = (t - t0) / 1e3
    print("%s: host-visible call median %.1f us (p10 %.1f); head-only launch: %d CTAs, start spread %.1f, "
          "last stage consumed median %.1f max %.1f, ticket max %.1f, end median %.1f max %.1f us" % (
              name, np.median(lat[50:]), np.percentile(lat[50:], 10), n, r[:, 0].max(), np.median(r[:, 1]), r[:, 1].max(),
              r[:, 2].max(), np.median(r[:, 3]), r[:, 3].max()), flush=True)
    b.close()

# config 2 through synchronous host calls: the job-list launch reads and writes the page-locked block
n, taps, rank, F = 2, 192000, 9, 256
b = pkg.ConvolverBatch(n, 0)
for c in range(n):
    assert b.init(c, synth.decaying_ir(c, taps), rank, (0.0, 0.5)[c])
hs = torch.rand((n, F)).pin_memory(); hd = torch.empty((n, F)).pin_memory()
hsn, hdn = hs.numpy(), hd.numpy()
lat = []
for i in range(300):
    t0 = time.perf_counter()
    b.process(hsn, hdn)
    lat.append((time.perf_counter() - t0) * 1e6)
    time.sleep(0.0005)
b.sync()
buf = (ctypes.c_ulonglong * (64 * 8))()
assert lib.b200conv_debug_frame_times(buf, 64 * 8) == 0
t = np.array(buf[:], dtype=np.float64).reshape(64, 8)
t0 = t[:, 0].min()
r = (t - t0) / 1e3
print("cfg2: host-visible call median %.1f us (p10 %.1f)" % (np.median(lat[50:]), np.percentile(lat[50:], 10)))
for c in (0, 32):
    print("   cta %2d  start %5.1f  P1 %5.1f  stream %5.1f  fwd %5.1f  last-stage %5.1f  ticket %5.1f  inverse %5.1f  end %5.1f" % (
        c, r[c, 0], r[c, 4], r[c, 5], r[c, 6], r[c, 1], r[c, 2], r[c, 7], r[c, 3]))
oth = [c for c in range(64) if c not in (0, 32)]
print("   other CTAs: stream end median %.1f max %.1f; ticket median %.1f max %.1f us" % (
    np.median(r[oth, 5]), r[oth, 5].max(), np.median(r[oth, 2]), r[oth, 2].max()))
# the same through the pointer-table call on pageable arrays (what bench_configs' host column measures)
x = np.random.rand(n, F).astype(np.float32); y = np.empty_like(x)
lat = []
for i in range(300):
    t0 = time.perf_counter()
    b.process(x, y)
    lat.append((time.perf_counter() - t0) * 1e6)
print("cfg2 pageable arrays back to back: median %.1f us" % np.median(lat[50:]))
b.close()
