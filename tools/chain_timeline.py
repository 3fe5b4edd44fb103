#!/usr/bin/env python
"""Developer tool: the three-kernel block of ranks 14..16 (k_fwd_half -> k_mac -> k_inv_half, chained with
programmatic serialisation) on a %globaltimer timeline: when do the CTAs of each kernel start and end?
Needs tools/ab/libb200conv_timing.so (tools/gen_timeline.py --build)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as ge
import synth
pkg = ge.load()
pkg.LIB_PATH = os.environ.get("B200CONV_LIB", os.path.join(ROOT, "tools", "ab", "libb200conv_timing.so"))
pkg._lib = None
lib = pkg.lib()
lib.b200conv_debug_frame_times.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
rank = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, taps = 64, 480000
F = 1 << (rank - 1)
b = pkg.ConvolverBatch(n, 0)
ir = synth.decaying_ir(0, taps)
b.init_many(list(range(n)), [ir] * n, rank, [0.0] * n)
blocks = 6
src = torch.rand((n, blocks * F), device="cuda") * 2 - 1
dst = torch.empty_like(src)
torch.cuda.synchronize()
for rep in range(2):
    for i in range(blocks):
        b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, blocks * F, F)
    b.sync()
buf = (ctypes.c_ulonglong * (8192 * 8))()
assert lib.b200conv_debug_frame_times(buf, 8192 * 8) == 0
t = np.array(buf[:], dtype=np.float64).reshape(8192, 8)
recent = t[:, 0].max() - 2e6                    # stamps of the last block only (ns)
mac = t[:4096]; mac = mac[mac[:, 0] > recent]
fwd = t[4096:4096 + 512]; fwd = fwd[fwd[:, 0] > recent]
inv = t[4096 + 512:4096 + 1024]; inv = inv[inv[:, 0] > recent]
t0 = fwd[:, 0].min()
def us(x): return (x - t0) / 1e3
print("rank %d, last block: times in us since the first transform CTA passed its wait" % rank)
print("  k_fwd_half: %d CTAs, start %.1f .. %.1f, end %.1f .. %.1f" % (len(fwd), us(fwd[:, 0]).min(), us(fwd[:, 0]).max(), us(fwd[:, 1]).min(), us(fwd[:, 1]).max()))
print("              passes start %.1f .. %.1f, passes end %.1f .. %.1f (median pass time %.1f us)" % (
    us(fwd[:, 2]).min(), us(fwd[:, 2]).max(), us(fwd[:, 3]).min(), us(fwd[:, 3]).max(), np.median(fwd[:, 3] - fwd[:, 2]) / 1e3))
fe = fwd[:, 1].max()
print("  k_mac: %d CTAs; started before the transform ended: %d; start p0/p10/p50/p90/p100 = %s" % (
    len(mac), int((mac[:, 0] < fe).sum()), " ".join("%.1f" % v for v in np.percentile(us(mac[:, 0]), [0, 10, 50, 90, 100]))))
print("         stream end p0/p50/p100 = %s; after wait p0/p50/p100 = %s" % (
    " ".join("%.1f" % v for v in np.percentile(us(mac[:, 1]), [0, 50, 100])), " ".join("%.1f" % v for v in np.percentile(us(mac[:, 2]), [0, 50, 100]))))
dur = (mac[:, 1] - mac[:, 0]) / 1e3
print("         per-CTA stream time p10/p50/p90 = %s us" % " ".join("%.1f" % v for v in np.percentile(dur, [10, 50, 90])))
print("  k_inv_half: %d CTAs, resident %.1f .. %.1f, past wait %.1f .. %.1f" % (len(inv), us(inv[:, 0]).min(), us(inv[:, 0]).max(), us(inv[:, 1]).min(), us(inv[:, 1]).max()))
print("              passes start %.1f .. %.1f, passes end %.1f .. %.1f (median pass time %.1f us)" % (
    us(inv[:, 2]).min(), us(inv[:, 2]).max(), us(inv[:, 3]).min(), us(inv[:, 3]).max(), np.median(inv[:, 3] - inv[:, 2]) / 1e3))
print("              end %.1f .. %.1f" % (us(inv[:, 4]).min(), us(inv[:, 4]).max()))
