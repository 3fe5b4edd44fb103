#!/usr/bin/env python
"""Developer tool: the anatomy of ONE isolated block of BASELINE config 5 split over N GPUs (fused
all-to-all reduce).  Needs the timing build (tools/ab/libb200conv_timing.so, -DB200CONV_TIMING):

    B200CONV_LIB=tools/ab/libb200conv_timing.so python -m torch.distributed.run --nproc-per-node N ... tools/cfg5_timeline.py

Per-CTA timestamps of the last launch on rank 0: start, end of the partition stream, ticket,
(finishers:) start / end of the inverse transform, end of the send, end of the launch."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import __graft_entry__ as ge
import synth

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = ge.load()
from lsp_dsp_units_b200 import sharding
C, F, taps = 8, 1024, 5760000
conv = sharding.PartitionShardedConvolver(pkg, C, 11, local, reduce="fused")
assert conv.init([synth.decaying_ir(c, taps) for c in range(C)])
b = conv.batch
src = torch.rand((C, F), device="cuda"); dst = torch.empty_like(src)
torch.cuda.synchronize()
for it in range(30):
    if world > 1:
        dist.barrier()
    b.process_device(dst.data_ptr(), src.data_ptr(), F, F); b.sync()
lib = pkg.lib()
ncta = 256
buf = (ctypes.c_ulonglong * (ncta * 8))()
lib.b200conv_debug_frame_times.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
assert lib.b200conv_debug_frame_times(buf, ncta * 8) == 0
t = np.array(buf[:], dtype=np.float64).reshape(ncta, 8)
t0 = t[:, 0].min()
rel = (t - t0) / 1e3
fin = t[:, 4] >= t0                       # finishers of THIS launch (stale stamps are older)
if rank == 0:
    print("world %d: CTA start max %.1f us; stream end median %.1f max %.1f; ticket max %.1f" % (
        world, rel[:, 0].max(), np.median(rel[:, 1]), rel[:, 1].max(), rel[:, 2].max()))
    if world > 1:
        f = rel[fin]
        print("  finishers (%d): inverse start %.1f..%.1f, inverse end %.1f..%.1f, send end %.1f..%.1f, launch end %.1f..%.1f" % (
            fin.sum(), f[:, 4].min(), f[:, 4].max(), f[:, 5].min(), f[:, 5].max(), f[:, 6].min(), f[:, 6].max(), f[:, 3].min(), f[:, 3].max()))
    else:
        last = rel[:, 3] > 0
        print("  finishers: launch end %.1f..%.1f" % (rel[last, 3].min(), rel[last, 3].max()))
conv.close()
if world > 1:
    dist.destroy_process_group()
