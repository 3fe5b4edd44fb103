import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, numpy as np
import __graft_entry__ as ge, synth
pkg = ge.load()
shapes = [("cfg1", 1, 65536, 11, 1024, (0.0,)), ("cfg2", 2, 192000, 9, 256, (0.0, 0.5)), ("8ch x 703 part", 8, 703 * 1024, 11, 1024, (0.0,)),
          ("8ch strong cfg3", 8, 480000, 11, 1024, (0.0,)), ("16ch", 16, 480000, 11, 1024, (0.0,))]
for name, n, taps, rank, block, phases in shapes:
    for stages in (2, 3, 4):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("mac_stages", stages)
        irs = [synth.decaying_ir(c, taps) for c in range(min(n, 2))]
        b.init_many(list(range(n)), [irs[c % len(irs)] for c in range(n)], rank, [phases[c % len(phases)] for c in range(n)])
        frames = 512
        src = torch.rand((n, frames * block), device="cuda") * 2 - 1
        dst = torch.empty_like(src)
        st = torch.cuda.ExternalStream(b.stream())
        torch.cuda.synchronize()
        def run():
            for i in range(frames):
                b.process_device(dst.data_ptr() + 4 * i * block, src.data_ptr() + 4 * i * block, frames * block, block, None)
        run(); b.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(4): run()
            e1.record(st)
        torch.cuda.synchronize()
        print(name, "stages", stages, "us/call %.2f" % (e0.elapsed_time(e1) * 1e3 / (4 * frames)), flush=True)
        b.close()
