"""Per-block time of channel shares of cfg 3 / partition shares of cfg 5 on one GPU, with the early input
transform allowed (default) and refused ("early_src" = 0): the launches between 150 and 450 MB per block."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import __graft_entry__ as ge, synth
pkg = ge.load()
shapes = [("cfg3 share of 2 GPUs: 32 ch", 32, 480000), ("cfg3 share of 4 GPUs: 16 ch", 16, 480000),
          ("cfg5 share of 4 GPUs: 1407 partitions", 1, 1407 * 1024), ("cfg5 share of 2 GPUs: 2813 partitions", 1, 2813 * 1024),
          ("cfg3: 64 ch", 64, 480000)]
for name, n, taps in shapes:
    for early in (1, 0):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("early_src", early)
        ir = synth.decaying_ir(0, taps)
        b.init_many(list(range(n)), [ir] * n, 11, [0.0] * n)
        frames, block = 512, 1024
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
            for _ in range(2): run()
            e1.record(st)
        torch.cuda.synchronize()
        print(name, "| early_src", early, "| us/block %.2f" % (e0.elapsed_time(e1) * 1e3 / (2 * frames)), flush=True)
        b.close()
