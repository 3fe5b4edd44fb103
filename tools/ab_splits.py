"""Per-block time of the 64-channel 10 s shape at ranks 14..16 against the k_mac bin tile ("mac_tile")
and with / without the MAC launched one block ahead ("chain_ahead"); set SPLITS=1 in the environment
to sweep the number of partition splits ("mac_splits") instead."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import __graft_entry__ as ge, synth
pkg = ge.load()
sweep_splits = os.environ.get("SPLITS") == "1"
for rank in (16, 15, 14):
    cases = [(0, 1, s) for s in (0, 1, 2, 3, 4)] if sweep_splits else [(tile, ahead, 0) for tile in (0, 512) for ahead in (1, 0)]
    for tile, ahead, splits in cases:
        n, taps = 64, 480000
        F = 1 << (rank - 1)
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("mac_splits", splits)
        b.set_option("mac_tile", tile)
        b.set_option("chain_ahead", ahead)
        ir = synth.decaying_ir(0, taps)
        b.init_many(list(range(n)), [ir] * n, rank, [0.0] * n)
        frames = 24
        src = torch.rand((n, frames * F), device="cuda") * 2 - 1
        dst = torch.empty_like(src)
        st = torch.cuda.ExternalStream(b.stream())
        torch.cuda.synchronize()
        def run():
            for i in range(frames):
                b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, frames * F, F, None)
        run(); b.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(3): run()
            e1.record(st)
        torch.cuda.synchronize()
        print("rank", rank, "tile", tile or 1024, "chain_ahead", ahead, "splits", splits or "auto",
              "us/block %.2f" % (e0.elapsed_time(e1) * 1e3 / (3 * frames)), flush=True)
        b.set_option("mac_tile", 0)
        b.close()
