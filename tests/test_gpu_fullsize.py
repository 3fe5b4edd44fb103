"""GPU parity at the FULL sizes of BASELINE.json configs 2-5 ("parity is checked over
the whole run"), through the C ABI.

* cfg 2 : 2 x 192 000 taps, phases 0 / 0.5, 256-sample calls, ranks 9 and 8, 2 s of noise plus a
          flush of L - 1 zeros, both instances against the CPU oracle (reference scheduler).
* cfg 3 : 64 x 480 000 taps, rank 11, 600 consecutive 1024-sample calls (the 477-slot ring wraps
          at full size), channels 0 and 37 against the oracle.
* cfg 4 : 4096 x 48 000 taps, 4 s of noise in 8192-sample calls (the multi-frame path), 16
          instances against the oracle.
* cfg 5 : 8 x 5 760 000 taps (5 625 partitions), noise over MORE than the whole IR length, two
          channels against float64 truth (FFT convolution) -- accumulation over thousands of
          partitions.

Tolerance: max |gpu - ref| <= 1e-5 of the reference peak (north_star).  The oracle runs on the
host cores of the GPU box: sizes are chosen so that each case finishes in seconds.
"""
import numpy as np
import pytest

import synth
from oracle.bindings import CpuConvolver, direct_convolve

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    p = ge.load()
    p.lib()
    return p


def rel_err(got, want):
    want = np.asarray(want, np.float64)
    return float(np.max(np.abs(np.asarray(got, np.float64) - want)) / np.max(np.abs(want)))


def oracle_run(ir, src, rank, phase, step):
    c = CpuConvolver("oracle")
    assert c.init(ir, rank, phase)
    return c.run(src, step)


@pytest.mark.parametrize("rank", [9, 8])
def test_config2_full_size_stereo_192000_taps(pkg, rank):
    taps, step, fs = 192000, 256, 48000
    n_in = 2 * fs
    total = ((n_in + taps - 1 + step - 1) // step) * step       # input + flush of L - 1 zeros
    phases = [0.0, 0.5]
    irs = [synth.decaying_ir(20 + c, taps) for c in range(2)]
    src = np.zeros((2, total), np.float32)
    for c in range(2):
        src[c, :n_in] = synth.noise(20 + c, n_in)
    b = pkg.ConvolverBatch(2, 0)
    for c in range(2):
        assert b.init(c, irs[c], rank, phases[c])
    F = 1 << (rank - 1)
    assert b.state(0)["frame_off"] == 0 and b.state(1)["frame_off"] == F // 2
    assert b.state(0)["bins"] == (taps + F - 1) // F
    out = np.empty_like(src)
    for i in range(0, total, step):
        out[:, i:i + step] = b.process(src[:, i:i + step])
    b.close()
    for c in range(2):
        want = oracle_run(irs[c], src[c], rank, phases[c], step)
        assert rel_err(out[c], want) <= TOL
        # and the oracle itself is the convolution (identity, float64)
        assert rel_err(out[c], direct_convolve(src[c], irs[c], total)) <= TOL


def test_config3_full_size_ring_wraps_against_oracle(pkg):
    torch = pytest.importorskip("torch")
    n, taps, rank, F, blocks = 64, 480000, 11, 1024, 600
    check = [0, 37]
    irs = {c: synth.decaying_ir(c, taps) for c in check}
    filler = synth.decaying_ir(99, taps)
    b = pkg.ConvolverBatch(n, 0)
    for c in range(n):
        assert b.init(c, irs.get(c, filler), rank, 0.0)
    st = b.state(0)
    assert st["bins"] == 469 and st["partitions"] == 470
    g = torch.Generator(device="cuda").manual_seed(3)
    src = torch.rand((n, blocks * F), generator=g, device="cuda") * 2 - 1
    for c in check:
        src[c] = torch.from_numpy(synth.noise(c, blocks * F)).cuda()
    dst = torch.empty_like(src)
    for i in range(blocks):                 # one launch per 1024-sample block, back to back
        b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, blocks * F, F)
    b.sync()
    assert b.state(0)["frames"] == blocks
    for c in check:
        x = src[c].cpu().numpy()
        want = oracle_run(irs[c], x, rank, 0.0, F)
        assert rel_err(dst[c].cpu().numpy(), want) <= TOL
    b.close()


def test_config4_full_size_4096_instances(pkg):
    torch = pytest.importorskip("torch")
    n, taps, rank, F = 4096, 48000, 11, 1024
    call, calls = 8192, 24                  # 4.1 s of audio per instance, offline-sized calls
    total = call * calls
    check = [0, 1, 255, 256, 1000, 1023, 1024, 2047, 2048, 2500, 3000, 3333, 3500, 4000, 4094, 4095]
    irs = {c: synth.decaying_ir(c, taps) for c in check}
    filler = [synth.decaying_ir(5000 + k, taps) for k in range(4)]
    b = pkg.ConvolverBatch(n, 0)
    for c in range(n):
        assert b.init(c, irs.get(c, filler[c % 4]), rank, 0.0)
    g = torch.Generator(device="cuda").manual_seed(4)
    src = torch.rand((n, total), generator=g, device="cuda") * 2 - 1
    for c in check:
        src[c] = torch.from_numpy(synth.noise(c, total)).cuda()
    dst = torch.empty_like(src)
    for i in range(calls):
        b.process_device(dst.data_ptr() + 4 * i * call, src.data_ptr() + 4 * i * call, total, call)
    b.sync()
    for c in check:
        x = src[c].cpu().numpy()
        want = oracle_run(irs[c], x, rank, 0.0, F)
        assert rel_err(dst[c].cpu().numpy(), want) <= TOL
    # no instance was skipped: every output row carries signal
    assert bool((dst.abs().amax(dim=1) > 0).all().item())
    b.close()


def test_config5_full_size_5625_partitions_float64_truth(pkg):
    torch = pytest.importorskip("torch")
    n, taps, rank, F = 8, 5760000, 11, 1024
    bins = (taps + F - 1) // F
    blocks = bins + 200                     # noise over more than the whole IR length
    check = [0, 5]
    irs = [synth.decaying_ir(100 + c, taps) for c in range(n)]
    b = pkg.ConvolverBatch(n, 0)
    for c in range(n):
        assert b.init(c, irs[c], rank, 0.0)
    assert b.state(0)["bins"] == 5625
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.rand((n, blocks * F), generator=g, device="cuda") * 2 - 1
    dst = torch.empty_like(src)
    for i in range(blocks):
        b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, blocks * F, F)
    b.sync()
    from scipy.signal import fftconvolve
    for c in check:
        x = src[c].cpu().numpy().astype(np.float64)
        want = fftconvolve(x, irs[c].astype(np.float64))[:blocks * F]
        got = dst[c].cpu().numpy()
        assert rel_err(got, want) <= TOL
        # the tail of the run (every partition active) on its own
        assert rel_err(got[-64 * F:], want[-64 * F:]) <= TOL * np.max(np.abs(want)) / np.max(np.abs(want[-64 * F:]))
    b.close()
