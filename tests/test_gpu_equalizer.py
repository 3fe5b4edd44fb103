"""GPU parity of the Equalizer FIR / FFT data path (b200conv_eq_*, scope-table row f2) through
the C ABI, against oracle/equalizer_oracle.c, the float64 model and the delayed-direct-convolution
identity.  Tolerance: 1e-5 of peak (the convolver's bar)."""
import ctypes

import numpy as np
import pytest

import synth
from equalizer_model import ModelEqualizer, band_kernel
from oracle.bindings import CpuEqualizer, direct_convolve

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL = 1e-5


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    p = ge.load()
    p.lib()
    return p


def _run(eq, x, step):
    out = np.empty_like(x)
    for i in range(0, x.shape[1], step):
        out[:, i:i + step] = eq.process(x[:, i:i + step])
    return out


@pytest.mark.parametrize("fir_rank,step", [(7, 31), (7, 128), (8, 256), (8, 1000), (9, 77), (10, 1024),
                                           (11, 5000), (12, 4096), (13, 3001), (14, 16384), (15, 50000)])
def test_delayed_direct_convolution_and_oracle(pkg, fir_rank, step):
    F = 1 << fir_rank
    n = 3
    total = (5 if fir_rank <= 12 else 3) * F + 13
    ks = [synth.decaying_ir(10 * fir_rank + c, F) for c in range(n)]
    x = np.stack([synth.noise(10 * fir_rank + c, total) for c in range(n)])
    eq = pkg.EqualizerBatch(n, fir_rank, device=0)
    assert eq.fir_size == F and eq.latency() == F
    for c in range(n):
        eq.set_kernel(c, ks[c])
    y = _run(eq, x, step)
    eq.close()
    for c in range(n):
        want = np.concatenate([np.zeros(F), direct_convolve(x[c], ks[c])])[:total]
        assert np.max(np.abs(y[c] - want)) <= TOL * np.max(np.abs(want))
    ref = CpuEqualizer(fir_rank)
    ref.set_kernel(ks[0])
    yo = ref.run(x[0], step)
    assert np.max(np.abs(y[0] - yo)) <= TOL * np.max(np.abs(yo))


@pytest.mark.parametrize("fir_rank", [7, 10, 13])
def test_latency_like_reference_utest(pkg, fir_rank):
    """src/test/utest/filters/equalizer.cpp:34-84: impulse in, peak of the response at
    get_latency() = nFirSize + nFirSize / 2 (FFT_RANK 13 in the reference's test)."""
    F = 1 << fir_rank
    eq = pkg.EqualizerBatch(1, fir_rank, device=0)
    eq.set_kernel(0, band_kernel(fir_rank, 0.01, 1.01))
    x = np.zeros((1, 4 * F), dtype=np.float32)
    x[0, 0] = 1.0
    y = eq.process(x)
    latency = eq.latency()
    eq.close()
    assert latency == F
    assert int(np.argmax(np.abs(y[0]))) == latency + F // 2


@pytest.mark.parametrize("fir_rank,step,swap_at", [(7, 31, 200), (8, 256, 256), (9, 100, 1300),
                                                   (10, 1024, 3000), (12, 999, 9000), (14, 16384, 20000)])
def test_smooth_handover_matches_oracle(pkg, fir_rank, step, swap_at):
    """Kernel replaced smoothly mid-stream on SOME instances (cross-fade of Equalizer.cpp:486-501),
    replaced at once on another, untouched on the last; then a second hand-over and the
    `replace while a cross-fade is pending` corner."""
    F = 1 << fir_rank
    n = 4
    total = 8 * F
    k0 = [band_kernel(fir_rank, 0.0, 0.2 + 0.1 * c) for c in range(n)]
    k1 = band_kernel(fir_rank, 0.2, 0.7, 2.0)
    k2 = synth.decaying_ir(3, F)
    x = np.stack([synth.noise(70 + c, total) for c in range(n)])

    eq = pkg.EqualizerBatch(n, fir_rank, device=0)
    refs = [CpuEqualizer(fir_rank) for _ in range(n)]
    models = [ModelEqualizer(fir_rank) for _ in range(n)]

    def set_kernel(c, k, smooth):
        eq.set_kernel(c, k, smooth)
        refs[c].set_kernel(k, smooth)
        models[c].set_kernel(k, smooth)

    for c in range(n):
        set_kernel(c, k0[c], False)
    events = {
        swap_at: [(0, k1, True), (1, k1, True), (2, k1, False)],
        swap_at + 3 * F: [(0, k2, True)],
        swap_at + 4 * F + 5: [(1, k2, True), (1, k0[1], False)],      # replace while a cross-fade is pending
    }
    marks = sorted(set(list(range(0, total, step)) + list(events) + [total]))
    y = np.empty_like(x)
    yo = np.empty_like(x)
    ym = np.empty(x.shape)
    for a, b in zip(marks[:-1], marks[1:]):
        for ev in events.get(a, []):
            set_kernel(*ev)
        y[:, a:b] = eq.process(x[:, a:b])
        for c in range(n):
            yo[c, a:b] = refs[c].process(x[c, a:b])
            ym[c, a:b] = models[c].process(x[c, a:b])
    eq.close()
    for c in range(n):
        peak = np.max(np.abs(ym[c]))
        assert np.max(np.abs(y[c] - yo[c])) <= TOL * peak, c
        assert np.max(np.abs(y[c] - ym[c])) <= TOL * peak, c


def test_clear_pointers_inplace_and_strides(pkg):
    fir_rank, n = 8, 5
    F = 1 << fir_rank
    ks = [band_kernel(fir_rank, 0.05 * c, 0.4 + 0.1 * c) for c in range(n)]
    x = np.stack([synth.noise(90 + c, 3 * F + 40) for c in range(n)])
    eq = pkg.EqualizerBatch(n, fir_rank, device=0)
    for c in range(n):
        eq.set_kernel(c, ks[c])
    first = eq.process(x)
    eq.clear()
    # pointer-per-instance API, in place, odd call sizes
    bufs = [x[c].copy() for c in range(n)]
    pos = 0
    for step in (1, 30, 255, 256, 257, 9999):
        m = min(step, x.shape[1] - pos)
        if m == 0:
            break
        views = [b[pos:pos + m] for b in bufs]
        eq.process_pointers(views, views, m)
        pos += m
    assert pos == x.shape[1]
    assert np.array_equal(np.stack(bufs), first)
    # planar matrix wider than the call
    eq.clear()
    wide_in = np.zeros((n, x.shape[1] + 37), dtype=np.float32)
    wide_in[:, :x.shape[1]] = x
    wide_out = np.full_like(wide_in, 7.0)
    lib = pkg.lib()
    rc = lib.b200conv_eq_process_planar(eq._h, wide_out.ctypes.data, wide_in.ctypes.data, wide_in.shape[1],
                                        x.shape[1])
    assert rc == 0
    assert np.array_equal(wide_out[:, :x.shape[1]], first)
    assert np.all(wide_out[:, x.shape[1]:] == 7.0)
    eq.close()


def test_device_api_on_a_caller_stream(pkg):
    fir_rank, n = 10, 64
    F = 1 << fir_rank
    ks = [band_kernel(fir_rank, 0.0, 0.02 * (c + 1)) for c in range(n)]
    x = np.stack([synth.noise(200 + c, 6 * F) for c in range(n)])
    eq = pkg.EqualizerBatch(n, fir_rank, device=0)
    for c in range(n):
        eq.set_kernel(c, ks[c])
    host = eq.process(x)
    eq.clear()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        dx = torch.from_numpy(x).cuda()
        dy = torch.empty_like(dx)
        for i in range(0, x.shape[1], F):
            eq.process_device(dy.data_ptr() + 4 * i, dx.shape[1], dx.data_ptr() + 4 * i, dx.shape[1], F,
                              s.cuda_stream)
        got = dy.cpu().numpy()          # ordered after the equalizer's work on the caller's stream
    eq.close()
    assert np.array_equal(got, host)
    want = np.concatenate([np.zeros(F), direct_convolve(x[5], ks[5])])[:x.shape[1]]
    assert np.max(np.abs(got[5] - want)) <= TOL * np.max(np.abs(want))


def test_bad_arguments(pkg):
    lib = pkg.lib()
    h = ctypes.c_void_p()
    assert lib.b200conv_eq_create(ctypes.byref(h), 0, 0, 8) != 0
    assert lib.b200conv_eq_create(ctypes.byref(h), 0, 1, 6) != 0
    assert lib.b200conv_eq_create(ctypes.byref(h), 0, 1, 16) != 0
    assert lib.b200conv_eq_create(ctypes.byref(h), 99, 1, 8) != 0
    assert b"device" in lib.b200conv_last_error()
    assert lib.b200conv_eq_create(ctypes.byref(h), 0, 2, 8) == 0
    k = np.zeros(256, dtype=np.float32)
    assert lib.b200conv_eq_set_kernel(h, 2, k.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 0) != 0
    assert lib.b200conv_eq_set_kernel(h, 0, None, 0) != 0
    assert lib.b200conv_eq_process_planar(h, None, None, 4, 4) != 0
    assert lib.b200conv_eq_process_planar(h, None, None, 0, 0) == 0          # nothing to do
    assert lib.b200conv_eq_latency(h) == 256
    assert lib.b200conv_eq_instances(h) == 2
    lib.b200conv_eq_free(h)
    lib.b200conv_eq_free(None)
    assert lib.b200conv_eq_process_planar(None, None, None, 4, 4) != 0
