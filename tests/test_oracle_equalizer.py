"""CPU checks of the Equalizer (FIR / FFT mode) oracle -- scope-table row f2.

The oracle (oracle/equalizer_oracle.c) restates Equalizer.cpp:474-518 over the restated fastconv
primitives.  It is pinned by identity where one exists (no hand-over: direct convolution delayed by
fir_size; the reference's own utest pins peak index == latency, src/test/utest/filters/
equalizer.cpp:34-84) and cross-checked against an independent float64 numpy model elsewhere."""
import numpy as np
import pytest

import synth
from equalizer_model import ModelEqualizer, band_kernel
from oracle.bindings import CpuEqualizer, direct_convolve


@pytest.mark.parametrize("fir_rank,step", [(7, 31), (8, 256), (9, 1000), (10, 77), (12, 4096)])
def test_oracle_is_delayed_direct_convolution(fir_rank, step):
    F = 1 << fir_rank
    k = synth.decaying_ir(fir_rank, F)
    x = synth.noise(fir_rank, 6 * F + 13)
    e = CpuEqualizer(fir_rank)
    e.set_kernel(k)
    y = e.run(x, step)
    want = np.concatenate([np.zeros(F), direct_convolve(x, k)])[:len(x)]
    assert np.max(np.abs(y - want)) <= 1e-5 * np.max(np.abs(want))


@pytest.mark.parametrize("fir_rank", [7, 10, 13])
def test_oracle_latency_like_reference_utest(fir_rank):
    """equalizer.cpp:34-84: feed a unit impulse, the peak of the response sits at get_latency()
    = nFirSize + nFirSize / 2 (Equalizer.cpp:347) for the linear-phase kernels reconfigure builds."""
    F = 1 << fir_rank
    e = CpuEqualizer(fir_rank)
    e.set_kernel(band_kernel(fir_rank, 0.01, 1.01))             # high-pass, like the utest's filter
    x = np.zeros(4 * F, dtype=np.float32)
    x[0] = 1.0
    y = e.process(x)
    assert int(np.argmax(np.abs(y))) == F + F // 2


@pytest.mark.parametrize("fir_rank,step,swap_at", [(7, 31, 200), (8, 256, 256), (9, 100, 1300), (10, 1024, 3000)])
def test_oracle_crossfade_matches_model(fir_rank, step, swap_at):
    F = 1 << fir_rank
    k0 = band_kernel(fir_rank, 0.0, 0.3)
    k1 = band_kernel(fir_rank, 0.2, 0.7, 2.0)
    k2 = synth.decaying_ir(3, F)
    x = synth.noise(40 + fir_rank, 8 * F)
    e, m = CpuEqualizer(fir_rank), ModelEqualizer(fir_rank)
    for q in (e, m):
        q.set_kernel(k0)
    ys, ym = [], []
    pos = 0
    events = {swap_at: (k1, True), swap_at + 3 * F: (k2, True), swap_at + 4 * F + 5: (k0, False)}
    marks = sorted(set(list(range(0, len(x), step)) + list(events) + [len(x)]))
    for a, b in zip(marks[:-1], marks[1:]):
        if a in events:
            for q in (e, m):
                q.set_kernel(events[a][0], events[a][1])
        ys.append(e.process(x[a:b]))
        ym.append(m.process(x[a:b]))
    ys, ym = np.concatenate(ys), np.concatenate(ym)
    assert np.max(np.abs(ys - ym)) <= 1e-5 * np.max(np.abs(ym))
    # the hand-over really changed the output (not a vacuous comparison)
    plain = CpuEqualizer(fir_rank)
    plain.set_kernel(k0)
    assert np.max(np.abs(plain.process(x) - ys)) > 0.05 * np.max(np.abs(ym))


def test_oracle_clear_and_inplace():
    fir_rank = 8
    F = 1 << fir_rank
    k = band_kernel(fir_rank, 0.1, 0.5)
    x = synth.noise(5, 3 * F + 40)
    e = CpuEqualizer(fir_rank)
    e.set_kernel(k)
    first = e.process(x)
    e.clear()
    again = e.process(x)
    assert np.array_equal(first, again)
