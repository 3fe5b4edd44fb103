"""Scope-table row f4, second sibling: the reference's own lsp::dspu::SpectralSplitter (compiled
verbatim into oracle/_ref) pinned by an independent float64 model (tests/splitter_model.py) and by
the properties its callers rely on (FFTCrossover: bands whose gains add up to one add up to the
delayed input).  CPU only."""
import numpy as np
import pytest

import splitter_model
import synth
from oracle.bindings import CpuSpectralSplitter

pytestmark = pytest.mark.skipif(not CpuSpectralSplitter.available(), reason="oracle/_ref has not been built")


def tables(rank, seed):
    N = 1 << rank
    rng = np.random.Generator(np.random.PCG64(seed))
    gain = rng.uniform(0.2, 1.5, N).astype(np.float32)
    half = rng.uniform(-1, 1, N // 2 + 1) + 1j * rng.uniform(-1, 1, N // 2 + 1)
    half[0] = half[0].real
    half[-1] = half[-1].real
    H = np.concatenate([half, np.conj(half[-2:0:-1])]).astype(np.complex64)
    return gain, H


@pytest.mark.parametrize("rank,chunk,phase,step", [(8, 0, 0.0, 31), (9, 0, 0.5, 256), (10, 8, 0.37, 1000),
                                                   (12, 10, 1.0, 4096), (7, 5, 0.25, 77), (11, 3, 0.0, 500)])
def test_against_the_model(rank, chunk, phase, step):
    n = 6 * (1 << rank) + 123
    src = synth.noise(rank, n)
    gain, H = tables(rank, rank)
    ss = CpuSpectralSplitter(13, 4)
    ss.set_rank(rank)
    if chunk:
        ss.set_chunk_rank(chunk)
    ss.set_phase(phase)
    m = splitter_model.ModelSpectralSplitter(rank, 4, chunk, phase)
    assert ss.bind_gain(0, gain) == 0
    m.bind(0, lambda X: X * gain.astype(np.float64))
    assert ss.bind_complex(1, H) == 0
    m.bind(1, lambda X: X * H.astype(np.complex128))
    assert ss.bind_sink(3) == 0                 # handler 2 stays unbound
    m.bind(3, "copy")
    got = ss.run(src, step)
    want = m.process(src)
    assert ss.bindings() == 3
    cr = min(max(chunk, 5), rank) if chunk > 0 else rank
    assert ss.chunk_rank() == cr and ss.latency() == (1 << cr)
    assert np.all(got[2] == 0.0)
    for h in (0, 1, 3):
        assert np.max(np.abs(got[h] - want[h])) <= 2e-5 * max(1.0, np.max(np.abs(want[h]))), h


def test_bands_that_add_up_to_one_return_the_delayed_input():
    # what FFTCrossover builds on: per-band real gains (FFTCrossover.cpp:124-140) that sum to one
    rank, n = 10, 8000
    N = 1 << rank
    src = synth.noise(3, n)
    k = np.minimum(np.arange(N), N - np.arange(N)) / (N / 2)
    lo = (1.0 / (1.0 + (k / 0.1) ** 4)).astype(np.float32)
    ss = CpuSpectralSplitter(rank, 2)
    ss.bind_gain(0, lo)
    ss.bind_gain(1, (1.0 - lo).astype(np.float32))
    out = ss.run(src, 333)
    lat = ss.latency()
    assert lat == N
    assert np.max(np.abs(out[0, lat:] + out[1, lat:] - src[:n - lat])) <= 2e-5


def test_nothing_bound_is_a_no_op_and_bind_checks():
    ss = CpuSpectralSplitter(9, 2)
    assert ss.bindings() == 0
    out = ss.process(synth.noise(1, 700))
    assert np.all(out == 0.0)
    assert ss.bind_sink(5) != 0                 # STATUS_OVERFLOW
    assert ss.unbind(1) != 0                    # STATUS_NOT_BOUND
    assert ss.bind_sink(1) == 0 and ss.bindings() == 1


def test_crossover_curves_of_the_reference_split_the_signal():
    """FFTCrossover = SpectralSplitter + one real curve per band (FFTCrossover.cpp:124-140,458-480).
    Three bands built with the reference's own curve functions: LPF 300 Hz | HPF 300 Hz + LPF 3 kHz |
    HPF 3 kHz, -24 dB/oct.  hipass(f) + lopass(f) = 1 at every frequency, so the bands of a two-way
    split add up to the delayed input."""
    from oracle.bindings import crossover_band_curve
    rank, sr, n = 12, 48000, 30000
    N = 1 << rank
    lo = crossover_band_curve(rank, sr, lpf=(300.0, -24.0))
    mid = crossover_band_curve(rank, sr, hpf=(300.0, -24.0), lpf=(3000.0, -24.0))
    hi = crossover_band_curve(rank, sr, hpf=(3000.0, -24.0))
    lo2 = crossover_band_curve(rank, sr, lpf=(3000.0, -24.0))
    assert lo.shape == (N,) and np.all(lo >= 0) and np.all(hi >= 0) and abs(lo[0] - 1.0) < 1e-6 and hi[0] == 0.0
    assert np.max(np.abs(lo2[1:] + hi[1:] - 1.0)) < 1e-6            # a two-way split is complementary
    assert np.allclose(lo[1:N // 2], lo[N - 1:N // 2:-1])             # symmetric: a real impulse response
    src = synth.noise(77, n)
    ss = CpuSpectralSplitter(rank, 3)
    for h, curve in enumerate((lo2, hi, mid)):
        ss.bind_gain(h, curve)
    out = ss.run(src, 1024)
    lat = ss.latency()
    # bin 0: lopass 1 + hipass 0
    assert np.max(np.abs(out[0, lat:] + out[1, lat:] - src[:n - lat])) <= 3e-5
    assert np.max(np.abs(out[2])) > 0.01
