"""Scope-table row f4: the reference's own lsp::dspu::SpectralProcessor (compiled verbatim into
oracle/_ref) pinned by the reference's unit test (src/test/utest/util/spectral_proc.cpp:33-66) and
by an independent float64 model (tests/spectral_model.py).  CPU only."""
import numpy as np
import pytest

import spectral_model
import synth
from oracle.bindings import CpuSpectralProcessor

pytestmark = pytest.mark.skipif(not CpuSpectralProcessor.available(), reason="oracle/_ref has not been built")


def test_reference_utest_simple():
    # spectral_proc.cpp:33-62 : 440 Hz sine, init(14), phase 0, rank 8, ONE call of 8192 samples;
    # src[i] == dst[latency + i] within 1e-5
    n = 8192
    src = np.sin(2 * np.pi * 440.0 / 48000.0 * np.arange(n)).astype(np.float32)
    sp = CpuSpectralProcessor(14)
    sp.set_phase(0.0)
    sp.set_rank(8)
    dst = sp.process(src)
    lat = sp.latency()
    assert lat == 256
    assert np.max(np.abs(src[:n - lat] - dst[lat:])) <= 1e-5


@pytest.mark.parametrize("rank,phase,step", [(8, 0.0, 31), (9, 0.5, 256), (10, 0.37, 1000), (12, 1.0, 4096), (7, 0.25, 77)])
def test_identity_hooks_and_phase_against_the_model(rank, phase, step):
    n = 6 * (1 << rank) + 123
    src = synth.noise(rank, n)
    N = 1 << rank
    rng = np.random.Generator(np.random.PCG64(rank))
    gain = rng.uniform(0.2, 1.5, N).astype(np.float32)
    half = rng.uniform(-1, 1, N // 2 + 1) + 1j * rng.uniform(-1, 1, N // 2 + 1)
    half[0] = half[0].real
    half[-1] = half[-1].real
    H = np.concatenate([half, np.conj(half[-2:0:-1])]).astype(np.complex64)       # conjugate-symmetric: real output
    for kind, table, hook in ((0, None, None), (2, gain, lambda X: X * gain.astype(np.float64)),
                              (1, H, lambda X: X * H.astype(np.complex128))):
        sp = CpuSpectralProcessor(14)
        sp.set_rank(rank)
        sp.set_phase(phase)
        if kind == 1:
            sp.bind_complex(table)
        elif kind == 2:
            sp.bind_gain(table)
        got = sp.run(src, step)
        want = spectral_model.ModelSpectralProcessor(rank, phase, hook).process(src)
        assert np.max(np.abs(got - want)) <= 2e-5 * max(1.0, np.max(np.abs(want)))
        assert sp.latency() == N
