"""The CPU restatement of SyncChirpProcessor::do_linear_convolutions (oracle/chirp_oracle.c)
against an independent float64 model (tests/chirp_model.py): plan arithmetic exact, results within
fp32 FFT error.  The reference has no golden vectors for this operator."""
import numpy as np
import pytest

import chirp_model
import synth
from oracle import bindings


@pytest.mark.parametrize("in_len,inv_len,limit", [
    ([1000], 300, 256), ([5000, 3000, 5000], 4096, 1024), ([40000, 100], 9000, 0),
    ([70000], 70000, 32768), ([128, 129, 127], 128, 128), ([10], 3, 1), ([2000], 2500, 100000),
])
def test_plan_matches_the_model(in_len, inv_len, limit):
    got = bindings.chirp_plan(in_len, inv_len, limit)
    want = chirp_model.plan(in_len, inv_len, limit)
    assert got == want


@pytest.mark.parametrize("in_len,inv_len,limit", [
    ([1000], 300, 256), ([5000, 3000, 4999], 4096, 1024), ([20000, 100], 9000, 0),
    ([300, 700], 512, 128), ([2000], 2500, 4096),
])
def test_linear_convolutions_match_float64(in_len, inv_len, limit):
    inputs = [synth.noise(40 + c, n) for c, n in enumerate(in_len)]
    inverse = synth.decaying_ir(41, inv_len)[::-1].copy()        # a time-reversed decay, like an inverse chirp
    scale = 0.37
    got = bindings.chirp_linear_convolutions(inputs, inverse, limit, scale)
    want, pl = chirp_model.linear_convolutions(inputs, inverse, limit, scale)
    assert got.shape == want.shape
    peak = np.max(np.abs(want))
    assert np.max(np.abs(got - want)) <= 1e-5 * peak
    # the mul_k2 quirk: samples past vConvLengths of a SHORT channel are left unscaled
    for ch in range(len(in_len)):
        a, c = pl["align_offsets"][ch], pl["conv_lengths"][ch]
        if a > 0:
            tail = slice(c, a + c)
            assert np.allclose(got[ch, tail], want[ch, tail], atol=1e-5 * peak)
