"""The batched device restatements of lsp::dsp::fastconv_* (include/b200conv.h) against their
contract and the CPU oracle's restated primitives."""
import ctypes

import numpy as np
import pytest

import synth
from oracle.bindings import CpuConvolver, direct_convolve

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    p = ge.load()
    p.lib()
    return p


def _check(pkg, rc):
    assert rc == 0, pkg.lib().b200conv_last_error().decode()


@pytest.mark.parametrize("rank", [8, 9, 10, 11, 12, 13, 14, 15, 16])
def test_parse_apply_restore(pkg, rank):
    lib = pkg.lib()
    n, F = 1 << rank, 1 << (rank - 1)
    count = 5 if rank <= 13 else 2
    rng = np.random.Generator(np.random.PCG64(rank))
    a = rng.uniform(-1, 1, (count, F)).astype(np.float32)
    b = rng.uniform(-1, 1, (count, F)).astype(np.float32)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    ia = torch.empty((count, n), device="cuda", dtype=torch.float32)     # image: 2^rank floats
    ib = torch.empty_like(ia)
    _check(pkg, lib.b200conv_fastconv_parse(0, ia.data_ptr(), da.data_ptr(), rank, count, None))
    _check(pkg, lib.b200conv_fastconv_parse(0, ib.data_ptr(), db.data_ptr(), rank, count, None))

    want = np.stack([np.convolve(a[i].astype(np.float64), b[i].astype(np.float64)) for i in range(count)])
    want = np.concatenate([want, np.zeros((count, 1))], axis=1)          # length 2F
    tol = 1e-5 * np.abs(want).max()

    # apply ACCUMULATES into dst (the reference depends on it, Convolver.cpp:278-283)
    dst = torch.ones((count, n), device="cuda", dtype=torch.float32)
    _check(pkg, lib.b200conv_fastconv_apply(0, dst.data_ptr(), ia.data_ptr(), ib.data_ptr(), rank, count, None))
    torch.cuda.synchronize()
    assert np.max(np.abs(dst.cpu().numpy() - 1.0 - want)) <= tol

    # parse_apply == parse + apply
    dst2 = torch.full((count, n), 2.0, device="cuda", dtype=torch.float32)
    _check(pkg, lib.b200conv_fastconv_parse_apply(0, dst2.data_ptr(), ia.data_ptr(), db.data_ptr(), rank, count, None))
    torch.cuda.synchronize()
    assert np.max(np.abs(dst2.cpu().numpy() - 2.0 - want)) <= tol

    # restore(parse(a)) == [a, 0 ...] and STORES
    dst3 = torch.full((count, n), 7.0, device="cuda", dtype=torch.float32)
    _check(pkg, lib.b200conv_fastconv_restore(0, dst3.data_ptr(), ia.data_ptr(), rank, count, None))
    torch.cuda.synchronize()
    got = dst3.cpu().numpy()
    assert np.max(np.abs(got[:, :F] - a)) <= 2e-6 and np.max(np.abs(got[:, F:])) <= 2e-6

    # same numbers as the oracle's restated primitives (different image layout, same contract)
    olib = CpuConvolver.lib("oracle")[0]
    fp = ctypes.POINTER(ctypes.c_float)
    P = lambda x: x.ctypes.data_as(fp)
    oa, ob, tmp = (np.zeros(2 * n, np.float32) for _ in range(3))
    olib.rs_fastconv_parse(P(oa), P(a[0]), ctypes.c_size_t(rank))
    olib.rs_fastconv_parse(P(ob), P(b[0]), ctypes.c_size_t(rank))
    od = np.ones(n, np.float32)
    olib.rs_fastconv_apply(P(od), P(tmp), P(oa), P(ob), ctypes.c_size_t(rank))
    assert np.max(np.abs(dst.cpu().numpy()[0] - od)) <= tol


def test_primitives_reject_bad_arguments(pkg):
    lib = pkg.lib()
    x = torch.zeros(4096, device="cuda")
    assert lib.b200conv_fastconv_parse(0, x.data_ptr(), x.data_ptr(), 7, 1, None) == pkg.ERR_ARG
    assert lib.b200conv_fastconv_parse(0, x.data_ptr(), x.data_ptr(), 17, 1, None) == pkg.ERR_ARG
    assert lib.b200conv_fastconv_restore(0, x.data_ptr(), x.data_ptr(), 9, 0, None) == pkg.ERR_ARG


@pytest.mark.parametrize("nx,nh,rank,count", [(1000, 300, 8, 3), (48000, 20000, 11, 2), (5, 7, 9, 1),
                                              (100000, 100000, 13, 1)])
def test_linear_convolve_matches_float64(pkg, nx, nh, rank, count):
    """Offline full linear convolution (scope-table row f1: SyncChirpProcessor::do_linear_convolution,
    reference src/main/util/SyncChirpProcessor.cpp:1406-1508) as one batched multi-frame pass."""
    from oracle.bindings import direct_convolve
    rng = np.random.Generator(np.random.PCG64(nx + nh))
    x = rng.uniform(-1, 1, (count, nx)).astype(np.float32)
    h = rng.uniform(-1, 1, nh).astype(np.float32)
    got = pkg.linear_convolve(x, h, rank=rank)
    assert got.shape == (count, nx + nh - 1)
    for i in range(count):
        want = direct_convolve(x[i], h)
        assert np.max(np.abs(got[i] - want)) <= 1e-5 * np.max(np.abs(want))


def test_direct_convolve_accumulates(pkg):
    """dsp::convolve on the device (scope row a16): dst[a+b] += src[a]*conv[b], batched."""
    lib = pkg.lib()
    rng = np.random.Generator(np.random.PCG64(16))
    for length, count, batch in ((128, 127, 3), (31, 0x2000, 1), (5, 1, 2), (1, 9, 4)):
        x = rng.uniform(-1, 1, (batch, count)).astype(np.float32)
        h = rng.uniform(-1, 1, (batch, length)).astype(np.float32)
        n = count + length - 1
        base = rng.uniform(-1, 1, (batch, n + 3)).astype(np.float32)       # row pitch n + 3
        dx, dh, dy = torch.from_numpy(x).cuda(), torch.from_numpy(h).cuda(), torch.from_numpy(base).cuda()
        _check(pkg, lib.b200conv_convolve(0, dy.data_ptr(), n + 3, dx.data_ptr(), count, dh.data_ptr(), length,
                                          length, count, batch, None))
        torch.cuda.synchronize()
        got = dy.cpu().numpy()
        for i in range(batch):
            want = base[i].astype(np.float64)
            want[:n] += np.convolve(x[i].astype(np.float64), h[i].astype(np.float64))
            assert np.max(np.abs(got[i] - want)) <= 1e-5 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("rank,count", [(9, 5000), (11, 2000), (12, 1000), (8, 6000)])
def test_large_batches_take_the_staged_transforms(pkg, rank, count):
    """Launches that loop over several jobs per CTA on the ping-pong ranks run k_fwd_staged /
    k_inv_staged (next job's input through TMA into a two-slot ring): parse -> restore round trip
    and parse_apply against float64 FFT convolution, every row checked."""
    lib = pkg.lib()
    n, F = 1 << rank, 1 << (rank - 1)
    rng = np.random.Generator(np.random.PCG64(100 + rank))
    a = rng.uniform(-1, 1, (count, F)).astype(np.float32)
    h = rng.uniform(-1, 1, F).astype(np.float32)
    da = torch.from_numpy(a).cuda()
    dh = torch.from_numpy(np.tile(h, (count, 1))).cuda()
    ia = torch.empty((count, n), device="cuda", dtype=torch.float32)
    ih = torch.empty_like(ia)
    _check(pkg, lib.b200conv_fastconv_parse(0, ia.data_ptr(), da.data_ptr(), rank, count, None))
    _check(pkg, lib.b200conv_fastconv_parse(0, ih.data_ptr(), dh.data_ptr(), rank, count, None))

    back = torch.full((count, n), 3.0, device="cuda", dtype=torch.float32)
    _check(pkg, lib.b200conv_fastconv_restore(0, back.data_ptr(), ia.data_ptr(), rank, count, None))
    torch.cuda.synchronize()
    got = back.cpu().numpy()
    assert np.max(np.abs(got[:, :F] - a)) <= 1e-5
    assert np.max(np.abs(got[:, F:])) <= 1e-5

    dst = torch.zeros((count, n), device="cuda", dtype=torch.float32)
    _check(pkg, lib.b200conv_fastconv_parse_apply(0, dst.data_ptr(), ih.data_ptr(), da.data_ptr(), rank, count, None))
    torch.cuda.synchronize()
    want = np.fft.irfft(np.fft.rfft(a.astype(np.float64), n, axis=1) * np.fft.rfft(h.astype(np.float64), n)[None, :],
                        n, axis=1)
    assert np.max(np.abs(dst.cpu().numpy() - want)) <= 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("in_len,inv_len,limit", [
    ([1000], 300, 256), ([5000, 3000, 4999], 4096, 1024), ([20000, 100], 9000, 128),
    ([300, 700], 512, 128), ([2000], 2500, 4096), ([120000, 120000, 90000, 120000], 48000, 0),
    ([70000], 96000, 32768),
])
def test_chirp_linear_convolutions_match_the_reference_operator(pkg, in_len, inv_len, limit):
    """Row f1: SyncChirpProcessor::do_linear_convolutions (SyncChirpProcessor.cpp:1374-1508) --
    prepend-padded inverse filter, per-channel align offsets, scale over the first vConvLengths
    samples -- against the CPU restatement of that operator AND the float64 model, at the
    reference's partition-rank rule (:1224-1250)."""
    import chirp_model
    from oracle import bindings
    inputs = [synth.noise(40 + c, n) for c, n in enumerate(in_len)]
    inverse = synth.decaying_ir(41, inv_len)[::-1].copy()
    scale = 0.37
    got = pkg.chirp_linear_convolutions(inputs, inverse, limit, scale, device=0)
    want = bindings.chirp_linear_convolutions(inputs, inverse, limit, scale)
    truth, pl = chirp_model.linear_convolutions(inputs, inverse, limit, scale)
    assert got.shape == want.shape == truth.shape
    peak = float(np.max(np.abs(truth)))
    assert np.max(np.abs(got - want.astype(np.float64))) <= 1e-5 * peak
    assert np.max(np.abs(got - truth)) <= 1e-5 * peak
    # layout: nothing before the align offset
    for ch in range(len(in_len)):
        assert not got[ch, :pl["align_offsets"][ch]].any()


def test_chirp_rejects_partitions_below_the_engine_ranks(pkg):
    with pytest.raises(pkg.B200ConvError):
        pkg.chirp_linear_convolutions([synth.noise(1, 100)], synth.noise(2, 50), 64, 1.0, device=0)


def test_shared_impulse_response_instances(pkg):
    """b200conv_init_shared: many channels through one reverb share one set of IR spectra; the
    lender cannot go away while borrowed from."""
    n, taps, rank, F = 6, 30000, 10, 512
    ir = synth.decaying_ir(3, taps)
    b = pkg.ConvolverBatch(n, 0)
    assert b.init(0, ir, rank, 0.0)
    for c in range(1, n):
        assert b.init_shared(c, 0 if c < 4 else 2, 0.25 * (c % 2))      # borrowing from a borrower resolves to the owner
    x = np.stack([synth.noise(60 + c, 12 * F) for c in range(n)])
    out = np.concatenate([b.process(x[:, i * 300:(i + 1) * 300].copy()) for i in range(12 * F // 300)], axis=1)
    for c in range(n):
        want = direct_convolve(x[c], ir, out.shape[1])
        assert np.max(np.abs(out[c] - want)) <= 1e-5 * np.max(np.abs(want))
    with pytest.raises(pkg.B200ConvError):
        b.destroy(0)
    with pytest.raises(pkg.B200ConvError):
        b.init(0, ir[:100], rank, 0.0)
    for c in range(1, n):
        b.destroy(c)
    b.destroy(0)
    assert b.rank(0) == 0
    b.close()
