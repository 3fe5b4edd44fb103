"""GPU parity: the CUDA engine, through the C ABI, against the CPU oracle (same seeded inputs).

Tolerance (BASELINE.json north_star): max |gpu - ref| <= 1e-5 of the reference peak
(<= -100 dB), zero latency, frame boundaries at the same sample indices.  Shapes mirror the
reference's own unit test src/test/utest/util/convolver.cpp and the BASELINE configs.
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
    p.lib()         # fails loudly when the CUDA library has not been built
    return p


def rel_err(got, want):
    return float(np.max(np.abs(np.asarray(got, np.float64) - want)) / np.max(np.abs(want)))


def oracle_run(ir, src, rank, phase, step):
    c = CpuConvolver("oracle")
    assert c.init(ir, rank, phase)
    return c.run(src, step)


# ---- the reference's own unit test, re-expressed against the GPU engine ----------------------

def test_utest_small(pkg):
    # convolver.cpp:88-136 : rel 1e-4 vs naive direct convolution, calls of 31
    ir, src = synth.utest_small()
    c = pkg.Convolver(0)
    assert c.init(ir, 9, 0.0)
    out = c.run(src, 31)
    want = direct_convolve(src, ir, src.size)
    assert synth.equals_relative(out, want, 1e-4)
    assert rel_err(out, want) <= TOL
    c.destroy()
    assert c.rank() == 0 and c.data_size() == 0


def test_utest_large(pkg):
    # convolver.cpp:184-223 : abs 1e-4, 8192 taps, rank 10, calls of 31
    ir, src = synth.utest_large()
    c = pkg.Convolver(0)
    assert c.init(ir, 10, 0.0)
    out = c.run(src, 31)
    assert np.max(np.abs(out - direct_convolve(src, ir, src.size))) <= 1e-4
    assert rel_err(out, oracle_run(ir, src, 10, 0.0, 31).astype(np.float64)) <= TOL


def test_utest_collisions_subsampled(pkg):
    # convolver.cpp:138-182 (disabled upstream): 65536 taps, rank 10, calls of 127, abs 1e-5,
    # tail flushed with data_size()-1 zeros.  Larger calls here keep the runtime bounded; one
    # offset also runs with the original 127-sample calls.
    rng = np.random.Generator(np.random.PCG64(11))
    L = 0x10000
    ir = rng.uniform(-1.0, 1.0, L).astype(np.float32)
    for i, step in ((1, 127 * 64), (513, 127 * 64), (40000, 127 * 64), (4095, 127)):
        c = pkg.Convolver(0)
        assert c.init(ir, 10, 0.0)
        total = L + c.data_size() - 1
        if step == 127:
            total = 3 * 4096      # bounded: 127-sample calls are ~100 kernel launches each
        src = np.zeros(total, dtype=np.float32)
        src[0] = 1.0
        src[i] = 1.0
        out = c.run(src, step)
        want = direct_convolve(src[:L], ir)[:total]
        assert np.max(np.abs(out - want)) <= 1e-5, (i, step)


def test_api_contract(pkg):
    c = pkg.Convolver(0)
    x = synth.noise(0, 300)
    assert not c.process(x).any()                               # Convolver.cpp:219-223
    assert c.init(np.ones(10, np.float32), 3, 0.0) and c.rank() == 8        # :87
    assert c.init(np.ones(10, np.float32), 20, 0.0) and c.rank() == 16
    assert c.data_size() == 10
    assert c.init(np.zeros(0, np.float32), 10, 0.0)             # :80-84
    assert c.rank() == 0 and c.data_size() == 0
    assert not c.process(x).any()
    c.init(np.ones(4, np.float32), 8, 0.0)
    assert c.process(np.zeros(0, np.float32)).size == 0


@pytest.mark.parametrize("taps,rank,phase,step", [(65536, 11, 0.0, 1024), (65536, 11, 0.37, 1000),
                                                  (5000, 8, 0.0, 77), (5000, 16, 0.5, 4096),
                                                  (70000, 13, 0.9, 333), (200, 12, 0.0, 256),
                                                  (129, 9, 0.0, 64), (1, 8, 0.0, 5),
                                                  (9000, 14, 0.0, 8192), (40000, 15, 0.1, 16384),
                                                  (40000, 16, 0.0, 32768), (3000, 10, 0.0, 512),
                                                  (3000, 12, 0.5, 2048)])
def test_matches_oracle_any_rank_phase_call_size(pkg, taps, rank, phase, step):
    ir = synth.decaying_ir(3, taps)
    n = min(3 * taps + 500, 90000)
    if step >= 4096:
        n = ((n + step - 1) // step) * step
    src = synth.noise(3, n)
    c = pkg.Convolver(0)
    assert c.init(ir, rank, phase)
    out = c.run(src, step)
    want = oracle_run(ir, src, rank, phase, step)
    truth = direct_convolve(src, ir, src.size)
    assert rel_err(out, want.astype(np.float64)) <= TOL
    assert rel_err(out, truth) <= TOL
    assert c.state()["frame_off"] == (int(np.float32(phase) * np.float32(1 << (c.rank() - 1)))
                                      + src.size) % (1 << (c.rank() - 1))


def test_golden_fixtures(pkg, golden):
    """Outputs frozen from the reference's Convolver.cpp compiled verbatim (tests/golden)."""
    for name in sorted({k.split(".")[0] for k in golden.files}):
        rank, phase, step, eff_rank, size = golden[name + ".meta"]
        c = pkg.Convolver(0)
        assert c.init(golden[name + ".ir"], int(rank), float(phase))
        assert (c.rank(), c.data_size()) == (int(eff_rank), int(size))
        out = c.run(golden[name + ".src"], int(step))
        want = golden[name + ".dst"].astype(np.float64)
        assert rel_err(out, want) <= TOL, name


def test_inplace_and_random_call_sizes(pkg):
    ir = synth.decaying_ir(5, 5000)
    src = synth.noise(5, 12000)
    rng = np.random.Generator(np.random.PCG64(3))
    c = pkg.Convolver(0)
    assert c.init(ir, 9, 0.37)
    buf = src.copy()
    i = 0
    while i < buf.size:
        n = int(rng.integers(1, 700))
        c.process(buf[i:i + n], out=buf[i:i + n])           # dst == src
        i += n
    assert rel_err(buf, direct_convolve(src, ir, src.size)) <= TOL


# ---- batches ------------------------------------------------------------------------------------

def test_config1_mono_65536_taps_1024_blocks(pkg):
    # BASELINE config 1 (shortened to 3 s of input + the reference's flush of L-1 zeros)
    L, rank, step = 65536, 11, 1024
    ir = synth.decaying_ir(0, L)
    x = synth.noise(0, 3 * 48000)
    total = ((x.size + L - 1 + step - 1) // step) * step
    src = np.zeros(total, np.float32)
    src[:x.size] = x
    c = pkg.Convolver(0)
    assert c.init(ir, rank, 0.0)
    out = c.run(src, step)
    assert rel_err(out, oracle_run(ir, src, rank, 0.0, step).astype(np.float64)) <= TOL
    assert rel_err(out, direct_convolve(src, ir, src.size)) <= TOL


def test_config2_stereo_phases_256_blocks(pkg):
    # BASELINE config 2 (IR shortened to 1 s): two instances, phases 0 and 0.5, rank 9 and 8
    for rank in (9, 8):
        L, step, n = 48000, 256, 256 * 400
        b = pkg.ConvolverBatch(2, 0)
        irs = [synth.decaying_ir(c, L) for c in range(2)]
        src = np.stack([synth.noise(c, n) for c in range(2)])
        for c, ph in enumerate((0.0, 0.5)):
            assert b.init(c, irs[c], rank, ph)
        out = np.empty_like(src)
        for i in range(0, n, step):
            out[:, i:i + step] = b.process(src[:, i:i + step])
        for c, ph in enumerate((0.0, 0.5)):
            assert rel_err(out[c], oracle_run(irs[c], src[c], rank, ph, step).astype(np.float64)) <= TOL
            assert rel_err(out[c], direct_convolve(src[c], irs[c], n)) <= TOL
        b.close()


def test_config3_batch_subset(pkg):
    # BASELINE config 3 at full IR length on 4 of the 64 channels (the CPU oracle is slow):
    # 480000-tap IRs, rank 11, 1024-sample calls
    L, rank, step, nblk, n = 480000, 11, 1024, 96, 4
    b = pkg.ConvolverBatch(n, 0)
    irs = [synth.decaying_ir(c, L) for c in range(n)]
    src = np.stack([synth.noise(c, nblk * step) for c in range(n)])
    for c in range(n):
        assert b.init(c, irs[c], rank, 0.0)
        assert b.state(c)["bins"] == 469
    out = np.empty_like(src)
    for i in range(0, nblk * step, step):
        out[:, i:i + step] = b.process(src[:, i:i + step])
    for c in range(n):
        assert rel_err(out[c], oracle_run(irs[c], src[c], rank, 0.0, step).astype(np.float64)) <= TOL
    b.close()


def test_ragged_batch_mixed_lengths_and_uninitialised(pkg):
    # ragged IR lengths in one batch, one instance never initialised, one destroyed mid-stream
    rank, step, n = 10, 512, 512 * 12
    lens = [1, 511, 512, 513, 7000, 0, 30000]
    b = pkg.ConvolverBatch(len(lens), 0)
    irs = [synth.decaying_ir(c, L) if L else None for c, L in enumerate(lens)]
    src = np.stack([synth.noise(c, n) for c in range(len(lens))])
    for c, ir in enumerate(irs):
        if ir is not None:
            assert b.init(c, ir, rank, 0.0)
    out = np.empty_like(src)
    half = n // 2
    for i in range(0, half, step):
        out[:, i:i + step] = b.process(src[:, i:i + step])
    b.destroy(1)
    for i in range(half, n, step):
        out[:, i:i + step] = b.process(src[:, i:i + step])
    for c, ir in enumerate(irs):
        if ir is None:
            assert not out[c].any()
        elif c == 1:
            assert rel_err(out[c][:half], direct_convolve(src[c][:half], ir, half)) <= TOL
            assert not out[c][half:].any()
        else:
            assert rel_err(out[c], direct_convolve(src[c], ir, n)) <= TOL
    # a different rank in a live batch is refused
    with pytest.raises(pkg.B200ConvError):
        b.init(5, np.ones(10, np.float32), 12, 0.0)
    b.close()


def test_reinit_discards_history(pkg):
    ir1, ir2 = synth.decaying_ir(1, 4000), synth.decaying_ir(2, 900)
    x = synth.noise(9, 4096)
    c = pkg.Convolver(0)
    assert c.init(ir1, 9, 0.0)
    c.run(x, 256)
    assert c.init(ir2, 9, 0.25)                 # Convolver.cpp:108-110: slab zeroed
    out = c.run(x, 100)
    assert rel_err(out, direct_convolve(x, ir2, x.size)) <= TOL


def test_partition_range_shards_sum_to_full(pkg):
    # SURVEY 8e / BASELINE config 5 in miniature: one IR split by partition range over 3 shards
    rank, F, L, n, step = 9, 256, 9000, 256 * 40, 256
    ir, x = synth.decaying_ir(2, L), synth.noise(2, n)
    bins = (L + F - 1) // F
    cuts = [0, 10, 23, bins]
    total = np.zeros(n, np.float64)
    for a, e in zip(cuts[:-1], cuts[1:]):
        b = pkg.ConvolverBatch(1, 0)
        assert b.init(0, ir[a * F:e * F], rank, 0.0, part_offset=a)
        out = np.concatenate([b.process(x[None, i:i + step])[0] for i in range(0, n, step)])
        total += out
        b.close()
    assert rel_err(total, direct_convolve(x, ir, n)) <= TOL


def test_linearity_and_impulse_at_full_size(pkg):
    """Size-independent properties on a full-length (10 s) IR: an impulse returns the IR itself
    (bit-aligned: zero latency), and the response to a sum is the sum of the responses."""
    L, rank, step = 480000, 11, 1024
    ir = synth.decaying_ir(7, L)
    nblk = 40
    n = nblk * step
    imp = np.zeros(n, np.float32)
    imp[5] = 1.0
    a, bsig = synth.noise(1, n), synth.noise(2, n)
    batch = pkg.ConvolverBatch(4, 0)
    for c in range(4):
        assert batch.init(c, ir, rank, 0.0)
    src = np.stack([imp, a, bsig, a + bsig])
    out = np.empty_like(src)
    for i in range(0, n, step):
        out[:, i:i + step] = batch.process(src[:, i:i + step])
    peak = np.abs(ir).max()
    assert np.max(np.abs(out[0][5:] - ir[:n - 5])) <= TOL * peak
    assert np.max(np.abs(out[0][:5])) <= TOL * peak          # nothing before the impulse (fp32 FFT noise only)
    assert np.max(np.abs(out[1] + out[2] - out[3])) <= 4 * TOL * np.abs(out[3]).max()
    batch.close()


@pytest.mark.parametrize("fused", [1, 0])
def test_device_pointer_api_and_stats(pkg, fused):
    torch = pytest.importorskip("torch")
    n, L, rank, F, nblk = 8, 20000, 11, 1024, 24
    b = pkg.ConvolverBatch(n, 0)
    b.set_option("fused", fused)
    irs = [synth.decaying_ir(c, L) for c in range(n)]
    for c in range(n):
        assert b.init(c, irs[c], rank, 0.0)
    src = np.stack([synth.noise(c, nblk * F) for c in range(n)])
    dsrc = torch.from_numpy(src).cuda()
    ddst = torch.empty_like(dsrc)
    torch.cuda.synchronize()
    b.reset_stats()
    stride = dsrc.shape[1]
    for i in range(nblk):
        b.process_device(ddst.data_ptr() + 4 * i * F, dsrc.data_ptr() + 4 * i * F, stride, F)
    b.sync()
    st = b.stats()
    assert st["launches"] == (1 if fused else 3) * nblk and st["mac_launches"] == nblk
    bins = (L + F - 1) // F
    assert st["mac_algo_bytes"] == nblk * n * (16 * F * bins + 24 * F)
    out = ddst.cpu().numpy()
    for c in range(n):
        assert rel_err(out[c], direct_convolve(src[c], irs[c], nblk * F)) <= TOL
    b.close()


@pytest.mark.parametrize("rank", [8, 9, 10, 11, 12, 13])
@pytest.mark.parametrize("opts", [dict(fused=1), dict(fused=0), dict(fused=1, pdl=0, fft_bias=0),
                                  dict(fused=1, mac_splits=7, mac_stages=2),
                                  dict(fused=1, mac_splits=1, mac_stages=5)])
def test_one_launch_per_block_matches_three_kernel_path(pkg, rank, opts):
    """k_frame (FFT + MAC + IFFT in one launch, ranks 8..13) against direct convolution on a ragged
    batch: IRs shorter than the number of partition splits, exactly one frame, many frames."""
    F = 1 << (rank - 1)
    lens = [1, F - 1, F, F + 1, 5 * F + 3, 40 * F + 7, 0, 200 * F]
    nblk = 24
    b = pkg.ConvolverBatch(len(lens), 0)
    b.set_option("multi_frame", 1)                           # two-frame calls as two single launches
    for k, v in opts.items():
        b.set_option(k, v)
    irs = [synth.decaying_ir(c, L) if L else None for c, L in enumerate(lens)]
    for c, ir in enumerate(irs):
        if ir is not None:
            assert b.init(c, ir, rank, 0.0)
    src = np.stack([synth.noise(40 + c, nblk * F) for c in range(len(lens))])
    out = np.empty_like(src)
    for i in range(0, nblk * F, 2 * F):                      # two frames per call
        out[:, i:i + 2 * F] = b.process(src[:, i:i + 2 * F])
    for c, ir in enumerate(irs):
        if ir is None:
            assert not out[c].any()
        else:
            assert rel_err(out[c], direct_convolve(src[c], ir, nblk * F)) <= TOL, (c, lens[c])
    b.close()


def test_planar_host_api_strided_views(pkg):
    """b200conv_process_planar: rows of one host matrix, strided views, in place."""
    n, L, rank, F, nblk = 5, 7000, 10, 512, 10
    b = pkg.ConvolverBatch(n, 0)
    irs = [synth.decaying_ir(c, L) for c in range(n)]
    for c in range(n):
        assert b.init(c, irs[c], rank, 0.0 if c % 2 == 0 else 0.3)
    src = np.stack([synth.noise(60 + c, nblk * F) for c in range(n)])
    buf = src.copy()
    for i in range(0, nblk * F, F):
        b.process(buf[:, i:i + F], buf[:, i:i + F])         # strided rows, dst == src
    for c in range(n):
        assert rel_err(buf[c], direct_convolve(src[c], irs[c], nblk * F)) <= TOL
    b.close()


@pytest.mark.parametrize("n,taps,frames,rank", [(8, 100 * 1024 + 3, 600, 11), (64, 480000, 150, 11),
                                                (6, 300000, 200, 12), (5, 300000, 120, 13), (7, 50000, 400, 9),
                                                (1500, 6000, 60, 11),      # more CTAs than the device holds at once
                                                (700, 40000, 40, 10)])
def test_overlapped_launches_are_bit_identical_to_serialised(pkg, n, taps, frames, rank):
    """Back-to-back blocks overlap on the GPU (programmatic dependent launch, ring_head
    hand-shake).  Any ordering bug would change bits: the overlapped run must equal the fully
    serialised one exactly, block for block."""
    torch = pytest.importorskip("torch")
    F = 1 << (rank - 1)
    irs = [synth.decaying_ir(c, taps) for c in range(min(n, 4))]
    g = torch.Generator(device="cuda").manual_seed(7)
    src = torch.rand((n, frames * F), generator=g, device="cuda") * 2 - 1
    outs = []
    for pdl in (1, 0):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("pdl", pdl)
        for c in range(n):
            assert b.init(c, irs[c % len(irs)], rank, 0.0)
        dst = torch.zeros_like(src)
        for i in range(frames):
            b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, frames * F, F)
        b.sync()
        outs.append(dst.cpu().numpy())
        b.close()
    assert np.array_equal(outs[0], outs[1])
    # and it is the right answer (first channel, float64 truth)
    want = direct_convolve(src[0].cpu().numpy(), irs[0], frames * F)
    assert rel_err(outs[0][0], want) <= TOL


def test_planar_zero_copy_from_pinned_host_memory(pkg):
    """Page-locked host matrices are read / written by the kernels directly (no staging copies);
    the result is bit-identical to the staged path and the input matrix is left untouched."""
    torch = pytest.importorskip("torch")
    n, L, rank, F, nblk = 6, 30000, 11, 1024, 12
    irs = [synth.decaying_ir(c, L) for c in range(n)]
    src = np.stack([synth.noise(70 + c, nblk * F) for c in range(n)])
    outs = []
    for zero_copy in (1, 0):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("zero_copy", zero_copy)
        for c in range(n):
            assert b.init(c, irs[c], rank, 0.0)
        hsrc = torch.from_numpy(src.copy()).pin_memory()
        hdst = torch.zeros_like(hsrc).pin_memory()
        hs, hd = hsrc.numpy(), hdst.numpy()
        for i in range(0, nblk * F, F):
            b.process(hs[:, i:i + F], hd[:, i:i + F])
        assert np.array_equal(hs, src)
        outs.append(hd.copy())
        b.close()
    assert np.array_equal(outs[0], outs[1])
    for c in range(n):
        assert rel_err(outs[0][c], direct_convolve(src[c], irs[c], nblk * F)) <= TOL


@pytest.mark.parametrize("rank", [8, 10, 11, 12])
@pytest.mark.parametrize("multi", [8, 4, 2])
def test_multi_frame_calls_share_one_pass_over_the_ir(pkg, rank, multi):
    """Calls that bring several whole frames (offline rendering) run k_mac_multi: one pass over
    the IR spectra serves up to 8 frames.  Mixed call lengths (groups of 8/4/2 and single-frame
    launches, then back to one-frame calls) on a ragged batch, against direct convolution."""
    F = 1 << (rank - 1)
    lens = [1, F, 3 * F + 5, 37 * F + 11, 0, 150 * F]
    b = pkg.ConvolverBatch(len(lens), 0)
    b.set_option("multi_frame", multi)
    irs = [synth.decaying_ir(c, L) if L else None for c, L in enumerate(lens)]
    for c, ir in enumerate(irs):
        if ir is not None:
            assert b.init(c, ir, rank, 0.0)
    calls = [3, 8, 13, 1, 1, 21, 2, 1, 16, 5]
    total = sum(calls) * F
    src = np.stack([synth.noise(80 + c, total) for c in range(len(lens))])
    out = np.empty_like(src)
    i = 0
    for nf in calls:
        out[:, i:i + nf * F] = b.process(src[:, i:i + nf * F])
        i += nf * F
    for c, ir in enumerate(irs):
        if ir is None:
            assert not out[c].any()
        else:
            assert rel_err(out[c], direct_convolve(src[c], ir, total)) <= TOL, (c, lens[c])
    b.close()


@pytest.mark.parametrize("block", [1024, 700])
def test_device_api_with_separate_row_strides(pkg, block):
    """b200conv_process_device2: the input and output matrices have their own row pitch
    (whole-frame path for 1024-sample calls, partial-frame path for 700)."""
    torch = pytest.importorskip("torch")
    n, L, rank, nblk = 5, 9000, 11, 9
    b = pkg.ConvolverBatch(n, 0)
    irs = [synth.decaying_ir(c, L) for c in range(n)]
    for c in range(n):
        assert b.init(c, irs[c], rank, 0.0)
    src = np.stack([synth.noise(90 + c, nblk * block) for c in range(n)])
    dsrc = torch.from_numpy(src).cuda()                     # row pitch nblk * block
    guard = 64
    ddst = torch.full((n, block + guard), 7.0, device="cuda")   # one block per row + a guard band
    out = np.empty_like(src)
    torch.cuda.synchronize()
    for i in range(nblk):
        b.process_device(ddst.data_ptr(), dsrc.data_ptr() + 4 * i * block, nblk * block, block,
                         dst_stride=block + guard)
        b.sync()
        blk = ddst.cpu().numpy()
        assert np.all(blk[:, block:] == 7.0)                # nothing written past the block
        out[:, i * block:(i + 1) * block] = blk[:, :block]
    for c in range(n):
        assert rel_err(out[c], direct_convolve(src[c], irs[c], nblk * block)) <= TOL
    b.close()


def test_failed_init_keeps_the_previous_state(pkg):
    """Convolver.cpp:103-108: init returns false on allocation failure and the old convolver keeps
    working (the new slab is allocated before the old one is released)."""
    import ctypes
    ir, x = synth.decaying_ir(1, 5000), synth.noise(1, 8 * 1024)
    b = pkg.ConvolverBatch(1, 0)
    assert b.init(0, ir, 11, 0.0)
    out = np.empty_like(x)
    out[:4096] = b.process(x[None, :4096])[0]
    # 2^40 taps cannot be allocated: the call must fail before it touches the (tiny) buffer
    rc = pkg.lib().b200conv_init(b._h, 0, ir.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 1 << 40, 11, 0.0)
    assert rc == pkg.ERR_NOMEM
    assert b.data_size(0) == 5000 and b.rank(0) == 11 and b.state(0)["frames"] == 4
    out[4096:] = b.process(x[None, 4096:])[0]
    assert rel_err(out, direct_convolve(x, ir, x.size)) <= TOL
    b.close()


def test_randomised_batches_against_direct_convolution(pkg):
    """Seeded random sweep: rank, batch size, IR lengths (ragged, some instances never initialised),
    phases, call sizes (whole frames, fragments, several frames), occasional re-init and destroy --
    every output sample against float64 direct convolution of what that instance has been fed."""
    rng = np.random.Generator(np.random.PCG64(2024))
    for case in range(24):
        rank = int(rng.integers(8, 12)) if case % 6 else int(rng.integers(12, 15))
        F = 1 << (rank - 1)
        n = int(rng.integers(1, 6))
        b = pkg.ConvolverBatch(n, 0)
        if case % 3 == 0:
            b.set_option("multi_frame", int(rng.choice([1, 2, 4, 8])))
        irs, fed, got = [None] * n, [[] for _ in range(n)], [[] for _ in range(n)]

        def init(c):
            L = int(rng.choice([1, F - 1, F, F + 1, 3 * F + 7, int(rng.integers(1, 40 * F))]))
            irs[c] = synth.decaying_ir(int(rng.integers(0, 1000)), L)
            phase = float(rng.choice([0.0, 0.0, 0.25, 0.5, 0.9]))
            assert b.init(c, irs[c], rank, phase)
            fed[c], got[c] = [], []

        for c in range(n):
            if rng.random() < 0.85:
                init(c)
        aligned = bool(rng.random() < 0.5)
        for _ in range(int(rng.integers(3, 9))):
            count = int(rng.choice([1, 2, 4, 9])) * F if aligned else int(rng.integers(1, 3 * F))
            x = rng.uniform(-1, 1, (n, count)).astype(np.float32)
            y = b.process(x)
            for c in range(n):
                if irs[c] is None:
                    assert not y[c].any()
                else:
                    fed[c].append(x[c])
                    got[c].append(y[c])
            if rng.random() < 0.15:
                c = int(rng.integers(0, n))
                # verify what this instance produced so far, then replace its IR (history discarded)
                if irs[c] is not None and fed[c]:
                    xin, out = np.concatenate(fed[c]), np.concatenate(got[c])
                    want = direct_convolve(xin, irs[c], xin.size)
                    assert np.max(np.abs(out - want)) <= TOL * max(np.max(np.abs(want)), 1e-3), (case, c)
                if rng.random() < 0.5:
                    init(c)
                else:
                    b.destroy(c)
                    irs[c] = None
        for c in range(n):
            if irs[c] is not None and fed[c]:
                xin, out = np.concatenate(fed[c]), np.concatenate(got[c])
                want = direct_convolve(xin, irs[c], xin.size)
                assert np.max(np.abs(out - want)) <= TOL * max(np.max(np.abs(want)), 1e-3), (case, c, rank)
        b.close()


@pytest.mark.parametrize("eager", [1, 0])
@pytest.mark.parametrize("pinned", [False, True])
def test_eager_pending_mac_for_synchronous_calls(pkg, eager, pinned):
    """Synchronous one-frame calls: after block t is delivered the partitions q >= 1 of block t+1
    are summed while the host is away; the next call finishes with partition 0 only.  Mixed with
    multi-frame calls, fragments, a re-init and a destroy (each must invalidate the pending rows)."""
    torch = pytest.importorskip("torch")
    n, rank, F = 3, 11, 1024
    lens = [50 * F + 3, 7 * F, 200 * F - 1]
    b = pkg.ConvolverBatch(n, 0)
    b.set_option("eager", eager)
    irs = [synth.decaying_ir(c, L) for c, L in enumerate(lens)]
    for c in range(n):
        assert b.init(c, irs[c], rank, 0.0)
    calls = [F] * 6 + [3 * F] + [F] * 3 + [700, 324] + [F] * 4 + [8 * F] + [F] * 5
    total = sum(calls)
    src = np.stack([synth.noise(30 + c, total) for c in range(n)])
    if pinned:
        hsrc, hdst = torch.from_numpy(src.copy()).pin_memory(), torch.zeros((n, total)).pin_memory()
        xin, out = hsrc.numpy(), hdst.numpy()
    else:
        xin, out = src, np.zeros_like(src)
    i = 0
    for k, cnt in enumerate(calls):
        b.process(xin[:, i:i + cnt], out[:, i:i + cnt])
        i += cnt
        if k == 12:
            cut = i                                     # instance 1 gets a new IR here
            irs1_new = synth.decaying_ir(77, 9 * F + 5)
            assert b.init(1, irs1_new, rank, 0.0)
    for c in (0, 2):
        assert rel_err(out[c], direct_convolve(src[c], irs[c], total)) <= TOL
    assert rel_err(out[1][:cut], direct_convolve(src[1][:cut], irs[1], cut)) <= TOL
    assert rel_err(out[1][cut:], direct_convolve(src[1][cut:], irs1_new, total - cut)) <= TOL
    b.close()


def test_back_to_back_synchronous_calls_without_programmatic_launch(pkg):
    """pdl = 0 with the eager pending MAC: the engine asks whether the previous pending MAC is
    still running (cudaEventQuery -> "not ready") and then launches with <<< >>>; the query's
    answer must not surface as a launch error."""
    torch = pytest.importorskip("torch")
    n, rank, F, blocks = 16, 11, 1024, 60
    irs = [synth.decaying_ir(c, 300000) for c in range(2)]
    src = np.stack([synth.noise(700 + c, blocks * F) for c in range(n)])
    hsrc = torch.from_numpy(src.copy()).pin_memory()
    hdst = torch.zeros((n, blocks * F)).pin_memory()
    b = pkg.ConvolverBatch(n, 0)
    b.set_option("pdl", 0)
    for c in range(n):
        assert b.init(c, irs[c % 2], rank, 0.0)
    xin, out = hsrc.numpy(), hdst.numpy()
    for k in range(blocks):
        b.process(xin[:, k * F:(k + 1) * F], out[:, k * F:(k + 1) * F])
    b.close()
    for c in (0, 1):
        assert rel_err(out[c], direct_convolve(src[c], irs[c], blocks * F)) <= TOL


@pytest.mark.parametrize("n,taps,rank", [(64, 480000, 11), (5, 40000, 9), (150, 9000, 8)])
def test_early_pending_mac_is_bit_identical_to_the_serialised_one(pkg, n, taps, rank):
    """Back-to-back synchronous one-frame calls on page-locked matrices: the pending MAC of block
    t+1 starts under the tail of the launch that delivers block t (programmatic serialization,
    ring_head poll, rows held back by griddepcontrol.wait).  Same arithmetic in the same order as
    with early_pend = 0 (2 = always early, 1 = automatic), so the outputs must be bit-identical;
    instance 0 is also checked against float64 truth."""
    torch = pytest.importorskip("torch")
    F = 1 << (rank - 1)
    blocks = 200 if n >= 64 else 400
    irs = [synth.decaying_ir(c, taps - 13 * c) for c in range(min(n, 4))]
    src = np.stack([synth.noise(300 + c, blocks * F) for c in range(n)])
    hsrc = torch.from_numpy(src.copy()).pin_memory()
    outs = []
    for early in (2, 0):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("early_pend", early)
        for c in range(n):
            assert b.init(c, irs[c % len(irs)], rank, 0.0)
        hdst = torch.zeros((n, blocks * F)).pin_memory()
        xin, out = hsrc.numpy(), hdst.numpy()
        for k in range(blocks):
            b.process(xin[:, k * F:(k + 1) * F], out[:, k * F:(k + 1) * F])
        outs.append(out.copy())
        b.close()
    assert np.array_equal(outs[0], outs[1])
    assert rel_err(outs[0][0], direct_convolve(src[0], irs[0], blocks * F)) <= TOL


@pytest.mark.parametrize("n,rank,taps,frames", [(300, 11, 3000, 8), (700, 9, 1500, 8), (120, 12, 5000, 16)])
def test_multi_frame_calls_on_many_instances(pkg, n, rank, taps, frames):
    """Enough (instance, frame) jobs per call that the transforms run as k_fwd_staged /
    k_inv_staged (several jobs per CTA, next input prefetched by TMA); every instance is checked
    against float64 FFT convolution, two calls so that ring history is exercised."""
    from scipy.signal import fftconvolve
    F = 1 << (rank - 1)
    total = 2 * frames * F
    irs = [synth.decaying_ir(c, taps - c) for c in range(4)]
    src = np.stack([synth.noise(500 + c, total) for c in range(n)])
    b = pkg.ConvolverBatch(n, 0)
    for c in range(n):
        assert b.init(c, irs[c % 4], rank, 0.0)
    out = np.empty_like(src)
    half = frames * F
    out[:, :half] = b.process(src[:, :half])
    out[:, half:] = b.process(src[:, half:])
    b.close()
    for k in range(4):
        idx = np.arange(k, n, 4)
        want = fftconvolve(src[idx].astype(np.float64), irs[k].astype(np.float64)[None, :], axes=1)[:, :total]
        assert np.max(np.abs(out[idx] - want)) <= TOL * np.max(np.abs(want))


def test_multi_frame_device_call_with_unaligned_rows(pkg):
    """Device matrices whose rows start 4 bytes off a 16-byte boundary (and an odd row pitch): the
    staged forward transform must fall back to direct loads for those jobs (TMA bulk copies need
    16-byte aligned sources); results identical to the aligned call."""
    torch = pytest.importorskip("torch")
    from scipy.signal import fftconvolve
    n, rank, taps, frames = 300, 11, 2500, 8
    F = 1 << (rank - 1)
    total = frames * F
    irs = [synth.decaying_ir(c, taps) for c in range(2)]
    src = np.stack([synth.noise(900 + c, total) for c in range(n)])
    outs = []
    for shift, pitch in ((0, total), (1, total + 3)):
        b = pkg.ConvolverBatch(n, 0)
        for c in range(n):
            assert b.init(c, irs[c % 2], rank, 0.0)
        dsrc = torch.zeros((n * pitch + 8,), device="cuda", dtype=torch.float32)
        ddst = torch.zeros_like(dsrc)
        view = dsrc[shift:shift + n * pitch].view(n, pitch)
        view[:, :total] = torch.from_numpy(src).cuda()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        b.process_device(ddst.data_ptr() + 4 * shift, dsrc.data_ptr() + 4 * shift, pitch, total, s.cuda_stream)
        s.synchronize()
        outs.append(ddst[shift:shift + n * pitch].view(n, pitch)[:, :total].cpu().numpy())
        b.close()
    assert np.array_equal(outs[0], outs[1])
    want = fftconvolve(src[::2].astype(np.float64), irs[0].astype(np.float64)[None, :], axes=1)[:, :total]
    assert np.max(np.abs(outs[0][::2] - want)) <= TOL * np.max(np.abs(want))


def test_ir_with_more_than_65535_partitions(pkg):
    """Maximum-size edge: a 3-minute IR at the smallest rank has 70 000 partitions of 128 taps
    (beyond the 65 535 grid-y limit and the 32 768-entry job ring of the IR ingest)."""
    rank, F = 8, 128
    L = 70000 * F - 5
    ir = synth.decaying_ir(4, L)
    n = 40 * F
    x = synth.noise(4, n)
    b = pkg.ConvolverBatch(1, 0)
    assert b.init(0, ir, rank, 0.0)
    assert b.state(0)["bins"] == 70000 and b.state(0)["partitions"] == 70001
    out = np.concatenate([b.process(x[None, i:i + 5 * F])[0] for i in range(0, n, 5 * F)])
    assert rel_err(out, direct_convolve(x, ir, n)) <= TOL
    b.close()


def test_overlapped_launch_soak(pkg):
    """20 000 back-to-back blocks (cycling over a 512-block buffer): the overlapped pipeline
    stays bit-identical to the serialised one over a long run (wrap-around of the ring, the job
    counters and the hand-shake epochs)."""
    torch = pytest.importorskip("torch")
    n, rank, F, taps, frames, window = 6, 11, 1024, 37 * 1024 + 9, 20000, 512
    irs = [synth.decaying_ir(c, taps) for c in range(n)]
    g = torch.Generator(device="cuda").manual_seed(11)
    src = torch.rand((n, window * F), generator=g, device="cuda") * 2 - 1
    outs = []
    for pdl in (1, 0):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("pdl", pdl)
        for c in range(n):
            assert b.init(c, irs[c], rank, 0.0)
        dst = torch.zeros_like(src)
        for i in range(frames):
            k = i % window
            b.process_device(dst.data_ptr() + 4 * k * F, src.data_ptr() + 4 * k * F, window * F, F)
        b.sync()
        assert b.state(0)["frames"] == frames
        outs.append(dst.cpu().numpy())
        b.close()
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("n,taps,rank", [(8, 30000, 11), (64, 200000, 11), (3, 50000, 13), (16, 9000, 9)])
def test_cascaded_batches_on_one_stream(pkg, n, taps, rank):
    """Two batches cascaded on ONE caller stream (B.src == A.dst), launched back to back: B's launch
    may become resident while A's is still in its inverse-transform tail (programmatic dependent
    launch), so B must not read its input before every earlier launch has completed.  The
    overlapped run must equal the serialised one bit for bit, and be the right answer."""
    torch = pytest.importorskip("torch")
    F = 1 << (rank - 1)
    frames = 48
    ira = [synth.decaying_ir(c, taps) for c in range(2)]
    irb = [synth.decaying_ir(10 + c, taps // 2 + 7) for c in range(2)]
    g = torch.Generator(device="cuda").manual_seed(11)
    src = torch.rand((n, frames * F), generator=g, device="cuda") * 2 - 1
    st = torch.cuda.Stream()
    outs = []
    for pdl in (1, 0):
        A, B = pkg.ConvolverBatch(n, 0), pkg.ConvolverBatch(n, 0)
        for b, irs in ((A, ira), (B, irb)):
            b.set_option("pdl", pdl)
            for c in range(n):
                assert b.init(c, irs[c % 2], rank, 0.0)
        mid = torch.zeros((n, F), device="cuda")        # ONE block, rewritten every step
        dst = torch.zeros_like(src)
        torch.cuda.synchronize()
        for i in range(frames):
            A.process_device(mid.data_ptr(), src.data_ptr() + 4 * i * F, frames * F, F, st.cuda_stream,
                             dst_stride=F)
            B.process_device(dst.data_ptr() + 4 * i * F, mid.data_ptr(), F, F, st.cuda_stream,
                             dst_stride=frames * F)
        st.synchronize()
        outs.append(dst.cpu().numpy())
        A.close()
        B.close()
    assert np.array_equal(outs[0], outs[1])
    x = src[0].cpu().numpy().astype(np.float64)
    want = np.convolve(np.convolve(x, ira[0].astype(np.float64))[:frames * F], irb[0].astype(np.float64))[:frames * F]
    assert rel_err(outs[0][0], want) <= TOL


@pytest.mark.parametrize("n,taps,rank,own", [(8, 20000, 11, True), (64, 100000, 10, True), (8, 20000, 11, False)])
def test_early_input_transform_is_bit_identical_and_respects_overlap(pkg, n, taps, rank, own):
    """Option "early_src": on the batch's own stream (or, with value 2, on a caller's) a block's input
    transform may run before the previous block's tail has finished -- unless a launch in flight
    writes that input.  Three patterns per setting: independent buffers (early), feedback
    (src(t) = dst(t-1): must fall back to the in-kernel wait), and a long back-to-back run that
    crosses the bounded chain length.  All must equal the serialised (pdl = 0) run bit for bit."""
    torch = pytest.importorskip("torch")
    F = 1 << (rank - 1)
    frames = 70
    irs = [0.5 * synth.decaying_ir(c, taps) for c in range(2)]
    g = torch.Generator(device="cuda").manual_seed(21)
    src = torch.rand((n, frames * F), generator=g, device="cuda") * 2 - 1
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    results = []
    for pdl, early in ((1, 2 if not own else 1), (0, 0)):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("pdl", pdl)
        b.set_option("early_src", early)
        for c in range(n):
            assert b.init(c, irs[c % 2], rank, 0.0)
        stream = None if own else st.cuda_stream
        dst = torch.zeros_like(src)
        fb = torch.zeros((n, (frames + 1) * F), device="cuda")
        fb[:, :F] = src[:, :F]
        torch.cuda.synchronize()
        for i in range(frames):                             # independent buffers
            b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, frames * F, F, stream)
        for i in range(frames):                             # feedback
            b.process_device(fb.data_ptr() + 4 * (i + 1) * F, fb.data_ptr() + 4 * i * F, (frames + 1) * F, F, stream)
        b.sync()
        st.synchronize()
        results.append((dst.cpu().numpy(), fb.cpu().numpy()))
        b.close()
    assert np.array_equal(results[0][0], results[1][0])
    assert np.array_equal(results[0][1], results[1][1])
    want = direct_convolve(src[0].cpu().numpy(), irs[0], frames * F)
    assert rel_err(results[0][0][0], want) <= TOL


@pytest.mark.parametrize("n,taps,rank", [(8, 20000, 11), (64, 100000, 10)])
def test_batch_fed_its_own_previous_output(pkg, n, taps, rank):
    """Feedback: block t's input is block t-1's OUTPUT buffer (same stream, back to back).  The
    launch of block t must see the finished output of block t-1."""
    torch = pytest.importorskip("torch")
    F = 1 << (rank - 1)
    frames = 40
    irs = [0.5 * synth.decaying_ir(c, taps) for c in range(2)]
    g = torch.Generator(device="cuda").manual_seed(12)
    first = torch.rand((n, F), generator=g, device="cuda") * 2 - 1
    outs = []
    for pdl in (1, 0):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("pdl", pdl)
        for c in range(n):
            assert b.init(c, irs[c % 2], rank, 0.0)
        buf = torch.zeros((n, (frames + 1) * F), device="cuda")
        buf[:, :F] = first
        torch.cuda.synchronize()
        for i in range(frames):
            b.process_device(buf.data_ptr() + 4 * (i + 1) * F, buf.data_ptr() + 4 * i * F, (frames + 1) * F, F)
        b.sync()
        outs.append(buf.cpu().numpy())
        b.close()
    assert np.array_equal(outs[0], outs[1])
    # float64 model of the recursion for channel 0
    h = irs[0].astype(np.float64)
    x = np.zeros(frames * F)
    cur = first[0].cpu().numpy().astype(np.float64)
    got = outs[0][0]
    hist = np.zeros(0)
    for i in range(3):                                  # three blocks are enough to pin the data flow
        hist = np.concatenate([hist, cur])
        y = np.convolve(hist, h)[i * F:(i + 1) * F]
        assert rel_err(got[(i + 1) * F:(i + 2) * F], y) <= 10 * TOL
        cur = got[(i + 1) * F:(i + 2) * F].astype(np.float64)


@pytest.mark.parametrize("taps,rank,phase,calls", [
    (65536, 11, 0.0, [1024] * 3 + [100, 28, 128, 768, 1, 1023]),
    (5000, 9, 0.37, [31] * 40),
    (70000, 13, 0.9, [333] * 30),
    (200, 12, 0.0, [256, 256, 1536, 2048, 7]),
    (129, 8, 0.5, [64, 64, 128, 1, 127]),
    (300000, 16, 0.0, [31, 4096, 28641, 32768, 5]),
    (31, 9, 0.0, [31] * 10),
])
def test_dump_fields_match_the_reference_scheduler(pkg, taps, rank, phase, calls):
    """Convolver::dump (Convolver.cpp:315-337): the 11 scalar fields the engine reports must equal
    the reference scheduler's after the same init / process history -- including nBlocksDone, which
    the reference raises at 128-sample boundaries and the engine reproduces arithmetically."""
    names = {"nDataBufferSize": "data_buffer_size", "nDirectSize": "direct_size", "nFrameSize": "frame_size",
             "nFrameOff": "frame_off", "nConvSize": "conv_size", "nLevels": "levels", "nBlocks": "blocks",
             "nBlocksDone": "blocks_done", "nRank": "rank", "nBlkInit": "blk_init", "fBlkCoef": "blk_coef"}
    ir = synth.decaying_ir(1, taps)
    ref = CpuConvolver("oracle")
    b = pkg.ConvolverBatch(1, 0)
    d = b.dump(0)
    assert all(not d[k] for k in d)                     # construct(): NULL / 0
    assert ref.init(ir, rank, phase) and b.init(0, ir, rank, phase)

    def same(where):
        want, got = ref.state(), b.dump(0)
        for k, o in names.items():
            assert got[k] == want[o], (where, k, got[k], want[o])
        assert all(got[k] for k in ("vDataBuffer", "vFrame", "vTaskData", "vConvData", "vDirectData", "vData"))

    same("after init")
    x = synth.noise(3, sum(calls))
    pos = 0
    for i, n in enumerate(calls):
        ref.process(x[pos:pos + n])
        b.process(x[None, pos:pos + n])
        pos += n
        same("after call %d (%d samples)" % (i, n))
    b.close()


def test_init_many_equals_one_by_one(pkg):
    """b200conv_init_many (one slab, one staged upload, one transform launch for all instances) must
    leave exactly the state N x b200conv_init leaves: bit-identical outputs, ragged lengths, phases,
    partition-range shards, a zero-length entry that destroys its instance, and re-initialisation of a
    subset while the other instances keep their history."""
    rank, F, n = 10, 512, 7
    lens = [30000, 1, 511, 513, 0, 70000, 4096]
    phases = [0.0, 0.5, 0.25, 0.0, 0.0, 0.9, 0.0]
    offs = [0, 0, 0, 0, 0, 0, 3]
    irs = [synth.decaying_ir(c, max(L, 1))[:L] for c, L in enumerate(lens)]
    x = np.stack([synth.noise(80 + c, 9 * F + 77) for c in range(n)])
    a, b = pkg.ConvolverBatch(n, 0), pkg.ConvolverBatch(n, 0)
    assert a.init(4, synth.decaying_ir(9, 100), rank, 0.0) and b.init(4, synth.decaying_ir(9, 100), rank, 0.0)
    for c in range(n):
        assert a.init(c, irs[c], rank, phases[c], part_offset=offs[c])
    assert b.init_many(list(range(n)), irs, rank, phases, offs)
    assert a.rank(4) == 0 and b.rank(4) == 0            # the zero-length entry destroyed instance 4
    for c in range(n):
        assert a.state(c) == b.state(c)
    step = 300
    for i in range(0, x.shape[1], step):
        ya, yb = a.process(x[:, i:i + step].copy()), b.process(x[:, i:i + step].copy())
        assert np.array_equal(ya, yb)
    # re-initialise two instances in one call: the others keep their history
    new = [synth.decaying_ir(50, 2000), synth.decaying_ir(51, 9000)]
    for c, ir in zip((2, 5), new):
        assert a.init(c, ir, rank, 0.0)
    assert b.init_many([2, 5], new, rank)
    for i in range(0, 4 * F, step):
        blk = x[:, i:i + step].copy()
        assert np.array_equal(a.process(blk), b.process(blk))
    # a rank clash fails as a whole and changes nothing
    with pytest.raises(pkg.B200ConvError):
        b.init_many([0], [irs[0]], rank + 1)
    assert b.rank(0) == rank
    a.close()
    b.close()


def test_switching_between_pipelined_profiled_and_synchronous_calls(pkg):
    """The sequence bench.py runs on one batch: back-to-back device calls on a caller's stream
    (pipelined launches), a profiled pass (every launch bracketed by events: not pipelined), and
    then synchronous host calls on page-locked buffers (the batch's own stream, eager pending MAC).
    The slot bookkeeping of the pipelined launches must survive every switch: outputs stay
    correct and no in-kernel wait times out."""
    torch = pytest.importorskip("torch")
    n, taps, rank, F = 8, 60000, 11, 1024
    irs = [synth.decaying_ir(c, taps) for c in range(2)]
    b = pkg.ConvolverBatch(n, 0)
    for c in range(n):
        assert b.init(c, irs[c % 2], rank, 0.0)
    blocks = 30
    x = np.stack([synth.noise(300 + c, 4 * blocks * F) for c in range(n)])
    src = torch.from_numpy(x).cuda()
    dst = torch.zeros_like(src)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    pos = 0

    def device_calls(count, stream):
        nonlocal pos
        for _ in range(count):
            b.process_device(dst.data_ptr() + 4 * pos, src.data_ptr() + 4 * pos, src.shape[1], F, stream)
            pos += F

    device_calls(blocks, st.cuda_stream)
    st.synchronize()
    b.set_profiling(True)
    device_calls(blocks // 2, st.cuda_stream)
    torch.cuda.synchronize()
    ms, launches = b.profile()
    assert launches == blocks // 2 and ms > 0
    b.set_profiling(False)
    device_calls(blocks // 2, None)                      # the batch's own stream
    b.sync()
    out = dst.cpu().numpy()
    hs = torch.empty((n, F)).pin_memory()
    hd = torch.empty((n, F)).pin_memory()
    for _ in range(blocks):                             # synchronous calls on page-locked buffers
        hs.copy_(torch.from_numpy(x[:, pos:pos + F]))
        b.process(hs.numpy(), hd.numpy())
        out[:, pos:pos + F] = hd.numpy()
        pos += F
    device_calls(blocks, None)
    b.sync()
    out[:, pos - blocks * F:pos] = dst[:, pos - blocks * F:pos].cpu().numpy()
    assert not b.reduce_timed_out()
    for c in (0, 5):
        want = direct_convolve(x[c], irs[c % 2], pos)
        assert rel_err(out[c, :pos], want) <= TOL
    b.close()


@pytest.mark.parametrize("n,taps,rank,step,phases", [(2, 60000, 9, 256, (0.0, 0.5)), (8, 30000, 10, 77, (0.0, 0.37)),
                                                     (3, 20000, 12, 1500, (0.25,)), (70, 9000, 8, 100, (0.0, 0.6))])
def test_spread_direct_form_answers_in_place_on_device(pkg, n, taps, rank, step, phases):
    """The job-list launch shares the direct-form answers of a call (the samples that finish the frame
    in progress, the first samples of the next one) among the CTAs of the job; the CTA that finishes
    the job copies them out after every CTA has read the caller's samples.  So dst == src (in place,
    device pointers, back-to-back launches with programmatic serialisation) must still be exact."""
    torch = pytest.importorskip("torch")
    calls = 60
    irs = [synth.decaying_ir(c, taps) for c in range(2)]
    x = np.stack([synth.noise(500 + c, calls * step) for c in range(n)])
    outs = []
    for pdl in (1, 0):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("pdl", pdl)
        for c in range(n):
            assert b.init(c, irs[c % 2], rank, phases[c % len(phases)])
        buf = torch.from_numpy(x).cuda()
        torch.cuda.synchronize()
        for i in range(calls):
            p = buf.data_ptr() + 4 * i * step
            b.process_device(p, p, calls * step, step)
        b.sync()
        outs.append(buf.cpu().numpy())
        b.close()
    assert np.array_equal(outs[0], outs[1])
    for c in (0, 1, n - 1):
        assert rel_err(outs[0][c], direct_convolve(x[c], irs[c % 2], calls * step)) <= TOL


def test_cascade_behind_a_batch_on_the_job_list_path(pkg):
    """Batch A is called with sizes that take the job-list path (its launches let their successor
    start early but are not in the launch history); batch B, on the same caller stream and allowed to
    transform its input early ("early_src" = 2), reads A's output block: B's launch must not start
    before A's has completed.  Equal to the serialised run bit for bit, and the right answer."""
    torch = pytest.importorskip("torch")
    n, rank, F = 4, 11, 1024
    ira = [synth.decaying_ir(c, 30000) for c in range(2)]
    irb = [synth.decaying_ir(10 + c, 15000) for c in range(2)]
    frames = 40
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.rand((n, frames * F), generator=g, device="cuda") * 2 - 1
    st = torch.cuda.Stream()
    outs = []
    for pdl in (1, 0):
        A, B = pkg.ConvolverBatch(n, 0), pkg.ConvolverBatch(n, 0)
        for b, irs in ((A, ira), (B, irb)):
            b.set_option("pdl", pdl)
            b.set_option("early_src", 2 if pdl else 0)
            for c in range(n):
                assert b.init(c, irs[c % 2], rank, 0.0)
        mid = torch.zeros((n, F), device="cuda")
        dst = torch.zeros_like(src)
        torch.cuda.synchronize()
        for i in range(frames):
            # A: the block in two uneven calls (job-list path), B: the whole block (one k_frame launch)
            A.process_device(mid.data_ptr(), src.data_ptr() + 4 * i * F, frames * F, 300, st.cuda_stream, dst_stride=F)
            A.process_device(mid.data_ptr() + 4 * 300, src.data_ptr() + 4 * (i * F + 300), frames * F, F - 300,
                             st.cuda_stream, dst_stride=F)
            B.process_device(dst.data_ptr() + 4 * i * F, mid.data_ptr(), F, F, st.cuda_stream, dst_stride=frames * F)
        st.synchronize()
        outs.append(dst.cpu().numpy())
        A.close()
        B.close()
    assert np.array_equal(outs[0], outs[1])
    x = src[1].cpu().numpy().astype(np.float64)
    want = np.convolve(np.convolve(x, ira[1].astype(np.float64))[:frames * F], irb[1].astype(np.float64))[:frames * F]
    assert rel_err(outs[0][1], want) <= TOL


def test_smaller_mac_bin_tiles(pkg):
    """Developer option "mac_tile" (ranks >= 14): 512-bin k_mac tiles with two partitions per stage
    (one bulk copy per row).  Not the bits of the default 1024-bin tiles -- the stage that holds
    partition 0 is taken last, so two partitions per stage change the order of the fp32 sum -- but the
    same answer to rounding."""
    torch = pytest.importorskip("torch")
    n, taps, rank = 3, 150000, 14
    F = 1 << (rank - 1)
    irs = [synth.decaying_ir(c, taps) for c in range(2)]
    x = np.stack([synth.noise(700 + c, 6 * F) for c in range(n)])
    outs = []
    try:
        for tile in (0, 512):
            b = pkg.ConvolverBatch(n, 0)
            b.set_option("mac_tile", tile)
            for c in range(n):
                assert b.init(c, irs[c % 2], rank, 0.0)
            src = torch.from_numpy(x).cuda()
            dst = torch.zeros_like(src)
            for i in range(6):
                b.process_device(dst.data_ptr() + 4 * i * F, src.data_ptr() + 4 * i * F, 6 * F, F)
            b.sync()
            outs.append(dst.cpu().numpy())
            b.close()
    finally:
        b = pkg.ConvolverBatch(1, 0)
        b.set_option("mac_tile", 0)                     # process-wide knob: back to the default
        b.close()
    for c in range(n):
        want = direct_convolve(x[c], irs[c % 2], 6 * F)
        assert rel_err(outs[0][c], want) <= TOL
        assert rel_err(outs[1][c], want) <= TOL
        assert rel_err(outs[1][c], outs[0][c].astype(np.float64)) <= 2e-6


@pytest.mark.parametrize("rank", [14, 15])
def test_mac_launched_ahead_survives_other_call_sizes_and_reinit(pkg, rank):
    """Ranks 14..16, whole-frame calls: the MAC of block t + 1 is launched behind block t (option
    "chain_ahead").  A block that then does not come as a whole frame (partial calls, a multi-frame
    call), or comes after an instance was re-initialised, must not use those rows; the answer is
    always the convolution, and equal to rounding to the schedule without the look-ahead."""
    torch = pytest.importorskip("torch")
    F = 1 << (rank - 1)
    n, taps = 3, 5 * F + 77
    irs = [synth.decaying_ir(c, taps) for c in range(3)]
    sizes = [F, F, 100, F - 100, F, 2 * F, F, F, 7, F, F - 7, F, F]
    total = sum(sizes)
    x = np.stack([synth.noise(900 + c, total) for c in range(n)])
    outs = []
    for ahead in (1, 0):
        b = pkg.ConvolverBatch(n, 0)
        b.set_option("chain_ahead", ahead)
        for c in range(n):
            assert b.init(c, irs[c], rank, 0.0)
        src = torch.from_numpy(x).cuda()
        dst = torch.zeros_like(src)
        pos = 0
        for s in sizes:
            b.process_device(dst.data_ptr() + 4 * pos, src.data_ptr() + 4 * pos, total, s)
            pos += s
        b.sync()
        out = dst.cpu().numpy()
        for c in range(n):
            assert rel_err(out[c], direct_convolve(x[c], irs[c], total)) <= TOL, (ahead, c)
        # re-initialise one instance between two whole-frame calls: its history restarts, the others carry on
        assert b.init(1, irs[2], rank, 0.0)
        tail = np.stack([synth.noise(950 + c, 3 * F) for c in range(n)])
        src2 = torch.from_numpy(tail).cuda()
        dst2 = torch.zeros_like(src2)
        for i in range(3):
            b.process_device(dst2.data_ptr() + 4 * i * F, src2.data_ptr() + 4 * i * F, 3 * F, F)
        b.sync()
        out2 = dst2.cpu().numpy()
        assert rel_err(out2[1], direct_convolve(tail[1], irs[2], 3 * F)) <= TOL
        want0 = direct_convolve(np.concatenate([x[0], tail[0]]), irs[0], total + 3 * F)[total:]
        assert rel_err(out2[0], want0) <= TOL
        outs.append(np.concatenate([out, out2], axis=1))
        b.close()
    assert rel_err(outs[0][2], outs[1][2].astype(np.float64)) <= 2e-6
