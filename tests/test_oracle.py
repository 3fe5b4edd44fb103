"""Pins the CPU oracle (oracle/) before anything is checked against it.

Mirrors the reference's own unit test src/test/utest/util/convolver.cpp (same shapes, same
tolerances, same identity: naive direct convolution), then checks the plain-C restatement
bit-for-bit against the reference's Convolver.cpp compiled verbatim (oracle/_ref) and against the
committed fixtures generated from that build.
"""
import numpy as np
import pytest

import synth
from oracle import bindings
from oracle.bindings import CpuConvolver, direct_convolve

HAVE_REF = CpuConvolver.available("reference")
IMPLS = ["oracle"] + (["reference"] if HAVE_REF else [])


equals_relative = synth.equals_relative


@pytest.mark.parametrize("impl", IMPLS)
def test_small(impl):
    # convolver.cpp:88-136 : 31-tap ramp IR, sparse input, rank 9, calls of 31, rel 1e-4
    ir, src = synth.utest_small()
    c = CpuConvolver(impl)
    assert c.init(ir, 9, 0.0)
    out = c.run(src, 31)
    assert equals_relative(out, direct_convolve(src, ir, src.size), 1e-4)
    c.destroy()
    assert c.data_size() == 0 and c.rank() == 0


@pytest.mark.parametrize("impl", IMPLS)
def test_large(impl):
    # convolver.cpp:184-223 : random 8192-tap IR, 32 random samples then zeros, rank 10, abs 1e-4
    ir, src = synth.utest_large()
    c = CpuConvolver(impl)
    assert c.init(ir, 10, 0.0)
    out = c.run(src, 31)
    assert np.max(np.abs(out - direct_convolve(src, ir, src.size))) <= 1e-4


@pytest.mark.parametrize("impl", IMPLS)
def test_collisions_subsampled(impl):
    # convolver.cpp:138-182 (disabled upstream; all 65535 offsets): 65536-tap random IR, rank 10,
    # impulses at 0 and i, calls of 127, tail flushed with data_size()-1 zeros, abs 1e-5.
    rng = np.random.Generator(np.random.PCG64(11))
    L = 0x10000
    ir = rng.uniform(-1.0, 1.0, L).astype(np.float32)
    for i in (1, 127, 128, 511, 512, 513, 4095, 40000, L - 1):
        c = CpuConvolver(impl)
        assert c.init(ir, 10, 0.0)
        src = np.zeros(L + c.data_size() - 1, dtype=np.float32)   # convolve_full: input + flush
        src[0] = 1.0
        src[i] = 1.0
        out = c.run(src, 127)
        want = direct_convolve(src[:L], ir)
        assert np.max(np.abs(out - want[:out.size])) <= 1e-5, i


@pytest.mark.parametrize("impl", IMPLS)
def test_api_contract(impl):
    c = CpuConvolver(impl)
    x = synth.noise(0, 300)
    # not initialised -> zeros (Convolver.cpp:219-223)
    assert not c.process(x).any()
    # rank is clamped to [8, 16] (Convolver.cpp:87)
    assert c.init(np.ones(10, np.float32), 3, 0.0) and c.rank() == 8
    assert c.init(np.ones(10, np.float32), 20, 0.0) and c.rank() == 16
    assert c.data_size() == 10
    # count == 0 -> destroy() and true (Convolver.cpp:80-84)
    assert c.init(np.zeros(0, np.float32), 10, 0.0)
    assert c.rank() == 0 and c.data_size() == 0
    assert not c.process(x).any()
    # count == 0 in process is a no-op
    c.init(np.ones(4, np.float32), 8, 0.0)
    assert c.process(np.zeros(0, np.float32)).size == 0


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("taps,rank,phase,step", [(65536, 11, 0.0, 1024), (65536, 11, 0.37, 1000),
                                                  (5000, 8, 0.0, 77), (5000, 16, 0.5, 4096),
                                                  (70000, 13, 0.9, 333), (200, 12, 0.0, 256),
                                                  (129, 9, 0.0, 64), (1, 8, 0.0, 5)])
def test_zero_latency_any_phase_any_call_size(impl, taps, rank, phase, step):
    ir = synth.decaying_ir(3, taps)
    src = synth.noise(3, min(3 * taps + 500, 90000))
    c = CpuConvolver(impl)
    assert c.init(ir, rank, phase)
    out = c.run(src, step)
    want = direct_convolve(src, ir, src.size)
    assert np.max(np.abs(out - want)) / np.max(np.abs(want)) <= 1e-5


def test_inplace_process_is_safe():
    # dst == src allowed (Convolver.cpp:291 copies src before :296 writes dst)
    ir = synth.decaying_ir(5, 3000)
    src = synth.noise(5, 4000)
    a = CpuConvolver("oracle"); a.init(ir, 9, 0.0)
    b = CpuConvolver("oracle"); b.init(ir, 9, 0.0)
    want = a.run(src, 100)
    buf = src.copy()
    for i in range(0, buf.size, 100):
        b.process(buf[i:i + 100], out=buf[i:i + 100])
    assert np.array_equal(buf, want)


def test_schedule_constants():
    # SURVEY section 8 table / App. A.3 (Convolver.cpp:199-210): nBlocks, nLevels and the per-sub-step
    # targets min(nBlocks, size_t(nBlkInit + fBlkCoef * sub_id)).
    for taps, rank, bins, blocks, levels, sched in [
            (65536, 11, 64, 63, 3, [1, 9, 9, 9, 9, 9, 9, 8]),
            (480000, 11, 469, 468, 3, [1, 66, 67, 67, 67, 67, 67, 66]),
            (48000, 11, 47, 46, 3, [1, 6, 7, 6, 7, 6, 7, 6]),
            (192000, 9, 750, 749, 1, [1, 748])]:
        c = CpuConvolver("oracle")
        assert c.init(np.ones(taps, np.float32) * 1e-3, rank, 0.0)
        s = c.state()
        assert (s["blocks"], s["levels"], s["frame_size"]) == (blocks, levels, 1 << (rank - 1))
        assert s["data_buffer_size"] == (bins + 1) << (rank - 1)
        done, steps = 0, []
        for sub in range(len(sched)):
            tgt = min(blocks, int(np.float32(s["blk_init"]) + np.float32(s["blk_coef"]) * np.float32(sub)))
            steps.append(tgt - done)
            done = tgt
        assert steps == sched and done == blocks


def test_phase_sets_frame_offset():
    # nFrameOff = size_t(phase * F) % F (Convolver.cpp:140)
    for rank, phase in [(9, 0.5), (11, 0.37), (8, 0.999), (10, 1.25)]:
        c = CpuConvolver("oracle")
        c.init(np.ones(5, np.float32), rank, phase)
        F = 1 << (rank - 1)
        assert c.state()["frame_off"] == int(np.float32(phase) * np.float32(F)) % F


def test_primitives_contract():
    """parse -> apply accumulates, parse_apply == parse + apply, restore stores."""
    import ctypes
    lib = CpuConvolver.lib("oracle")[0]
    fp = ctypes.POINTER(ctypes.c_float)
    P = lambda a: a.ctypes.data_as(fp)
    rng = np.random.Generator(np.random.PCG64(5))
    for rank in (1, 2, 3, 8, 11, 16):
        n = 1 << rank
        a = rng.uniform(-1, 1, n // 2).astype(np.float32)
        b = rng.uniform(-1, 1, n // 2).astype(np.float32)
        ia, ib, tmp = (np.zeros(2 * n, np.float32) for _ in range(3))
        lib.rs_fastconv_parse(P(ia), P(a), ctypes.c_size_t(rank))
        lib.rs_fastconv_parse(P(ib), P(b), ctypes.c_size_t(rank))
        want = np.convolve(a.astype(np.float64), b.astype(np.float64))
        want = np.concatenate([want, np.zeros(n - want.size)])
        tol = 2e-6 * max(1.0, np.abs(want).max())
        dst = np.ones(n, np.float32)            # apply must ADD to what is there
        lib.rs_fastconv_apply(P(dst), P(tmp), P(ia), P(ib), ctypes.c_size_t(rank))
        assert np.max(np.abs(dst - 1.0 - want)) <= tol
        dst2 = np.ones(n, np.float32)
        lib.rs_fastconv_parse_apply(P(dst2), P(tmp), P(ia), P(b), ctypes.c_size_t(rank))
        assert np.max(np.abs(dst2 - 1.0 - want)) <= tol
        # restore(parse(a)) == [a, 0...] and overwrites
        dst3 = np.full(n, 7.0, np.float32)
        img = ia.copy()
        lib.rs_fastconv_restore(P(dst3), P(img), ctypes.c_size_t(rank))
        assert np.max(np.abs(dst3[:n // 2] - a)) <= 1e-6 and np.max(np.abs(dst3[n // 2:])) <= 1e-6


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
def test_restatement_matches_verbatim_reference():
    """Same kernels below, so the restated scheduler must reproduce the reference's bit for bit."""
    rng = np.random.Generator(np.random.PCG64(2))
    for taps, rank, phase, step, n in [(31, 9, 0.0, 31, 8000), (8192, 10, 0.0, 31, 9000),
                                       (65536, 11, 0.0, 1024, 40000), (65536, 11, 0.37, 1000, 30000),
                                       (5000, 8, 0.0, 77, 12000), (5000, 16, 0.5, 4096, 70000),
                                       (70000, 13, 0.9, 333, 80000), (129, 9, 0.0, 64, 2000),
                                       (192000, 9, 0.5, 256, 20000)]:
        ir = rng.uniform(-1, 1, taps).astype(np.float32)
        src = rng.uniform(-1, 1, n).astype(np.float32)
        a, b = CpuConvolver("oracle"), CpuConvolver("reference")
        assert a.init(ir, rank, phase) and b.init(ir, rank, phase)
        assert (a.rank(), a.data_size()) == (b.rank(), b.data_size())
        assert np.array_equal(a.run(src, step), b.run(src, step))


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
def test_reference_dump_fields():
    # Convolver::dump writes 18 named fields (Convolver.cpp:315-337)
    c = CpuConvolver("reference")
    c.init(np.ones(300, np.float32), 9, 0.0)
    n, names = c.dump_names()
    assert n == 18 and names[0] == "pDataBuffer" and names[-1] == "vData" and "fBlkCoef" in names


def test_golden_fixtures(golden):
    """The restatement reproduces the outputs frozen from the verbatim reference build."""
    names = sorted({k.split(".")[0] for k in golden.files})
    assert len(names) == 8
    for name in names:
        rank, phase, step, eff_rank, size = golden[name + ".meta"]
        c = CpuConvolver("oracle")
        assert c.init(golden[name + ".ir"], int(rank), float(phase))
        assert (c.rank(), c.data_size()) == (int(eff_rank), int(size))
        out = c.run(golden[name + ".src"], int(step))
        assert np.array_equal(out, golden[name + ".dst"]), name


def test_cpu_bench_driver_runs():
    rate, sec = bindings.cpu_bench("oracle", 2, 4096, 11, 1024, 1, 4, 2)
    assert rate > 0 and sec > 0
