"""Seeded synthetic signals shared by the parity tests and bench.py.

* input  : white noise, uniform [-1, 1), float32, PCG64(0x5EED0000 + c)
* IR     : h[n] = g * u[n] * exp(-n / tau), u uniform [-1, 1) from PCG64(0x1A000000 + c),
           tau = L / ln(1000) (-60 dB at the end), g such that sum(h^2) = 1
"""
import numpy as np


def noise(c, n):
    return np.random.Generator(np.random.PCG64(0x5EED0000 + c)).uniform(-1.0, 1.0, n).astype(np.float32)


def decaying_ir(c, taps):
    u = np.random.Generator(np.random.PCG64(0x1A000000 + c)).uniform(-1.0, 1.0, taps)
    h = u * np.exp(-np.arange(taps) / (taps / np.log(1000.0)))
    h /= np.sqrt(np.sum(h * h))
    return h.astype(np.float32)


def utest_small():
    """Signals of the reference's test_small (src/test/utest/util/convolver.cpp:88-112)."""
    ir = np.arange(1, 32, dtype=np.float32)
    src = np.zeros(0x2000 + ir.size, dtype=np.float32)
    vals = (1.0, 0.1, 0.01)
    for j, i in enumerate(range(0, 0x2000, 5)):
        src[i] = vals[j % 3]
    return ir, src


def utest_large(seed=7):
    """Signals of the reference's test_large (convolver.cpp:184-203): FloatBuffer's random fill is
    replaced by a seeded generator."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ir = rng.uniform(-1.0, 1.0, 0x2000).astype(np.float32)
    src = np.zeros(0x20 + ir.size, dtype=np.float32)
    src[:0x20] = rng.uniform(-1.0, 1.0, 0x20).astype(np.float32)
    return ir, src


def equals_relative(a, b, tol):
    """lsp-test-fw's FloatBuffer::equals_relative as used at convolver.cpp:123 (recalled: the
    test framework is not in the reference tree): when either value is exactly zero the other
    must be below `tol` in magnitude, otherwise |a/b - 1| < tol."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    zero = (a == 0.0) | (b == 0.0)
    ok_zero = np.maximum(np.abs(a), np.abs(b)) < tol
    with np.errstate(divide="ignore", invalid="ignore"):
        ok_rel = np.abs(a / np.where(b == 0.0, 1.0, b) - 1.0) < tol
    return bool(np.all(np.where(zero, ok_zero, ok_rel)))
