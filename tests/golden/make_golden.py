"""Generates tests/golden/convolver_golden.npz from the reference's own Convolver.cpp.

Run in the build container (needs /root/reference to build oracle/_ref):

    python tests/golden/make_golden.py

Each case stores the IR, the input, the call pattern and the output produced by
oracle/_ref/libref_convolver.so (reference src/main/util/Convolver.cpp compiled verbatim over the
restated lsp::dsp:: kernels).  The reference ships no golden vectors for this path;
these fixtures freeze its behaviour as observed here so that the GPU box, where /root/reference
does not exist, can still check against it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import bindings  # noqa: E402
import synth  # noqa: E402

#        name            taps   rank  phase  step   samples
CASES = [("r8_phase",     5000,  8,   0.25,  77,    6000),
         ("r11_aligned",  6000,  11,  0.0,   1024,  8192),
         ("r9_halfphase", 3000,  9,   0.5,   256,   4096),
         ("r12_short_ir", 200,   12,  0.0,   333,   5000),
         ("clamp_low",    300,   3,   0.0,   100,   1000),
         ("clamp_high",   300,   20,  0.9,   100,   1000)]


def main():
    bindings.build()
    out = {}
    cases = []
    ir, src = synth.utest_small()
    cases.append(("utest_small", ir, src, 9, 0.0, 31))
    ir, src = synth.utest_large()
    cases.append(("utest_large", ir, src, 10, 0.0, 31))
    for k, (name, taps, rank, phase, step, n) in enumerate(CASES):
        cases.append((name, synth.decaying_ir(100 + k, taps), synth.noise(100 + k, n), rank, phase, step))

    for name, ir, src, rank, phase, step in cases:
        c = bindings.CpuConvolver("reference")
        assert c.init(ir, rank, phase)
        out[name + ".ir"] = ir
        out[name + ".src"] = src
        out[name + ".dst"] = c.run(src, step)
        out[name + ".meta"] = np.array([rank, phase, step, c.rank(), c.data_size()], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "convolver_golden.npz"), **out)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
