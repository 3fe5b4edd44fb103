"""Float64 model of SyncChirpProcessor::do_linear_convolutions (reference
src/main/util/SyncChirpProcessor.cpp:1224-1250,1299-1331,1374-1508), independent of the fastconv
primitives: per channel the linear convolution of the input (zero padded at the tail to
vPaddedLengths) with the inverse filter prepended by vInversePrepends zeros, written at
vAlignOffsets, then scaled over the FIRST vConvLengths samples of the row (dsp::mul_k2 at :1508
starts at index 0, not at the align offset)."""
import numpy as np

MAX_PART_SIZE = 32768


def plan(in_len, inverse_len, part_size_limit):
    limit = min(part_size_limit, MAX_PART_SIZE) or MAX_PART_SIZE
    part, exponent = 1, 0
    while part < limit:
        part <<= 1
        exponent += 1
    partitions = [max(n, inverse_len) // part + 1 for n in in_len]
    padded = [p * part for p in partitions]
    conv = [2 * p for p in padded]
    alloc = max(conv)
    middle = alloc // 2 - 1
    return {"partition_size": part, "conv_rank": exponent + 1, "image": 1 << (exponent + 2),
            "allocation_size": alloc, "partitions": partitions, "padded": padded,
            "prepends": [p - inverse_len for p in padded], "conv_lengths": conv,
            "align_offsets": [middle - c // 2 + 1 for c in conv]}


def linear_convolutions(inputs, inverse, part_size_limit, scale):
    from scipy.signal import fftconvolve
    inverse = np.asarray(inverse, np.float64)
    pl = plan([len(x) for x in inputs], inverse.size, part_size_limit)
    out = np.zeros((len(inputs), pl["allocation_size"]))
    for ch, x in enumerate(inputs):
        x = np.asarray(x, np.float64)
        pre = np.concatenate([np.zeros(pl["prepends"][ch]), inverse])
        y = fftconvolve(x, pre) if x.size * pre.size > (1 << 22) else np.convolve(x, pre)
        a = pl["align_offsets"][ch]
        n = min(y.size, pl["allocation_size"] - a)
        out[ch, a:a + n] = y[:n]
        out[ch, :pl["conv_lengths"][ch]] *= scale
    return out, pl
