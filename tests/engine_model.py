"""Executable numpy model of the B200 engine's arithmetic (TEST INFRASTRUCTURE).

This mirrors, index for index, what the CUDA kernels in lsp-dsp-units_b200/csrc do, so the index
algebra can be checked on a CPU-only box before GPU time is spent:

* fwd_half_spectrum  = kernel K1: two P-point Stockham FFTs (even / odd bins of the packed
                       M-point transform of a half-zero frame) + real-FFT split post-pass,
                       output = M packed complex bins, (DC, Nyquist) folded into bin 0;
* mac                = kernel K2: complex multiply-accumulate over IR partitions with the bin-0
                       special case, split over partition chunks into partial rows;
* inv_first_half     = kernel K3: merge pre-pass, two P-point inverse FFTs, pruned combine that
                       yields only the first F of the 2F output samples;
* ModelConvolver     = host scheduler: frame bookkeeping, ring slots, partial-frame path.

The IR spectra are stored in the "folded overlap" form G_q = H_q + (-1)^k H_{q-1}
(q = 0..bins, H_{-1} = H_bins = 0): the second half of every partition product is aliased onto the
first half of the next frame inside the frequency-domain accumulator, so one inverse transform per
frame gives the finished output block and no time-domain overlap tail has to be kept.
"""
import numpy as np

C64 = np.complex64
F32 = np.float32


def twiddle_table(N):
    """w[i] = exp(-2*pi*i/N), i < N, computed in double and rounded to fp32 (as the device table)."""
    a = -2.0 * np.pi * np.arange(N) / N
    return (np.cos(a).astype(F32) + 1j * np.sin(a).astype(F32)).astype(C64)


def stockham(buf, tw, N, inverse):
    """In-place-with-register-staging Stockham FFT of len(buf)=P points, radix 4 (+ one leading
    radix-2 pass when log2 P is odd).  tw is the N-point table; P divides N."""
    P = buf.size
    logp = P.bit_length() - 1
    x = buf.astype(C64).copy()
    Ns = 1
    if logp & 1:
        # radix-2 first pass, Ns = 1: no twiddles
        j = np.arange(P // 2)
        a, b = x[j], x[j + P // 2]
        y = np.empty_like(x)
        y[2 * j] = a + b
        y[2 * j + 1] = a - b
        x = y
        Ns = 2
    while Ns < P:
        j = np.arange(P // 4)
        k = j % Ns
        step = N // (4 * Ns)                      # table stride for exp(-2 pi i k / (4 Ns))
        v = [x[j + r * (P // 4)] for r in range(4)]
        for r in (1, 2, 3):
            w = tw[(k * r * step) % N]
            if inverse:
                w = np.conj(w)
            v[r] = (v[r] * w).astype(C64)
        # radix-4 butterfly
        s0, s1 = v[0] + v[2], v[0] - v[2]
        s2, s3 = v[1] + v[3], v[1] - v[3]
        rot = (1j if inverse else -1j) * s3       # forward: -i * (v1 - v3)
        o = [s0 + s2, s1 + rot, s0 - s2, s1 - rot]
        j0 = (j // Ns) * (4 * Ns) + k
        y = np.empty_like(x)
        for r in range(4):
            y[j0 + r * Ns] = o[r].astype(C64)
        x = y
        Ns *= 4
    return x


def fwd_half_spectrum(frame, rank, tw):
    """K1.  frame: F real samples (zero padding to 2F implied).  Returns M=F packed complex bins."""
    N = 1 << rank
    M = N // 2
    P = M // 2
    f = frame.astype(F32)
    z = (f[0::2] + 1j * f[1::2]).astype(C64)                   # P points
    A = stockham(z, tw, N, False)                               # even bins of Z
    B = stockham((z * tw[2 * np.arange(P)]).astype(C64), tw, N, False)   # odd bins (pre-twiddle w_M^m)
    X = np.empty(M, C64)

    def Z(k):
        k = k % M
        return A[k // 2] if (k & 1) == 0 else B[k // 2]

    for k in range(0, M // 2 + 1):
        zk, zm = Z(k), Z(M - k)
        if k == 0:
            X[0] = (zk.real + zk.imag) + 1j * (zk.real - zk.imag)
            continue
        e = F32(0.5) * (zk + np.conj(zm))
        o = F32(0.5) * (zk - np.conj(zm))
        X[k] = C64(e - 1j * tw[k] * o)
        if k != M - k:
            # X[M-k] = conj(e) - i w^(M-k) * (-conj(o)),  w^(M-k) = -conj(w^k)
            X[M - k] = C64(np.conj(e) - 1j * np.conj(tw[k]) * np.conj(o))
    return X


def inv_first_half(Y, rank, tw):
    """K3.  Y: M packed bins.  Returns the first F samples of the 2F-point inverse real FFT."""
    N = 1 << rank
    M = N // 2
    P = M // 2
    A = np.empty(P, C64)
    B = np.empty(P, C64)

    def put(k, val):
        if (k & 1) == 0:
            A[k // 2] = val
        else:
            B[k // 2] = val

    for k in range(0, M // 2 + 1):
        if k == 0:
            dc, ny = Y[0].real, Y[0].imag
            put(0, C64((dc + ny) + 1j * (dc - ny)))
            continue
        yk, ym = Y[k], Y[M - k]
        e = yk + np.conj(ym)
        o = np.conj(tw[k]) * (yk - np.conj(ym))                 # w^-k
        put(k, C64(e + 1j * o))
        if k != M - k:
            # Z[M-k] = conj(e) + i * w^-(M-k) * (ym - conj(yk)) = conj(e) + i * conj(o)
            put(M - k, C64(np.conj(e) + 1j * np.conj(o)))
    a = stockham(A, tw, N, True)
    b = stockham(B, tw, N, True)
    m = np.arange(P)
    z = (a + np.conj(tw[2 * m]) * b) * F32(1.0 / N)
    out = np.empty(M, F32)
    out[0::2] = z.real
    out[1::2] = z.imag
    return out


def fold_partitions(H):
    """H: [bins][M] packed spectra of zero-padded IR partitions -> G: [bins+1][M]."""
    bins, M = H.shape
    sign = np.where(np.arange(M) & 1, -1.0, 1.0).astype(F32)
    G = np.zeros((bins + 1, M), C64)
    G[:bins] += H
    # sign[0] = +1 covers packed bin 0: DC (k=0) and Nyquist (k=M, M even) both take '+'
    G[1:] += H * sign
    return G


def mac(G, ring, s0, qa, qb, splits):
    """K2.  sum_{q in [qa,qb)} G[q] * ring[(s0 + q) % S], bin 0 = (DC*DC, Ny*Ny); `splits`
    partial rows summed afterwards in fp32 like K3's pre-pass does."""
    S = ring.shape[0]
    M = G.shape[1]
    parts = []
    nq = qb - qa
    for c in range(splits):
        a = qa + (nq * c) // splits
        b = qa + (nq * (c + 1)) // splits
        acc = np.zeros(M, C64)
        d = F32(0)
        for q in range(a, b):
            g, x = G[q], ring[(s0 + q) % S]
            acc = (acc + g * x).astype(C64)
            d = F32(d + g[0].imag * x[0].imag)
        acc[0] = C64((acc[0].real + d) + 1j * d)
        parts.append(acc)
    y = parts[0]
    for p in parts[1:]:
        y = (y + p).astype(C64)
    return y


class ModelConvolver:
    """Host scheduler + kernels for one instance, any call size, any phase."""

    def __init__(self, ir, rank, phase, splits=3, part_offset=0):
        rank = min(max(int(rank), 8), 16)
        self.rank = rank
        self.F = F = 1 << (rank - 1)
        self.tw = twiddle_table(1 << rank)
        ir = np.asarray(ir, F32)
        bins = (ir.size + F - 1) // F
        padded = np.zeros(bins * F, F32)
        padded[:ir.size] = ir
        H = np.stack([fwd_half_spectrum(padded[p * F:(p + 1) * F], rank, self.tw) for p in range(bins)])
        self.G = fold_partitions(H)
        self.q_lo = part_offset                      # partition-range sharding: global index of G[0]
        self.nq = bins + 1
        self.S = self.q_lo + self.nq                 # ring slots
        self.ring = np.zeros((self.S, F), C64)
        self.head = padded[:F].copy() if part_offset == 0 else np.zeros(F, F32)
        self.cur = np.zeros(F, F32)
        self.pend = np.zeros(F, F32)
        self.pend_valid = False
        self.off = int(F32(phase) * F32(F)) % F      # Convolver.cpp:140
        self.t = 0                                   # completed frames
        self.splits = splits

    def _slot0(self, t):
        # frame t lives in slot (-t) mod S; X_{t-q} is at (slot0(t) + q) mod S
        return (-t) % self.S

    def _push_frame(self, frame):
        self.ring[self._slot0(self.t)] = fwd_half_spectrum(frame, self.rank, self.tw)

    def _mac(self, qa_global):
        """sum over global q in [max(qa,q_lo), q_lo+nq) of G_q X_{t-q} for the current self.t."""
        qa = max(qa_global, self.q_lo)
        qb = self.q_lo + self.nq
        # local rows index q - q_lo
        Gl = self.G
        s0 = self._slot0(self.t)
        S = self.S
        parts_ring = self.ring
        # emulate with global q: ring slot (s0 + q) % S, G row q - q_lo
        M = self.F
        acc_parts = []
        nq = qb - qa
        for c in range(self.splits):
            a = qa + (nq * c) // self.splits
            b = qa + (nq * (c + 1)) // self.splits
            acc = np.zeros(M, C64)
            d = F32(0)
            for q in range(a, b):
                g, x = Gl[q - self.q_lo], parts_ring[(s0 + q) % S]
                acc = (acc + g * x).astype(C64)
                d = F32(d + g[0].imag * x[0].imag)
            acc[0] = C64((acc[0].real + d) + 1j * d)
            acc_parts.append(acc)
        y = acc_parts[0]
        for p in acc_parts[1:]:
            y = (y + p).astype(C64)
        return y

    def process(self, src):
        src = np.asarray(src, F32)
        out = np.zeros_like(src)
        F = self.F
        i = 0
        while i < src.size:
            if self.off == F:
                # deferred completion of a frame that was delivered in pieces
                self._push_frame(self.cur)
                self.t += 1
                self.off = 0
                self.cur[:] = 0
                self.pend_valid = False
            n = min(src.size - i, F - self.off)
            if self.off == 0 and n == F:
                # full aligned frame: FFT -> MAC over all q -> IFFT -> out
                self._push_frame(src[i:i + F])
                out[i:i + F] = inv_first_half(self._mac(0), self.rank, self.tw)
                self.t += 1
                self.pend_valid = False
            else:
                if not self.pend_valid:
                    # contributions of completed frames to the frame in progress (q >= 1):
                    # evaluated with X_t = 0 in slot0(t) excluded by qa = 1
                    self.pend = inv_first_half(self._mac(1), self.rank, self.tw)
                    self.pend_valid = True
                self.cur[self.off:self.off + n] = src[i:i + n]
                for m in range(self.off, self.off + n):
                    own = np.dot(self.cur[:m + 1].astype(F32), self.head[m::-1][:m + 1].astype(F32))
                    out[i + m - self.off] = F32(self.pend[m] + own)
                self.off += n
            i += n
        return out
