"""Float64 numpy model of the data path of dspu::Equalizer in EQM_FIR / EQM_FFT mode
(reference src/main/filters/Equalizer.cpp:474-518), written with np.convolve instead of the
fastconv primitives: an independent cross-check of oracle/equalizer_oracle.c and of the GPU path.

Also the synthetic linear-phase kernels the equalizer tests use: windowed band filters of
2**fir_rank taps whose peak sits at fir_size / 2 (what Equalizer::reconfigure builds, :325-334)."""
import numpy as np


class ModelEqualizer:
    def __init__(self, fir_rank):
        self.F = 1 << fir_rank
        self.inb = np.zeros(self.F)
        self.outb = np.zeros(2 * self.F)
        self.conv = np.zeros(self.F)
        self.new = np.zeros(self.F)
        self.xfade = False
        self.nbuf = 0

    def set_kernel(self, ir, smooth=False):             # Equalizer.cpp:336-345
        if smooth:
            self.new = np.asarray(ir, dtype=np.float64).copy()
            self.xfade = True
        else:
            self.conv = np.asarray(ir, dtype=np.float64).copy()

    def clear(self):                                    # Equalizer.cpp:273-278
        self.inb[:] = 0
        self.outb[:] = 0
        self.nbuf = 0

    def process(self, x):
        F = self.F
        x = np.asarray(x, dtype=np.float64)
        y = np.empty_like(x)
        pos = 0
        while pos < len(x):
            if self.nbuf >= F:
                self.outb[:F] = self.outb[F:]
                self.outb[F:] = 0
                self.outb[:2 * F - 1] += np.convolve(self.inb, self.conv)
                if self.xfade:
                    half = F // 2
                    fft = np.zeros(2 * F)
                    self.conv = self.new.copy()
                    fft[:2 * F - 1] = np.convolve(self.inb, self.conv)
                    ramp = np.arange(F) / F
                    self.outb[half:half + F] *= 1.0 - ramp
                    self.outb[half:half + F] += fft[half:half + F] * ramp
                    self.outb[F + half:] = fft[F + half:]
                    self.xfade = False
                self.nbuf = 0
            n = min(len(x) - pos, F - self.nbuf)
            self.inb[self.nbuf:self.nbuf + n] = x[pos:pos + n]
            y[pos:pos + n] = self.outb[self.nbuf:self.nbuf + n]
            self.nbuf += n
            pos += n
        return y

    def run(self, x, step):
        return np.concatenate([self.process(x[i:i + step]) for i in range(0, len(x), step)])


def band_kernel(fir_rank, lo, hi, gain=1.0):
    """Linear-phase band filter: magnitude `gain` on normalised frequencies [lo, hi) (1 = Nyquist),
    zero phase, made causal by a half-length rotation and windowed (Blackman-Nuttall coefficients
    of reference src/main/misc/windows.cpp:205-208)."""
    F = 1 << fir_rank
    f = np.abs(np.fft.fftfreq(F)) * 2.0
    mag = np.where((f >= lo) & (f < hi), gain, 0.0)
    h = np.real(np.fft.ifft(mag))
    h = np.roll(h, F // 2)
    i = np.arange(F)
    w = (0.3635819 - 0.4891775 * np.cos(2 * np.pi * i / (F - 1)) + 0.1365995 * np.cos(4 * np.pi * i / (F - 1))
         - 0.0106411 * np.cos(6 * np.pi * i / (F - 1)))
    return (h * w).astype(np.float32)
