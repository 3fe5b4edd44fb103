"""oracle/bindings.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers:

* ``liboracle.so``              plain-C restatement (oracle/convolver_oracle.c), always buildable;
* ``_ref/libref_convolver.so``  the reference's own ``Convolver.cpp`` compiled verbatim over the
                                restated ``lsp::dsp::`` kernels (only buildable where
                                ``/root/reference`` exists; the prebuilt file travels to the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``lsp-dsp-units_b200``) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_FP = ctypes.POINTER(ctypes.c_float)
_SZ = ctypes.c_size_t


def build(quiet=True):
    """Compile liboracle.so and (where the reference tree is present) _ref/libref_convolver.so."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _ptr(a):
    return a.ctypes.data_as(_FP)


class _State(ctypes.Structure):
    _fields_ = [(n, _SZ) for n in ("data_buffer_size", "direct_size", "frame_size", "frame_off",
                                   "conv_size", "levels", "blocks", "blocks_done", "rank",
                                   "blk_init")] + [("blk_coef", ctypes.c_float)]


def _load(path, prefix):
    lib = ctypes.CDLL(path)
    f = lambda n: getattr(lib, prefix + n)
    f("create").restype = ctypes.c_void_p
    f("create").argtypes = []
    f("free").argtypes = [ctypes.c_void_p]
    f("destroy").argtypes = [ctypes.c_void_p]
    f("init").argtypes = [ctypes.c_void_p, _FP, _SZ, _SZ, ctypes.c_float]
    f("init").restype = ctypes.c_int
    f("process").argtypes = [ctypes.c_void_p, _FP, _FP, _SZ]
    f("data_size").argtypes = [ctypes.c_void_p]
    f("data_size").restype = _SZ
    f("rank").argtypes = [ctypes.c_void_p]
    f("rank").restype = _SZ
    f("bench").argtypes = [_SZ] * 7 + [ctypes.POINTER(ctypes.c_double)]
    f("bench").restype = ctypes.c_double
    return lib


class CpuConvolver:
    """One CPU convolver with the reference's init/process/destroy surface.

    ``impl="oracle"`` -> the plain-C restatement; ``impl="reference"`` -> the verbatim build.
    """

    _libs = {}

    @classmethod
    def lib(cls, impl):
        if impl not in cls._libs:
            if impl == "oracle":
                path, prefix = os.path.join(_HERE, "liboracle.so"), "orc_"
            elif impl == "reference":
                path, prefix = os.path.join(_HERE, "_ref", "libref_convolver.so"), "refconv_"
            else:
                raise ValueError(impl)
            if not os.path.exists(path):
                if impl == "oracle":
                    build()
                else:
                    raise FileNotFoundError(path)
            cls._libs[impl] = (_load(path, prefix), prefix)
        return cls._libs[impl]

    @classmethod
    def available(cls, impl):
        try:
            cls.lib(impl)
            return True
        except (OSError, RuntimeError):
            return False

    def __init__(self, impl="oracle"):
        self._lib, self._p = self.lib(impl)
        self.impl = impl
        self._h = self._f("create")()

    def _f(self, name):
        return getattr(self._lib, self._p + name)

    def init(self, ir, rank, phase=0.0):
        ir = np.ascontiguousarray(ir, dtype=np.float32)
        return bool(self._f("init")(self._h, _ptr(ir), ir.size, rank, phase))

    def destroy(self):
        self._f("destroy")(self._h)

    def process(self, src, out=None):
        """Convolver::process(dst, src, count) on a float32 array; returns dst."""
        src = np.ascontiguousarray(src, dtype=np.float32)
        if out is None:
            out = np.empty_like(src)
        self._f("process")(self._h, _ptr(out), _ptr(src), src.size)
        return out

    def run(self, src, step):
        """Feed ``src`` in calls of ``step`` samples (reference utest helper convolver.cpp:43-53)."""
        src = np.ascontiguousarray(src, dtype=np.float32)
        out = np.zeros_like(src)
        for i in range(0, src.size, step):
            n = min(step, src.size - i)
            self._f("process")(self._h, _ptr(out[i:]), _ptr(src[i:]), n)
        return out

    def data_size(self):
        return int(self._f("data_size")(self._h))

    def rank(self):
        return int(self._f("rank")(self._h))

    def state(self):
        if self.impl != "oracle":
            raise NotImplementedError
        st = _State()
        self._lib.orc_get_state.argtypes = [ctypes.c_void_p, ctypes.POINTER(_State)]
        self._lib.orc_get_state(self._h, ctypes.byref(st))
        return {n: getattr(st, n) for n, _ in _State._fields_}

    def dump_names(self):
        if self.impl != "reference":
            raise NotImplementedError
        buf = ctypes.create_string_buffer(1024)
        self._lib.refconv_dump_names.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _SZ]
        self._lib.refconv_dump_names.restype = _SZ
        n = self._lib.refconv_dump_names(self._h, buf, 1024)
        return int(n), buf.value.decode().split(",")

    def __del__(self):
        try:
            if self._h:
                self._f("free")(self._h)
                self._h = None
        except Exception:
            pass


def cpu_bench(impl, instances, taps, rank, block, warm_blocks, blocks, threads):
    """Time ``blocks`` process() calls of ``block`` samples on ``instances`` CPU convolvers.

    Returns (output samples per second over all instances, elapsed seconds)."""
    lib, prefix = CpuConvolver.lib(impl)
    chk = ctypes.c_double(0.0)
    sec = getattr(lib, prefix + "bench")(instances, taps, rank, block, warm_blocks, blocks,
                                         threads, ctypes.byref(chk))
    if sec <= 0:
        raise RuntimeError("cpu bench failed")
    return instances * block * blocks / sec, sec


def direct_convolve(src, ir, count=None):
    """Float64 linear convolution -- the identity the reference's utest pins against
    (naive direct form, src/test/utest/util/convolver.cpp:32-40).  Small problems use the direct
    form itself; large ones the float64 FFT form (agrees with it to ~1e-13 of peak)."""
    src = np.asarray(src, dtype=np.float64)
    ir = np.asarray(ir, dtype=np.float64)
    if src.size * ir.size <= (1 << 24):
        y = np.convolve(src, ir)
    else:
        from scipy.signal import fftconvolve
        y = fftconvolve(src, ir)
    return y if count is None else y[:count]


class CpuEqualizer:
    """Data path of ``dspu::Equalizer`` in EQM_FIR / EQM_FFT mode (oracle/equalizer_oracle.c):
    ``set_kernel(ir, smooth)`` / ``clear()`` / ``process(x)`` with ``fir_size`` samples of latency."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
            lib.orc_eq_create.restype = ctypes.c_void_p
            lib.orc_eq_create.argtypes = [_SZ]
            lib.orc_eq_free.argtypes = [ctypes.c_void_p]
            lib.orc_eq_set_kernel.argtypes = [ctypes.c_void_p, _FP, ctypes.c_int]
            lib.orc_eq_clear.argtypes = [ctypes.c_void_p]
            lib.orc_eq_process.argtypes = [ctypes.c_void_p, _FP, _FP, _SZ]
            lib.orc_eq_fir_size.argtypes = [ctypes.c_void_p]
            lib.orc_eq_fir_size.restype = _SZ
            cls._lib = lib
        return cls._lib

    def __init__(self, fir_rank):
        self._h = self.lib().orc_eq_create(fir_rank)
        if not self._h:
            raise MemoryError("orc_eq_create")
        self.fir_size = 1 << fir_rank

    def set_kernel(self, ir, smooth=False):
        ir = np.ascontiguousarray(ir, dtype=np.float32)
        assert ir.size == self.fir_size
        self.lib().orc_eq_set_kernel(self._h, _ptr(ir), int(bool(smooth)))

    def clear(self):
        self.lib().orc_eq_clear(self._h)

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty_like(x)
        self.lib().orc_eq_process(self._h, _ptr(y), _ptr(x), x.size)
        return y

    def run(self, x, step):
        return np.concatenate([self.process(x[i:i + step]) for i in range(0, len(x), step)])

    def __del__(self):
        try:
            if self._h:
                self.lib().orc_eq_free(self._h)
                self._h = None
        except Exception:
            pass


# ---- SyncChirpProcessor::do_linear_convolutions (oracle/chirp_oracle.c) ---------------------------

class _ChirpPlan(ctypes.Structure):
    _fields_ = [(n, _SZ) for n in ("partition_size", "conv_rank", "image", "allocation_size")]


def _chirp_lib():
    lib = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
    SZP = ctypes.POINTER(_SZ)
    lib.orc_chirp_plan.argtypes = [ctypes.POINTER(_ChirpPlan), SZP, SZP, SZP, SZP, SZP, SZP, _SZ, _SZ, _SZ]
    lib.orc_chirp_plan.restype = None
    lib.orc_chirp_linear_convolutions.argtypes = [_FP, ctypes.POINTER(_FP), SZP, _SZ, _FP, _SZ, _SZ, ctypes.c_float]
    lib.orc_chirp_linear_convolutions.restype = ctypes.c_int
    return lib


def chirp_plan(in_len, inverse_len, part_size_limit):
    """calculateConvolutionPartitionSize + calculateConvolutionParameters
    (SyncChirpProcessor.cpp:1224-1250, 1299-1331) -> dict of scalars and per-channel lists."""
    n = len(in_len)
    arr = lambda: (_SZ * n)()
    plan, parts, padded, prep, clen, align = _ChirpPlan(), arr(), arr(), arr(), arr(), arr()
    _chirp_lib().orc_chirp_plan(ctypes.byref(plan), parts, padded, prep, clen, align, (_SZ * n)(*in_len), n,
                                inverse_len, part_size_limit)
    out = {k: int(getattr(plan, k)) for k, _ in _ChirpPlan._fields_}
    out.update(partitions=list(parts), padded=list(padded), prepends=list(prep), conv_lengths=list(clen),
               align_offsets=list(align))
    return out


def chirp_linear_convolutions(inputs, inverse, part_size_limit, scale):
    """SyncChirpProcessor::do_linear_convolutions (:1374-1508) on plain arrays: returns the
    [nchannels][allocation_size] float32 result."""
    inputs = [np.ascontiguousarray(x, dtype=np.float32) for x in inputs]
    inverse = np.ascontiguousarray(inverse, dtype=np.float32)
    n = len(inputs)
    lens = [x.size for x in inputs]
    plan = chirp_plan(lens, inverse.size, part_size_limit)
    res = np.empty((n, plan["allocation_size"]), dtype=np.float32)
    rc = _chirp_lib().orc_chirp_linear_convolutions(_ptr(res), (_FP * n)(*[_ptr(x) for x in inputs]),
                                                    (_SZ * n)(*lens), n, _ptr(inverse), inverse.size,
                                                    part_size_limit, scale)
    if rc != 0:
        raise RuntimeError("orc_chirp_linear_convolutions failed")
    return res


# ---- lsp::dspu::SpectralProcessor, the reference class itself (oracle/ref_wrap_spectral.cpp) -------

class CpuSpectralProcessor:
    """The reference's own ``dspu::SpectralProcessor`` (SpectralProcessor.cpp compiled verbatim into
    ``oracle/_ref``) with one of two host callbacks: ``bind_complex(H)`` multiplies the packed complex
    spectrum (``2**rank`` bins) by ``H``, ``bind_gain(g)`` scales bin ``k`` by the real ``g[k]``."""

    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(_HERE, "_ref", "libref_convolver.so"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_convolver.so"))
            lib.refsp_create.restype = ctypes.c_void_p
            lib.refsp_create.argtypes = [_SZ]
            lib.refsp_free.argtypes = [ctypes.c_void_p]
            lib.refsp_set_rank.argtypes = [ctypes.c_void_p, _SZ]
            lib.refsp_set_phase.argtypes = [ctypes.c_void_p, ctypes.c_float]
            for name in ("refsp_rank", "refsp_latency", "refsp_remaining"):
                getattr(lib, name).restype = _SZ
                getattr(lib, name).argtypes = [ctypes.c_void_p]
            lib.refsp_reset.argtypes = [ctypes.c_void_p]
            lib.refsp_update_settings.argtypes = [ctypes.c_void_p]
            lib.refsp_bind.argtypes = [ctypes.c_void_p, ctypes.c_int, _FP, _SZ]
            lib.refsp_process.argtypes = [ctypes.c_void_p, _FP, _FP, _SZ]
            cls._lib = lib
        return cls._lib

    def __init__(self, max_rank):
        self._h = self.lib().refsp_create(max_rank)

    def set_rank(self, rank):
        self.lib().refsp_set_rank(self._h, rank)

    def set_phase(self, phase):
        self.lib().refsp_set_phase(self._h, phase)

    def rank(self):
        return int(self.lib().refsp_rank(self._h))

    def latency(self):
        return int(self.lib().refsp_latency(self._h))

    def remaining(self):
        return int(self.lib().refsp_remaining(self._h))

    def reset(self):
        self.lib().refsp_reset(self._h)

    def update_settings(self):
        self.lib().refsp_update_settings(self._h)

    def unbind(self):
        self.lib().refsp_bind(self._h, 0, None, 0)

    def bind_complex(self, H):
        H = np.ascontiguousarray(H, dtype=np.complex64).view(np.float32)
        self.lib().refsp_bind(self._h, 1, _ptr(H), H.size)

    def bind_gain(self, g):
        g = np.ascontiguousarray(g, dtype=np.float32)
        self.lib().refsp_bind(self._h, 2, _ptr(g), g.size)

    def process(self, src):
        src = np.ascontiguousarray(src, dtype=np.float32)
        out = np.empty_like(src)
        self.lib().refsp_process(self._h, _ptr(out), _ptr(src), src.size)
        return out

    def run(self, src, step):
        src = np.ascontiguousarray(src, dtype=np.float32)
        out = np.zeros_like(src)
        for i in range(0, src.size, step):
            out[i:i + step] = self.process(src[i:i + step])
        return out

    def __del__(self):
        try:
            if self._h:
                self.lib().refsp_free(self._h)
                self._h = None
        except Exception:
            pass


# ---- lsp::dspu::SpectralSplitter, the reference class itself (oracle/ref_wrap_splitter.cpp) --------

class CpuSpectralSplitter:
    """The reference's own ``dspu::SpectralSplitter`` (SpectralSplitter.cpp compiled verbatim into
    ``oracle/_ref``).  Handler ``h`` gets one of: ``bind_complex(h, H)`` (spectrum times a packed
    complex table of ``2**rank`` bins), ``bind_gain(h, g)`` (real gain per bin: what ``FFTCrossover``
    does per band), ``bind_sink(h)`` (no spectral function: the input frames themselves).  Every
    bound handler has a sink; ``process`` returns ``[handlers][count]`` (rows of unbound handlers
    stay zero)."""

    _lib = None

    @classmethod
    def available(cls):
        path = os.path.join(_HERE, "_ref", "libref_convolver.so")
        if not os.path.exists(path):
            return False
        try:
            return hasattr(ctypes.CDLL(path), "refss_create")
        except OSError:
            return False

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_convolver.so"))
            lib.refss_create.restype = ctypes.c_void_p
            lib.refss_create.argtypes = [_SZ, _SZ]
            lib.refss_free.argtypes = [ctypes.c_void_p]
            lib.refss_set_rank.argtypes = [ctypes.c_void_p, _SZ]
            lib.refss_set_chunk_rank.argtypes = [ctypes.c_void_p, ctypes.c_long]
            lib.refss_set_phase.argtypes = [ctypes.c_void_p, ctypes.c_float]
            for name in ("refss_rank", "refss_latency", "refss_bindings"):
                getattr(lib, name).restype = _SZ
                getattr(lib, name).argtypes = [ctypes.c_void_p]
            lib.refss_chunk_rank.restype = ctypes.c_long
            lib.refss_chunk_rank.argtypes = [ctypes.c_void_p]
            lib.refss_clear.argtypes = [ctypes.c_void_p]
            lib.refss_update_settings.argtypes = [ctypes.c_void_p]
            lib.refss_bind.restype = ctypes.c_int
            lib.refss_bind.argtypes = [ctypes.c_void_p, _SZ, ctypes.c_int, _FP, _SZ]
            lib.refss_process.argtypes = [ctypes.c_void_p, _FP, _SZ, _FP, _SZ]
            cls._lib = lib
        return cls._lib

    def __init__(self, max_rank, handlers):
        self.handlers = handlers
        self._h = self.lib().refss_create(max_rank, handlers)
        if not self._h:
            raise ValueError("SpectralSplitter::init failed")

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib().refss_free(self._h)
            self._h = None

    def set_rank(self, rank):
        self.lib().refss_set_rank(self._h, rank)

    def set_chunk_rank(self, rank):
        self.lib().refss_set_chunk_rank(self._h, rank)

    def set_phase(self, phase):
        self.lib().refss_set_phase(self._h, phase)

    def rank(self):
        return int(self.lib().refss_rank(self._h))

    def chunk_rank(self):
        return int(self.lib().refss_chunk_rank(self._h))

    def latency(self):
        return int(self.lib().refss_latency(self._h))

    def bindings(self):
        return int(self.lib().refss_bindings(self._h))

    def clear(self):
        self.lib().refss_clear(self._h)

    def update_settings(self):
        self.lib().refss_update_settings(self._h)

    def bind_complex(self, handler, H):
        H = np.ascontiguousarray(H, dtype=np.complex64).view(np.float32)
        return self.lib().refss_bind(self._h, handler, 1, _ptr(H), H.size)

    def bind_gain(self, handler, gain):
        gain = np.ascontiguousarray(gain, dtype=np.float32)
        return self.lib().refss_bind(self._h, handler, 2, _ptr(gain), gain.size)

    def bind_sink(self, handler):
        return self.lib().refss_bind(self._h, handler, 3, None, 0)

    def unbind(self, handler):
        return self.lib().refss_bind(self._h, handler, 0, None, 0)

    def process(self, src):
        src = np.ascontiguousarray(src, dtype=np.float32)
        out = np.zeros((self.handlers, src.size), dtype=np.float32)
        if src.size:
            self.lib().refss_process(self._h, _ptr(out), src.size, _ptr(src), src.size)
        return out

    def run(self, src, step):
        src = np.ascontiguousarray(src, dtype=np.float32)
        return np.concatenate([self.process(src[i:i + step]) for i in range(0, src.size, step)], axis=1)


def crossover_band_curve(rank, sample_rate, hpf=None, lpf=None, gain=1.0, flatten=1.0):
    """The real gain curve (``2**rank`` bins) of one ``FFTCrossover`` band: the reference's own
    ``crossover::hipass_fft_set / lopass_fft_apply / lopass_fft_set`` (misc/fft_crossover.cpp compiled
    verbatim) combined as ``FFTCrossover::update_band`` does (FFTCrossover.cpp:458-480).
    ``hpf`` / ``lpf``: ``(frequency, slope in dB/oct, negative)`` or None."""
    lib = CpuSpectralSplitter.lib()
    lib.refx_band_curve.argtypes = [_FP, _SZ, ctypes.c_float, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                    ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float]
    lib.refx_band_curve.restype = None
    curve = np.zeros(1 << rank, dtype=np.float32)
    h = hpf if hpf is not None else (0.0, 0.0)
    l = lpf if lpf is not None else (0.0, 0.0)
    lib.refx_band_curve(_ptr(curve), rank, float(sample_rate), int(hpf is not None), h[0], h[1],
                        int(lpf is not None), l[0], l[1], gain, flatten)
    return curve
