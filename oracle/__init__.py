"""TEST INFRASTRUCTURE ONLY -- ctypes front-end to the two CPU oracle builds.

    kind="ref"  -> oracle/_ref/libbfir_ref.so     the UNMODIFIED reference sources (compiled from
                   /root/reference/brutefir through oracle/ref_shim; see oracle/Makefile)
    kind="ref_mkl" -> oracle/_ref/libbfir_ref_mkl.so the same reference objects with MKL DFTI (PyTorch's
                   libtorch_cpu.so) behind the FFTW calls: the tuned CPU FFT for BASELINE TIMINGS only
    kind="port" -> oracle/_build/libbfir_oracle.so the restatement in oracle/bfir_oracle.cpp

Both export the same entry points (prefix ``ref_`` / ``orc_``), so every test can be run against
either.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs
may import this package; the product (foo-dsp-bfir_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libbfir_ref.so")
REF_MKL_SO = os.path.join(HERE, "_ref", "libbfir_ref_mkl.so")
PORT_SO = os.path.join(HERE, "_build", "libbfir_oracle.so")
_PATHS = {"ref": REF_SO, "ref_mkl": REF_MKL_SO, "port": PORT_SO}

# sample formats, reference brutefir/global.h:24-37
S8, S16_LE, S16_BE, S24_LE, S24_BE, S32_LE, S32_BE, FLOAT_LE, FLOAT_BE, FLOAT64_LE, FLOAT64_BE = range(1, 12)
FORMAT_BYTES = {S8: 1, S16_LE: 2, S16_BE: 2, S24_LE: 3, S24_BE: 3, S32_LE: 4, S32_BE: 4,
                FLOAT_LE: 4, FLOAT_BE: 4, FLOAT64_LE: 8, FLOAT64_BE: 8}
MIX_INPUT, MIX_INPUT_ADD, MIX_OUTPUT = 1, 2, 3


class Overflow(ctypes.Structure):  # bfoverflow_t, global.h:96-102
    _fields_ = [("n_overflows", ctypes.c_uint), ("intlargest", ctypes.c_int32),
                ("largest", ctypes.c_double), ("max", ctypes.c_double)]

    def as_tuple(self):
        return (self.n_overflows, self.intlargest, self.largest, self.max)


def build(kind="all", quiet=True):
    """Compile the oracle libraries (gcc only). `ref` needs /root/reference and is skipped without it."""
    out = subprocess.run(["make", "-C", HERE, kind], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def available(kind):
    return os.path.exists(_PATHS[kind])


def _mkl_library():
    """PyTorch's libtorch_cpu.so (exports MKL's Dfti* entry points), located without importing torch"""
    import importlib.util
    spec = importlib.util.find_spec("torch")
    if spec is None or not spec.origin:
        return None
    path = os.path.join(os.path.dirname(spec.origin), "lib", "libtorch_cpu.so")
    return path if os.path.exists(path) else None


def best_timing_kind():
    """The fastest honest CPU arm for baseline timings: the reference sources on MKL's FFT when that library loads
    and its provider reports itself, else best_kind()."""
    if available("ref_mkl") and _mkl_library() is not None:
        try:
            if lib("ref_mkl").fft_provider():
                return "ref_mkl"
        except OSError:
            pass
    return best_kind()


def best_kind():
    """'ref' (the real reference sources) when its prebuilt .so is here, else the port."""
    return "ref" if available("ref") else "port"


_libs = {}


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def lib(kind):
    if kind in _libs:
        return _libs[kind]
    path = _PATHS[kind]
    if not os.path.exists(path):
        build("port" if kind == "port" else "ref")
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if kind == "ref_mkl":
        os.environ.setdefault("BFIR_MKL_LIB", _mkl_library() or "")
    L = ctypes.CDLL(path)
    p = "orc_" if kind == "port" else "ref_"
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double

    def sig(name, res, *args):
        f = getattr(L, p + name)
        f.restype = res
        f.argtypes = list(args)
        return f

    ns = type("OracleNS", (), {})()
    ns.kind = kind
    ns.conv_new = sig("conv_new", vp, ci, ci, ci, ci)
    ns.conv_delete = sig("conv_delete", None, vp)
    ns.conv_cbufsize = sig("conv_cbufsize", ci, vp)
    ns.conv_raw2cbuf = sig("conv_raw2cbuf", ci, vp, vp, vp, vp, ci, ci, ci)
    ns.conv_time2freq = sig("conv_time2freq", None, vp, vp, vp)
    ns.conv_freq2time = sig("conv_freq2time", None, vp, vp, vp)
    ns.conv_mixnscale = sig("conv_mixnscale", None, vp, ctypes.POINTER(vp), vp, ctypes.POINTER(cd), ci, ci)
    ns.conv_convolve_inplace = sig("conv_convolve_inplace", None, vp, vp, vp)
    ns.conv_convolve = sig("conv_convolve", None, vp, vp, vp, vp)
    ns.conv_convolve_add = sig("conv_convolve_add", None, vp, vp, vp, vp)
    ns.conv_crossfade_inplace = sig("conv_crossfade_inplace", None, vp, vp, vp, vp)
    ns.conv_dirac_convolve = sig("conv_dirac_convolve", None, vp, vp, vp)
    ns.conv_dirac_convolve_inplace = sig("conv_dirac_convolve_inplace", None, vp, vp)
    ns.conv_convolve_eval = sig("conv_convolve_eval", None, vp, vp, vp, vp)
    ns.conv_td_block_length = sig("conv_td_block_length", ci, vp, ci)
    ns.conv_td_new = sig("conv_td_new", vp, vp, vp, ci)
    ns.conv_td_coeffs = sig("conv_td_coeffs", None, vp, vp, ci)
    ns.conv_td_convolve = sig("conv_td_convolve", None, vp, vp, vp)
    ns.conv_td_free = sig("conv_td_free", None, vp)
    ns.conv_cbuf2raw = sig("conv_cbuf2raw", ci, vp, vp, vp, ci, ci, ci, ci, ci, ctypes.POINTER(Overflow))
    ns.conv_coeffs2cbuf = sig("conv_coeffs2cbuf", ci, vp, vp, ci, cd, vp)
    ns.conv_runtime_coeffs2cbuf = sig("conv_runtime_coeffs2cbuf", None, vp, vp, vp)
    ns.conv_dither_table_size = sig("conv_dither_table_size", ci, vp)
    ns.conv_dither_table = sig("conv_dither_table", ctypes.POINTER(ctypes.c_int8), vp)
    ns.conv_dither_ptr = sig("conv_dither_ptr", ci, vp, ci)
    ns.conv_dither_map = sig("conv_dither_map", None, vp, vp)
    ns.raw2real = sig("raw2real", None, ci, vp, vp, ci, ci, ci, ci, ci, ci)
    ns.bfir_new = sig("bfir_new", vp, ci, ci, ci, ci, ci, ci, ci, ci)
    ns.bfir_delete = sig("bfir_delete", None, vp)
    ns.bfir_is_initialized = sig("bfir_is_initialized", ci, vp)
    ns.bfir_set_coeff = sig("bfir_set_coeff", ci, vp, ctypes.POINTER(vp), ci, ci, ci, cd)
    ns.bfir_run = sig("bfir_run", ci, vp, vp, vp)
    ns.bfir_reset = sig("bfir_reset", None, vp)
    ns.bfir_get_overflow = sig("bfir_get_overflow", None, vp, ci, ctypes.POINTER(Overflow))
    ns.bfir_dither_ptr = sig("bfir_dither_ptr", ci, vp, ci)
    ns.bfir_blockcounter = sig("bfir_blockcounter", ctypes.c_uint, vp)
    ns.preprocess_coeff = sig("preprocess_coeff", ci, vp, vp, ci, ci, ci, ci, cd, vp)
    ns.equalizer_render = sig("equalizer_render", ci, ci, ci, ci, ci, ci, ci,
                              ctypes.POINTER(cd), ctypes.POINTER(cd), ctypes.POINTER(cd), vp, ci)
    ns.fft_provider = sig("fft_provider", ctypes.c_char_p)
    if kind != "port":   # the reference's own offline drivers (brutefir/preprocessor.cpp), in-memory sound files
        wp = ctypes.c_wchar_p
        ns.vfile_register = sig("vfile_register", None, wp, ci, ci, ci, vp)
        ns.vfile_clear = sig("vfile_clear", None)
        ns.set_noise = sig("set_noise", None, vp, ctypes.c_longlong)
        ns.convolve_impulses = sig("convolve_impulses", ci, ctypes.POINTER(wp), vp, ci, ci, ci, vp, ctypes.c_longlong, ctypes.POINTER(ci))
        ns.calculate_attenuation = sig("calculate_attenuation", ci, wp, ci, ci, ctypes.POINTER(cd), ctypes.POINTER(ci),
                                       ctypes.POINTER(ci), ctypes.POINTER(ci))
    _libs[kind] = ns
    if kind in ("ref", "ref_mkl"):
        # Reference quirk (fftw_convolver.cpp:543-549): convolver_runtime_coeffs2cbuf keeps a
        # function-static scratch sized by its FIRST caller and shared by every later instance, so a
        # larger convolver overruns it. Prime it once with the largest cbuf any caller here uses.
        h = ns.conv_new(65536, 8, 1, 44100)
        src = np.zeros(65536, dtype=np.float64)
        dst = np.zeros(131072, dtype=np.float64)
        ns.conv_runtime_coeffs2cbuf(h, _ptr(src), _ptr(dst))
        ns.conv_delete(h)
    return ns


def real_dtype(realsize):
    return np.float32 if realsize == 4 else np.float64


class Convolver:
    """The reference's `fftw_convolver` (brutefir/fftw_convolver.hpp:28-166) on host numpy cbufs."""

    def __init__(self, length, realsize, kind=None, n_channels=1, sample_rate=44100):
        self.ns = lib(kind or best_kind())
        self.L, self.N, self.realsize = length, 2 * length, realsize
        self.dtype = real_dtype(realsize)
        self.h = self.ns.conv_new(length, realsize, n_channels, sample_rate)
        if not self.h:
            raise ValueError("invalid convolver parameters")

    def close(self):
        if self.h:
            self.ns.conv_delete(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cbuf(self, n=None):
        return np.zeros(self.N if n is None else n, dtype=self.dtype)

    def cbufsize(self):
        return self.ns.conv_cbufsize(self.h)

    def raw2cbuf(self, raw, cbuf, next_cbuf, fmt, index, spacing):
        assert self.ns.conv_raw2cbuf(self.h, _ptr(raw), _ptr(cbuf), _ptr(next_cbuf), fmt, index, spacing) == 0

    def time2freq(self, x, out=None):
        out = self.cbuf() if out is None else out
        self.ns.conv_time2freq(self.h, _ptr(x), _ptr(out))
        return out

    def freq2time(self, x, out=None):
        out = self.cbuf() if out is None else out
        self.ns.conv_freq2time(self.h, _ptr(x), _ptr(out))
        return out

    def mixnscale(self, bufs, scales, mixmode, out=None):
        out = self.cbuf() if out is None else out
        n = len(bufs)
        arr = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sc = (ctypes.c_double * n)(*[float(s) for s in scales])
        self.ns.conv_mixnscale(self.h, arr, _ptr(out), sc, n, mixmode)
        return out

    def convolve(self, x, c, out=None):
        out = self.cbuf() if out is None else out
        self.ns.conv_convolve(self.h, _ptr(x), _ptr(c), _ptr(out))
        return out

    def convolve_add(self, x, c, out):
        self.ns.conv_convolve_add(self.h, _ptr(x), _ptr(c), _ptr(out))
        return out

    def convolve_inplace(self, x, c):
        self.ns.conv_convolve_inplace(self.h, _ptr(x), _ptr(c))
        return x

    def crossfade_inplace(self, x, xfade, buffer):
        self.ns.conv_crossfade_inplace(self.h, _ptr(x), _ptr(xfade), _ptr(buffer))
        return x

    def dirac_convolve(self, x, out=None):
        out = self.cbuf() if out is None else out
        self.ns.conv_dirac_convolve(self.h, _ptr(x), _ptr(out))
        return out

    def dirac_convolve_inplace(self, x):
        self.ns.conv_dirac_convolve_inplace(self.h, _ptr(x))
        return x

    def convolve_eval(self, x, buffer, out=None):
        out = self.cbuf() if out is None else out
        self.ns.conv_convolve_eval(self.h, _ptr(x), _ptr(buffer), _ptr(out))
        return out

    # small one-shot convolver (fftw_convolver.cpp:698-777): returns (blocklen, handle)
    def td_block_length(self, n_coeffs):
        return self.ns.conv_td_block_length(self.h, n_coeffs)

    def td_new(self, coeffs):
        coeffs = np.ascontiguousarray(coeffs, dtype=self.dtype)
        bl = self.td_block_length(len(coeffs))
        if bl < 0:
            return -1, None
        return bl, self.ns.conv_td_new(self.h, _ptr(coeffs), len(coeffs))

    def td_coeffs(self, tdc, blocklen):
        out = np.zeros(2 * blocklen, dtype=self.dtype)
        self.ns.conv_td_coeffs(tdc, _ptr(out), out.nbytes)
        return out

    def td_convolve(self, tdc, overlap_block):
        self.ns.conv_td_convolve(self.h, tdc, _ptr(overlap_block))
        return overlap_block

    def td_free(self, tdc):
        self.ns.conv_td_free(tdc)

    def cbuf2raw(self, cbuf, out_raw, fmt, index, spacing, apply_dither, dither_channel, overflow):
        assert self.ns.conv_cbuf2raw(self.h, _ptr(cbuf), _ptr(out_raw), fmt, index, spacing,
                                     int(apply_dither), dither_channel, ctypes.byref(overflow)) == 0

    def coeffs2cbuf(self, coeffs, scale=1.0, n=None, out=None):
        out = self.cbuf() if out is None else out
        coeffs = np.ascontiguousarray(coeffs, dtype=self.dtype)
        rc = self.ns.conv_coeffs2cbuf(self.h, _ptr(coeffs), len(coeffs) if n is None else n, scale, _ptr(out))
        return out if rc == 0 else None

    def runtime_coeffs2cbuf(self, src, out=None):
        out = self.cbuf() if out is None else out
        src = np.ascontiguousarray(src, dtype=self.dtype)
        self.ns.conv_runtime_coeffs2cbuf(self.h, _ptr(src), _ptr(out))
        return out

    def preprocess_coeff(self, coeffs, blocks, scale=1.0):
        coeffs = np.ascontiguousarray(coeffs, dtype=self.dtype)
        out = np.zeros((blocks, self.N), dtype=self.dtype)
        rc = self.ns.preprocess_coeff(self.h, _ptr(coeffs), self.L, blocks, len(coeffs), self.realsize, scale, _ptr(out))
        return out if rc == 0 else None

    def dither_table(self):
        n = self.ns.conv_dither_table_size(self.h)
        p = self.ns.conv_dither_table(self.h)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def dither_map(self):
        out = np.zeros(511, dtype=self.dtype)
        self.ns.conv_dither_map(self.h, _ptr(out))
        return out

    def dither_ptr(self, ch):
        return self.ns.conv_dither_ptr(self.h, ch)


class Engine:
    """The reference's `brutefir` class (brutefir/brutefir.hpp:15-52)."""

    def __init__(self, filter_length, filter_blocks, realsize, channels, in_format, out_format,
                 sampling_rate, apply_dither, kind=None):
        kind = kind or best_kind()
        if kind == "ref" and channels > 8:  # BF_MAXCHANNELS, global.h:21
            kind = "port"
        self.ns = lib(kind)
        self.kind = kind
        self.L, self.P, self.C, self.realsize = filter_length, filter_blocks, channels, realsize
        self.in_format, self.out_format = in_format, out_format
        self.dtype = real_dtype(realsize)
        self.h = self.ns.bfir_new(filter_length, filter_blocks, realsize, channels, in_format, out_format,
                                  sampling_rate, int(apply_dither))
        if not self.h:
            raise ValueError("invalid engine parameters")

    def close(self):
        if self.h:
            self.ns.bfir_delete(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def is_initialized(self):
        return bool(self.ns.bfir_is_initialized(self.h))

    def set_coeff(self, coeffs, coeff_blocks, scale=1.0, length=None):
        arrs = [np.ascontiguousarray(c, dtype=self.dtype) for c in coeffs]
        length = len(arrs[0]) if length is None else length
        ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        return self.ns.bfir_set_coeff(self.h, ptrs, len(arrs), length, coeff_blocks, scale)

    def run(self, inbuf, outbuf=None):
        if outbuf is None:
            outbuf = np.zeros(self.L * self.C * FORMAT_BYTES[self.out_format], dtype=np.uint8)
        rc = self.ns.bfir_run(self.h, _ptr(inbuf), _ptr(outbuf))
        return rc, outbuf

    def reset(self):
        self.ns.bfir_reset(self.h)

    def overflow(self, ch):
        o = Overflow()
        self.ns.bfir_get_overflow(self.h, ch, ctypes.byref(o))
        return o

    def dither_ptr(self, ch):
        return self.ns.bfir_dither_ptr(self.h, ch)

    def blockcounter(self):
        return self.ns.bfir_blockcounter(self.h)


def equalizer_render(block_length, n_blocks, realsize, sampling_rate, freq, mag, phase, kind=None, n_channels=1):
    ns = lib(kind or best_kind())
    taps = block_length * n_blocks
    out = np.zeros(taps // 2, dtype=real_dtype(realsize))
    n = len(freq)
    arr = lambda v: (ctypes.c_double * n)(*[float(x) for x in v])
    rc = ns.equalizer_render(block_length, n_blocks, realsize, n_channels, sampling_rate, n, arr(freq), arr(mag),
                             arr(phase), _ptr(out), len(out))
    if rc < 0:
        raise ValueError("equalizer_render failed")
    return out


def ref_convolve_impulses(responses, scales, filter_length, realsize, rate=44100, kind="ref"):
    """preprocessor::convolve_impulses of the reference itself (brutefir/preprocessor.cpp:34-233) on in-memory sound
    files: responses = list of [frames, channels] arrays. Returns [g_frames, channels] in the engine's precision."""
    ns = lib(kind)
    ns.vfile_clear()
    names = []
    for i, r in enumerate(responses):
        a = np.ascontiguousarray(np.asarray(r, dtype=np.float64))
        ns.vfile_register("impulse%d.wav" % i, a.shape[1], a.shape[0], rate, _ptr(a))
        names.append("impulse%d.wav" % i)
    C = np.asarray(responses[0]).shape[1]
    frames = max(np.asarray(r).shape[0] for r in responses)
    out = np.zeros((frames + filter_length) * C, dtype=real_dtype(realsize))
    sc = np.ascontiguousarray(np.asarray(scales, dtype=np.float64))
    ch = ctypes.c_int(0)
    n = ns.convolve_impulses((ctypes.c_wchar_p * len(names))(*names), _ptr(sc), len(names), filter_length, realsize, _ptr(out),
                             out.nbytes, ctypes.byref(ch))
    if n <= 0:
        raise ValueError("reference convolve_impulses failed")
    return out[:n * ch.value].reshape(n, ch.value)


def ref_calculate_attenuation(response, filter_length, realsize, noise, rate=44100, kind="ref"):
    """preprocessor::calculate_attenuation of the reference itself (brutefir/preprocessor.cpp:250-412); `noise` is what
    its buffer::load_white_noise returns (interleaved, filter_length * blocks * channels samples). Returns dB."""
    ns = lib(kind)
    a = np.ascontiguousarray(np.asarray(response, dtype=np.float64))
    ns.vfile_register("response.wav", a.shape[1], a.shape[0], rate, _ptr(a))
    nz = np.ascontiguousarray(np.asarray(noise, dtype=np.float64).ravel())
    ns.set_noise(_ptr(nz), nz.size)
    att, c, f, r = ctypes.c_double(0), ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    if not ns.calculate_attenuation("response.wav", filter_length, realsize, ctypes.byref(att), ctypes.byref(c), ctypes.byref(f), ctypes.byref(r)):
        raise ValueError("reference calculate_attenuation failed")
    return att.value
