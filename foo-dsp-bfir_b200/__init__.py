"""foo-dsp-bfir_b200 -- B200 (sm_100a) partitioned-convolution engine behind the BruteFIR surface.

This module is the thin Python host mirror of the two reference classes the C ABI replaces
(``include/bfir_b200.h``):

* :class:`Brutefir`       <-> ``class brutefir``        (reference brutefir/brutefir.hpp:15-52)
* :class:`FftwConvolver`  <-> ``class fftw_convolver``  (reference brutefir/fftw_convolver.hpp:28-166)

Method names, argument order and return codes follow the reference.  Everything executes in
``libbfir_b200.so`` (hand-written CUDA); there is no CPU path -- loading fails loudly when the library
has not been built (``python __graft_entry__.py`` or ``make -C foo-dsp-bfir_b200``), and creating an
engine fails loudly when no CUDA device is present.

The directory name contains hyphens, so import it with
``importlib.import_module("foo-dsp-bfir_b200")`` (tests/conftest.py and bench.py do).
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BFIR_LIB") or os.path.join(HERE, "libbfir_b200.so")   # BFIR_LIB: A/B builds of the library (tools)

OK, ERR_NONFINITE, ERR_COEFF, ERR_NOT_READY, ERR_INVALID, ERR_CUDA = 0, -1, -2, -3, -4, -5

# sample formats (reference brutefir/global.h:24-37)
S8, S16_LE, S16_BE, S24_LE, S24_BE, S32_LE, S32_BE, FLOAT_LE, FLOAT_BE, FLOAT64_LE, FLOAT64_BE = range(1, 12)
FORMAT_BYTES = {S8: 1, S16_LE: 2, S16_BE: 2, S24_LE: 3, S24_BE: 3, S32_LE: 4, S32_BE: 4,
                FLOAT_LE: 4, FLOAT_BE: 4, FLOAT64_LE: 8, FLOAT64_BE: 8}
MIXMODE_INPUT, MIXMODE_INPUT_ADD, MIXMODE_OUTPUT = 1, 2, 3


class BfirError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bfir_b200 error %d: %s" % (code, msg))
        self.code = code


class Overflow(ctypes.Structure):  # bfir_overflow_t == struct bfoverflow_t (global.h:96-102)
    _fields_ = [("n_overflows", ctypes.c_uint), ("intlargest", ctypes.c_int32),
                ("largest", ctypes.c_double), ("max", ctypes.c_double)]

    def as_tuple(self):
        return (self.n_overflows, self.intlargest, self.largest, self.max)


class Config(ctypes.Structure):  # bfir_config_t
    _fields_ = [(n, ctypes.c_int) for n in (
        "filter_length", "filter_blocks", "realsize", "channels", "in_format", "out_format",
        "sampling_rate", "apply_dither", "n_streams", "device", "part_begin", "part_count", "n_groups",
        "xbar_inputs", "xbar_outputs")]


# every symbol include/bfir_b200.h declares: (name, restype, argtypes)
_vp, _ci, _cd, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
_pp = ctypes.POINTER(ctypes.c_void_p)
PRINT_CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p)
API = [
    ("bfir_create", _ci, [_pp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ci]),
    ("bfir_create_ex", _ci, [_pp, ctypes.POINTER(Config)]),
    ("bfir_destroy", None, [_vp]),
    ("bfir_is_initialized", _ci, [_vp]),
    ("bfir_set_coeff", _ci, [_vp, _pp, _ci, _ci, _ci, _cd]),
    ("bfir_set_coeff_crossfade", _ci, [_vp, _pp, _ci, _ci, _ci, _cd]),
    ("bfir_set_coeff_device", _ci, [_vp, _vp, ctypes.c_longlong, _ci, _ci, _ci, _cd, _ci]),
    ("bfir_set_crossbar", _ci, [_vp, ctypes.POINTER(_cd), ctypes.POINTER(_cd)]),
    ("bfir_run", _ci, [_vp, _vp, _vp]),
    ("bfir_host_alloc", _vp, [ctypes.c_size_t]),
    ("bfir_host_free", None, [_vp]),
    ("bfir_run_device", _ci, [_vp, _vp, _vp]),
    ("bfir_run_async", ctypes.c_longlong, [_vp, _vp, _vp]),
    ("bfir_run_device_pipelined", _ci, [_vp, _vp, _vp]),
    ("bfir_join", _ci, [_vp]),
    ("bfir_run_device_pair", _ci, [_vp, _vp, _vp, _vp, _vp, _ci]),
    ("bfir_run_async_pair", ctypes.c_longlong, [_vp, _vp, _vp, _vp, _vp]),
    ("bfir_run_device_quad", _ci, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    ("bfir_run_device_quad_staged", _ci, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    ("bfir_run_partial_quad_device", _ci, [_vp, ctypes.POINTER(_vp)]),
    ("bfir_run_finish_quad_device", _ci, [_vp, ctypes.POINTER(_vp)]),
    ("bfir_run_shard_quad_staged", _ci, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    ("bfir_run_shard_oct_staged", _ci, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    ("bfir_run_async_quad", ctypes.c_longlong, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    ("bfir_run_device_oct", _ci, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _ci]),
    ("bfir_get_mac_profile", _ci, [_vp, _ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong), _ci]),
    ("bfir_wait", _ci, [_vp, ctypes.c_longlong]),
    ("bfir_sync", _ci, [_vp]),
    ("bfir_reset", _ci, [_vp]),
    ("bfir_get_overflow", _ci, [_vp, _ci, ctypes.POINTER(Overflow)]),
    ("bfir_check_overflows", _ci, [_vp]),
    ("bfir_get_dither_ptr", _ci, [_vp, _ci, ctypes.POINTER(_ci)]),
    ("bfir_get_blockcounter", _ci, [_vp, ctypes.POINTER(ctypes.c_uint)]),
    ("bfir_run_partial_device", _ci, [_vp, _vp]),
    ("bfir_run_finish_device", _ci, [_vp, _vp]),
    ("bfir_acc_device_ptr", _vp, [_vp, ctypes.POINTER(_sz)]),
    ("bfir_peer_setup", _ci, [_vp, _ci, _ci]),
    ("bfir_peer_export", _ci, [_vp, _vp]),
    ("bfir_peer_import", _ci, [_vp, _ci, _vp]),
    ("bfir_peer_set_ptr", _ci, [_vp, _ci, _vp]),
    ("bfir_peer_recv_ptr", _vp, [_vp]),
    ("bfir_peer_own_channels", _ci, [_vp, ctypes.POINTER(_ci), ctypes.POINTER(_ci)]),
    ("bfir_set_groups", _ci, [_vp, _ci]),
    ("bfir_get_groups", _ci, [_vp]),
    ("bfir_set_coeff_map", _ci, [_vp, ctypes.POINTER(ctypes.c_int), _ci]),
    ("bfir_get_mac_split", _ci, [_vp]),
    ("bfir_get_quad_split", _ci, [_vp]),
    ("bfir_set_stream", _ci, [_vp, _vp]),
    ("bfir_set_profiling", _ci, [_vp, _ci]),
    ("bfir_get_profile", _ci, [_vp, ctypes.POINTER(_cd), ctypes.POINTER(ctypes.c_ulonglong), _ci]),
    ("bfir_set_print_callback", None, [PRINT_CB]),
    ("bfir_last_error", ctypes.c_char_p, []),
    ("bfir_kernel_launch_count", ctypes.c_ulonglong, []),
    ("bfir_conv_create", _ci, [_pp, _ci, _ci, _ci, _ci]),
    ("bfir_conv_destroy", None, [_vp]),
    ("bfir_conv_cbufsize", _ci, [_vp]),
    ("bfir_conv_alloc", _vp, [_vp, _sz]),
    ("bfir_conv_free", None, [_vp, _vp]),
    ("bfir_conv_upload", _ci, [_vp, _vp, _vp, _sz]),
    ("bfir_conv_download", _ci, [_vp, _vp, _vp, _sz]),
    ("bfir_conv_sync", _ci, [_vp]),
    ("bfir_conv_raw2cbuf", _ci, [_vp, _vp, _vp, _vp, _ci, _ci, _ci]),
    ("bfir_conv_time2freq", _ci, [_vp, _vp, _vp]),
    ("bfir_conv_mixnscale", _ci, [_vp, _pp, _vp, ctypes.POINTER(_cd), _ci, _ci]),
    ("bfir_conv_convolve_inplace", _ci, [_vp, _vp, _vp]),
    ("bfir_conv_convolve", _ci, [_vp, _vp, _vp, _vp]),
    ("bfir_conv_convolve_add", _ci, [_vp, _vp, _vp, _vp]),
    ("bfir_conv_crossfade_inplace", _ci, [_vp, _vp, _vp, _vp]),
    ("bfir_conv_dirac_convolve", _ci, [_vp, _vp, _vp]),
    ("bfir_conv_dirac_convolve_inplace", _ci, [_vp, _vp]),
    ("bfir_conv_freq2time", _ci, [_vp, _vp, _vp]),
    ("bfir_conv_convolve_eval", _ci, [_vp, _vp, _vp, _vp]),
    ("bfir_conv_td_block_length", _ci, [_ci]),
    ("bfir_conv_td_new", _ci, [_vp, ctypes.POINTER(_vp), _vp, _ci]),
    ("bfir_conv_td_blocklen", _ci, [_vp]),
    ("bfir_conv_td_coeffs", _vp, [_vp]),
    ("bfir_conv_td_convolve", _ci, [_vp, _vp, _vp]),
    ("bfir_conv_td_free", None, [_vp]),
    ("bfir_conv_cbuf2raw", _ci, [_vp, _vp, _vp, _ci, _ci, _ci, _ci, _ci, ctypes.POINTER(Overflow)]),
    ("bfir_conv_coeffs2cbuf", _ci, [_vp, _vp, _ci, _cd, _vp]),
    ("bfir_conv_runtime_coeffs2cbuf", _ci, [_vp, _vp, _vp]),
    ("bfir_conv_dither_table_size", _ci, [_vp]),
    ("bfir_conv_dither_table", _ci, [_vp, _vp, _ci]),
    ("bfir_conv_dither_map", _ci, [_vp, _vp]),
    ("bfir_conv_dither_ptr", _ci, [_vp, _ci]),
    ("bfir_eq_create", _ci, [_pp, _ci, _ci, _ci, _ci]),
    ("bfir_eq_destroy", None, [_vp]),
    ("bfir_eq_taps", _ci, [_vp]),
    ("bfir_eq_render", _ci, [_vp, _ci, ctypes.POINTER(_cd), ctypes.POINTER(_cd), ctypes.POINTER(_cd), _vp]),
    ("bfir_eq_render_device", _vp, [_vp, _ci, ctypes.POINTER(_cd), ctypes.POINTER(_cd), ctypes.POINTER(_cd)]),
]

_lib = None


def load_library(path=None):
    """dlopen libbfir_b200.so and bind every exported symbol. Raises if the library is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s not built: run `python __graft_entry__.py` (or make -C foo-dsp-bfir_b200). "
            "There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, res, args in API:
        f = getattr(lib, name)  # AttributeError here = header and library out of sync
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load_library().bfir_last_error().decode("utf-8", "replace")


def kernel_launch_count():
    return int(load_library().bfir_kernel_launch_count())


def _check(rc):
    if rc < 0:
        raise BfirError(rc, last_error())
    return rc


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    if hasattr(a, "data_ptr"):  # torch tensor
        return ctypes.c_void_p(a.data_ptr())
    if hasattr(a, "ptr"):
        return ctypes.c_void_p(a.ptr)
    return a


def _buf(a, nbytes, what):
    """_ptr() for a block buffer: numpy arrays and torch tensors must be contiguous and hold at least `nbytes`
    (the library copies / transforms exactly that many bytes from the bare pointer); raw addresses pass through."""
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("%s: numpy buffer is not C-contiguous" % what)
        if a.nbytes < nbytes:
            raise ValueError("%s: buffer holds %d bytes, the block needs %d" % (what, a.nbytes, nbytes))
    elif hasattr(a, "data_ptr") and hasattr(a, "is_contiguous"):
        if not a.is_contiguous():
            raise ValueError("%s: tensor is not contiguous" % what)
        if a.numel() * a.element_size() < nbytes:
            raise ValueError("%s: tensor holds %d bytes, the block needs %d" % (what, a.numel() * a.element_size(), nbytes))
    return _ptr(a)


def real_dtype(realsize):
    return np.float32 if realsize == 4 else np.float64


class _CudaArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (int(n),), "typestr": typestr, "version": 2}


def as_torch(ptr, n, typestr):
    """View `n` elements of device memory at `ptr` as a torch tensor (plumbing for NCCL reduces)."""
    import torch
    return torch.as_tensor(_CudaArray(ptr, n, typestr), device="cuda")


class Brutefir:
    """``class brutefir`` (reference brutefir/brutefir.hpp:15-52) on the GPU.

    ``run`` takes/returns HOST buffers of ``filter_length * channels`` interleaved samples exactly like
    the reference; ``run_device`` is the asynchronous device-buffer variant used for throughput.
    Extra keyword arguments (``n_streams``, ``device``, ``part_begin``, ``part_count``) map to
    ``bfir_config_t`` and have no reference counterpart.
    """

    def __init__(self, filter_length, filter_blocks, realsize, channels, in_format, out_format,
                 sampling_rate, apply_dither, n_streams=1, device=-1, part_begin=0, part_count=0, n_groups=0,
                 xbar_inputs=0, xbar_outputs=0):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        cfg = Config(filter_length, filter_blocks, realsize, channels, in_format, out_format,
                     sampling_rate, int(bool(apply_dither)), n_streams, device, part_begin, part_count, n_groups,
                     xbar_inputs, xbar_outputs)
        rc = self.lib.bfir_create_ex(ctypes.byref(self.h), ctypes.byref(cfg))
        if rc != OK:
            self.h = ctypes.c_void_p()
            # the reference constructor cannot fail loudly (is_initialized() stays false); we do
            raise BfirError(rc, last_error())
        self.filter_length, self.filter_blocks, self.realsize = filter_length, filter_blocks, realsize
        self.channels, self.n_streams = channels, n_streams
        self.in_format, self.out_format = in_format, out_format
        self.dtype = real_dtype(realsize)
        self.n_inputs = xbar_inputs if xbar_inputs > 0 else channels
        self.n_outputs = xbar_outputs if xbar_outputs > 0 else channels
        self.in_bytes = n_streams * filter_length * self.n_inputs * FORMAT_BYTES[in_format]
        self.out_bytes = n_streams * filter_length * self.n_outputs * FORMAT_BYTES[out_format]

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.bfir_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def is_initialized(self):
        return bool(self.lib.bfir_is_initialized(self.h))

    @staticmethod
    def _check_lengths(arrs, length):
        # the library reads `length` elements from EVERY array (brutefir::set_coeff takes one length for all,
        # brutefir.hpp:31-35): a shorter array would be read past its end
        for n, a in enumerate(arrs):
            if len(a) < length:
                raise ValueError("coefficient array %d has %d elements, length is %d" % (n, len(a), length))

    def set_coeff(self, coeffs, coeff_blocks, scale=1.0, length=None):
        """set_coeff(void **coeffs, n_coeffs, length, coeff_blocks, scale): returns 0 or -2."""
        arrs = [np.ascontiguousarray(c, dtype=self.dtype).ravel() for c in coeffs]
        length = (len(arrs[0]) if arrs else 0) if length is None else length
        self._check_lengths(arrs, length)
        ptrs = (ctypes.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
        rc = self.lib.bfir_set_coeff(self.h, ptrs, len(arrs), length, coeff_blocks, float(scale))
        if rc not in (OK, ERR_COEFF):
            raise BfirError(rc, last_error())
        return rc

    def set_coeff_crossfade(self, coeffs, coeff_blocks, scale=1.0, length=None):
        """stage a new coefficient set; the next run() cross-fades old -> new (crossfade_inplace ramp)"""
        arrs = [np.ascontiguousarray(c, dtype=self.dtype).ravel() for c in coeffs]
        length = len(arrs[0]) if length is None else length
        self._check_lengths(arrs, length)
        ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        rc = self.lib.bfir_set_coeff_crossfade(self.h, ptrs, len(arrs), length, coeff_blocks, float(scale))
        if rc not in (OK, ERR_COEFF):
            raise BfirError(rc, last_error())
        return rc

    def set_coeff_device(self, d_coeffs, channel_stride, n_coeffs, length, coeff_blocks, scale=1.0, crossfade=False):
        rc = self.lib.bfir_set_coeff_device(self.h, _ptr(d_coeffs), int(channel_stride), n_coeffs, length, coeff_blocks,
                                            float(scale), int(bool(crossfade)))
        if rc not in (OK, ERR_COEFF):
            raise BfirError(rc, last_error())
        return rc

    def set_crossbar(self, in_gains, out_gains):
        """in_gains [filters][inputs], out_gains [outputs][filters] (mixnscale scales, n_bufs > 1)"""
        a = np.ascontiguousarray(in_gains, dtype=np.float64)
        b = np.ascontiguousarray(out_gains, dtype=np.float64)
        assert a.shape == (self.channels, self.n_inputs) and b.shape == (self.n_outputs, self.channels)
        _check(self.lib.bfir_set_crossbar(self.h, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                          b.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))

    def run(self, inbuf, outbuf=None):
        """run(void *inbuf, void *outbuf) -> (rc, outbuf); rc is 0 or -1 like the reference."""
        if outbuf is None:
            outbuf = np.empty(self.out_bytes, dtype=np.uint8)
        rc = self.lib.bfir_run(self.h, _buf(inbuf, self.in_bytes, "inbuf"), _buf(outbuf, self.out_bytes, "outbuf"))
        if rc not in (OK, ERR_NONFINITE):
            raise BfirError(rc, last_error())
        return rc, outbuf

    def run_device(self, d_in, d_out):
        _check(self.lib.bfir_run_device(self.h, _buf(d_in, self.in_bytes, "d_in"), _buf(d_out, self.out_bytes, "d_out")))

    def run_device_pipelined(self, d_in, d_out):
        _check(self.lib.bfir_run_device_pipelined(self.h, _buf(d_in, self.in_bytes, "d_in"), _buf(d_out, self.out_bytes, "d_out")))

    def join(self):
        _check(self.lib.bfir_join(self.h))

    def run_device_pair(self, d_in0, d_in1, d_out0, d_out1, pipelined=False):
        """Two consecutive blocks with one partition-sum launch (offline / pipelined callers). `pipelined`: False/0 joined,
        True/1 no join between calls (stream-ordered with one group), "staged"/2 the engine's stage pipeline -- inputs
        must be complete at call time, outputs are visible after join() / sync() (include/bfir_b200.h)."""
        pipelined = 2 if pipelined == "staged" else int(pipelined)
        _check(self.lib.bfir_run_device_pair(self.h, _buf(d_in0, self.in_bytes, "d_in0"), _buf(d_in1, self.in_bytes, "d_in1"),
                                             _buf(d_out0, self.out_bytes, "d_out0"), _buf(d_out1, self.out_bytes, "d_out1"), pipelined))

    def run_device_quad(self, d_ins, d_outs, staged=False):
        """Four consecutive blocks with one partition-sum launch; staged=True: through the stage pipeline (inputs
        complete at call time, outputs visible after join() / sync())."""
        a = (_vp * 4)(*[_buf(x, self.in_bytes, "d_in") for x in d_ins])
        b = (_vp * 4)(*[_buf(x, self.out_bytes, "d_out") for x in d_outs])
        _check((self.lib.bfir_run_device_quad_staged if staged else self.lib.bfir_run_device_quad)(self.h, a, b))

    def run_device_oct(self, d_ins, d_outs, staged=False):
        """Eight consecutive blocks with one partition-sum launch, through the stage pipeline; staged=False joins at the
        end of the call, staged=True leaves the pipeline open (inputs complete at call time, outputs after join() / sync())."""
        a = (_vp * 8)(*[_buf(x, self.in_bytes, "d_in") for x in d_ins])
        b = (_vp * 8)(*[_buf(x, self.out_bytes, "d_out") for x in d_outs])
        _check(self.lib.bfir_run_device_oct(self.h, a, b, 1 if staged else 0))

    def run_partial_quad_device(self, d_ins):
        """partition shard with the fused reduce, four blocks: transforms, one four-block partition sum over the own
        partitions, pushes to the owners, arrival flag (bfir_run_partial_quad_device)"""
        a = (_vp * 4)(*[_buf(x, self.in_bytes, "d_in") for x in d_ins])
        _check(self.lib.bfir_run_partial_quad_device(self.h, a))

    def run_finish_quad_device(self, d_outs):
        """wait for every source rank's arrival flag on the device, then sum and emit the own channels of the four blocks"""
        b = (_vp * 4)(*[_ptr(x) for x in d_outs])
        _check(self.lib.bfir_run_finish_quad_device(self.h, b))

    def run_shard_quad_staged(self, d_ins, d_outs):
        """both halves of the four-block shard call through the stage pipeline (bfir_run_shard_quad_staged): inputs
        complete at call time, own-channel outputs visible after join() / sync()"""
        a = (_vp * 4)(*[_buf(x, self.in_bytes, "d_in") for x in d_ins])
        b = (_vp * 4)(*[_ptr(x) for x in d_outs])
        _check(self.lib.bfir_run_shard_quad_staged(self.h, a, b))

    def run_shard_oct_staged(self, d_ins, d_outs):
        """the eight-block shard call through the stage pipeline (bfir_run_shard_oct_staged)"""
        a = (_vp * 8)(*[_buf(x, self.in_bytes, "d_in") for x in d_ins])
        b = (_vp * 8)(*[_ptr(x) for x in d_outs])
        _check(self.lib.bfir_run_shard_oct_staged(self.h, a, b))

    def run_async_quad(self, ins, outs):
        """Four consecutive blocks of PINNED host buffers through the stage pipeline; returns the ticket of the fourth."""
        a = (_vp * 4)(*[_buf(x, self.in_bytes, "in") for x in ins])
        b = (_vp * 4)(*[_buf(x, self.out_bytes, "out") for x in outs])
        t = self.lib.bfir_run_async_quad(self.h, a, b)
        if t < 0:
            raise BfirError(int(t), last_error())
        return t

    def run_async_pair(self, in0, in1, out0, out1):
        """Two consecutive blocks of PINNED host buffers; returns the ticket of the second block."""
        t = self.lib.bfir_run_async_pair(self.h, _buf(in0, self.in_bytes, "in0"), _buf(in1, self.in_bytes, "in1"),
                                         _buf(out0, self.out_bytes, "out0"), _buf(out1, self.out_bytes, "out1"))
        if t < 0:
            raise BfirError(int(t), last_error())
        return t

    def run_async(self, inbuf, outbuf):
        """Queue one block on PINNED host buffers; returns a ticket for wait(). Buffers stay untouched until then."""
        t = self.lib.bfir_run_async(self.h, _buf(inbuf, self.in_bytes, "inbuf"), _buf(outbuf, self.out_bytes, "outbuf"))
        if t < 0:
            raise BfirError(int(t), last_error())
        return t

    def wait(self, ticket):
        """-> 0, or -1 when a NaN/Inf probe fired in a block since the last wait/sync."""
        rc = self.lib.bfir_wait(self.h, ticket)
        if rc not in (OK, ERR_NONFINITE):
            raise BfirError(rc, last_error())
        return rc

    def run_partial_device(self, d_in):
        _check(self.lib.bfir_run_partial_device(self.h, _buf(d_in, self.in_bytes, "d_in")))

    def run_finish_device(self, d_out):
        _check(self.lib.bfir_run_finish_device(self.h, _ptr(d_out)))

    def acc_device_ptr(self):
        n = ctypes.c_size_t()
        p = self.lib.bfir_acc_device_ptr(self.h, ctypes.byref(n))
        return p, n.value

    def sync(self):
        rc = self.lib.bfir_sync(self.h)
        if rc not in (OK, ERR_NONFINITE):
            raise BfirError(rc, last_error())
        return rc

    # fused partition-shard reduce (peer stores over NVLink instead of an NCCL reduce)
    def peer_setup(self, rank, world):
        _check(self.lib.bfir_peer_setup(self.h, rank, world))

    def peer_export(self):
        buf = ctypes.create_string_buffer(64)
        _check(self.lib.bfir_peer_export(self.h, buf))
        return bytes(buf.raw)

    def peer_import(self, peer_rank, handle):
        _check(self.lib.bfir_peer_import(self.h, peer_rank, ctypes.create_string_buffer(handle, 64)))

    def peer_set_ptr(self, peer_rank, ptr):
        _check(self.lib.bfir_peer_set_ptr(self.h, peer_rank, ctypes.c_void_p(ptr)))

    def peer_recv_ptr(self):
        return self.lib.bfir_peer_recv_ptr(self.h)

    def peer_own_channels(self):
        a, b = ctypes.c_int(), ctypes.c_int()
        _check(self.lib.bfir_peer_own_channels(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def set_groups(self, n):
        _check(self.lib.bfir_set_groups(self.h, int(n)))

    def get_groups(self):
        return _check(self.lib.bfir_get_groups(self.h))

    def set_coeff_map(self, mapping):
        """filter channel c uses coefficient set mapping[c]; None restores the identity (bfir_set_coeff_map)"""
        if mapping is None:
            return _check(self.lib.bfir_set_coeff_map(self.h, None, 0))
        m = [int(x) for x in mapping]
        return _check(self.lib.bfir_set_coeff_map(self.h, (ctypes.c_int * len(m))(*m), len(m)))

    def get_mac_split(self):
        return _check(self.lib.bfir_get_mac_split(self.h))

    def get_quad_split(self):
        return _check(self.lib.bfir_get_quad_split(self.h))

    def set_stream(self, cuda_stream):
        _check(self.lib.bfir_set_stream(self.h, ctypes.c_void_p(cuda_stream)))

    def reset(self):
        _check(self.lib.bfir_reset(self.h))

    def set_profiling(self, max_blocks):
        _check(self.lib.bfir_set_profiling(self.h, int(max_blocks)))

    def get_profile(self, reset=True):
        """-> ({'fwd_ms','mac_ms','inv_ms'} summed over `blocks` profiled block steps, blocks)"""
        ms = (ctypes.c_double * 3)()
        n = ctypes.c_ulonglong()
        _check(self.lib.bfir_get_profile(self.h, ms, ctypes.byref(n), int(reset)))
        return {"fwd_ms": ms[0], "mac_ms": ms[1], "inv_ms": ms[2]}, int(n.value)

    def get_mac_profile(self, blocks_per_launch, reset=True):
        """-> (summed ms, launches) of the profiled partition sums that covered `blocks_per_launch` blocks"""
        ms, n = ctypes.c_double(), ctypes.c_ulonglong()
        _check(self.lib.bfir_get_mac_profile(self.h, int(blocks_per_launch), ctypes.byref(ms), ctypes.byref(n), int(reset)))
        return float(ms.value), int(n.value)

    def check_overflows(self):
        return _check(self.lib.bfir_check_overflows(self.h))

    def overflow(self, channel):
        o = Overflow()
        _check(self.lib.bfir_get_overflow(self.h, channel, ctypes.byref(o)))
        return o

    def dither_ptr(self, channel):
        v = ctypes.c_int()
        _check(self.lib.bfir_get_dither_ptr(self.h, channel, ctypes.byref(v)))
        return v.value

    def blockcounter(self):
        v = ctypes.c_uint()
        _check(self.lib.bfir_get_blockcounter(self.h, ctypes.byref(v)))
        return v.value


class DeviceBuf:
    """A device buffer owned by a convolver (the reference's callers own host cbufs instead)."""

    def __init__(self, conv, nbytes):
        self.conv, self.nbytes = conv, nbytes
        self.ptr = conv.lib.bfir_conv_alloc(conv.h, nbytes)
        if not self.ptr:
            raise BfirError(ERR_CUDA, "bfir_conv_alloc(%d) failed" % nbytes)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        _check(self.conv.lib.bfir_conv_upload(self.conv.h, ctypes.c_void_p(self.ptr), _ptr(arr), arr.nbytes))
        return self

    def download(self, dtype, count=None):
        dtype = np.dtype(dtype)
        count = self.nbytes // dtype.itemsize if count is None else count
        out = np.empty(count, dtype=dtype)
        _check(self.conv.lib.bfir_conv_download(self.conv.h, _ptr(out), ctypes.c_void_p(self.ptr), out.nbytes))
        return out

    def offset(self, nbytes):
        return self.ptr + nbytes

    def free(self):
        if self.ptr:
            self.conv.lib.bfir_conv_free(self.conv.h, ctypes.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FftwConvolver:
    """``class fftw_convolver`` (reference brutefir/fftw_convolver.hpp:28-166) on device cbufs."""

    def __init__(self, length, realsize, n_dither_channels=1, sampling_rate=44100):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        rc = self.lib.bfir_conv_create(ctypes.byref(self.h), length, realsize, n_dither_channels, sampling_rate)
        if rc != OK:
            self.h = ctypes.c_void_p()
            raise BfirError(rc, last_error())
        self.length, self.n_fft, self.realsize = length, 2 * length, realsize
        self.dtype = real_dtype(realsize)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.bfir_conv_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # buffers
    def cbuf(self, data=None, n_cbufs=1.0):
        b = DeviceBuf(self, int(self.convolver_cbufsize() * n_cbufs))
        if data is not None:
            b.upload(np.ascontiguousarray(data, dtype=self.dtype))
        return b

    def rawbuf(self, data=None, nbytes=None):
        b = DeviceBuf(self, nbytes if nbytes is not None else np.asarray(data).nbytes)
        if data is not None:
            b.upload(data)
        return b

    def get(self, buf, count=None):
        return buf.download(self.dtype, count)

    # entry points, reference names
    def convolver_cbufsize(self):
        return self.lib.bfir_conv_cbufsize(self.h)

    def convolver_raw2cbuf(self, rawbuf, cbuf, next_cbuf, fmt, byte_offset, sample_spacing):
        _check(self.lib.bfir_conv_raw2cbuf(self.h, _ptr(rawbuf), _ptr(cbuf), _ptr(next_cbuf), fmt, byte_offset, sample_spacing))

    def convolver_time2freq(self, input_cbuf, output_cbuf):
        _check(self.lib.bfir_conv_time2freq(self.h, _ptr(input_cbuf), _ptr(output_cbuf)))

    def convolver_mixnscale(self, input_cbufs, output_cbuf, scales, mixmode):
        n = len(input_cbufs)
        arr = (ctypes.c_void_p * n)(*[b.ptr for b in input_cbufs])
        sc = (ctypes.c_double * n)(*[float(s) for s in scales])
        return self.lib.bfir_conv_mixnscale(self.h, arr, _ptr(output_cbuf), sc, n, mixmode)

    def convolver_convolve_inplace(self, cbuf, coeffs):
        _check(self.lib.bfir_conv_convolve_inplace(self.h, _ptr(cbuf), _ptr(coeffs)))

    def convolver_convolve(self, input_cbuf, coeffs, output_cbuf):
        _check(self.lib.bfir_conv_convolve(self.h, _ptr(input_cbuf), _ptr(coeffs), _ptr(output_cbuf)))

    def convolver_convolve_add(self, input_cbuf, coeffs, output_cbuf):
        _check(self.lib.bfir_conv_convolve_add(self.h, _ptr(input_cbuf), _ptr(coeffs), _ptr(output_cbuf)))

    def convolver_crossfade_inplace(self, input_cbuf, crossfade_cbuf, buffer_cbuf):
        _check(self.lib.bfir_conv_crossfade_inplace(self.h, _ptr(input_cbuf), _ptr(crossfade_cbuf), _ptr(buffer_cbuf)))

    def convolver_dirac_convolve(self, input_cbuf, output_cbuf):
        _check(self.lib.bfir_conv_dirac_convolve(self.h, _ptr(input_cbuf), _ptr(output_cbuf)))

    def convolver_dirac_convolve_inplace(self, cbuf):
        _check(self.lib.bfir_conv_dirac_convolve_inplace(self.h, _ptr(cbuf)))

    def convolver_freq2time(self, input_cbuf, output_cbuf):
        _check(self.lib.bfir_conv_freq2time(self.h, _ptr(input_cbuf), _ptr(output_cbuf)))

    def convolver_convolve_eval(self, input_cbuf, buffer_cbuf, output_cbuf):
        _check(self.lib.bfir_conv_convolve_eval(self.h, _ptr(input_cbuf), _ptr(buffer_cbuf), _ptr(output_cbuf)))

    # td_conv_t (fftw_convolver.cpp:698-777): the handle is an opaque pointer, freed with convolver_td_free
    def convolver_td_block_length(self, n_coeffs):
        return self.lib.bfir_conv_td_block_length(n_coeffs)

    def convolver_td_new(self, coeffs, n_coeffs=None):
        """Returns the td handle, or None where the reference returns NULL / has no defined result."""
        coeffs = np.ascontiguousarray(coeffs, dtype=self.dtype)
        n = len(coeffs) if n_coeffs is None else n_coeffs
        h = _vp()
        rc = self.lib.bfir_conv_td_new(self.h, ctypes.byref(h), _ptr(coeffs), n)
        if rc == ERR_INVALID:
            return None
        _check(rc)
        return h

    def convolver_td_coeffs(self, tdc):
        bl = self.lib.bfir_conv_td_blocklen(tdc)
        out = np.zeros(2 * bl, dtype=self.dtype)
        _check(self.lib.bfir_conv_download(self.h, _ptr(out), ctypes.c_void_p(self.lib.bfir_conv_td_coeffs(tdc)), out.nbytes))
        return out

    def convolver_td_convolve(self, tdc, overlap_block):
        _check(self.lib.bfir_conv_td_convolve(self.h, tdc, _ptr(overlap_block)))

    def convolver_td_free(self, tdc):
        self.lib.bfir_conv_td_free(tdc)

    def convolver_cbuf2raw(self, cbuf, outbuf, fmt, byte_offset, sample_spacing, apply_dither, dither_channel, overflow):
        _check(self.lib.bfir_conv_cbuf2raw(self.h, _ptr(cbuf), _ptr(outbuf), fmt, byte_offset, sample_spacing,
                                           int(bool(apply_dither)), dither_channel, ctypes.byref(overflow)))

    def convolver_coeffs2cbuf(self, coeffs, n_coeffs, scale, dest):
        """Returns 0, or -2 where the reference returns NULL (NaN/Inf among coefficients)."""
        coeffs = np.ascontiguousarray(coeffs, dtype=self.dtype)
        rc = self.lib.bfir_conv_coeffs2cbuf(self.h, _ptr(coeffs), n_coeffs, float(scale), _ptr(dest))
        if rc not in (OK, ERR_COEFF):
            raise BfirError(rc, last_error())
        return rc

    def convolver_runtime_coeffs2cbuf(self, src, dest):
        _check(self.lib.bfir_conv_runtime_coeffs2cbuf(self.h, _ptr(src), _ptr(dest)))

    def sync(self):
        _check(self.lib.bfir_conv_sync(self.h))

    # dither inspection
    def dither_table(self):
        n = _check(self.lib.bfir_conv_dither_table_size(self.h))
        out = np.empty(n, dtype=np.int8)
        _check(self.lib.bfir_conv_dither_table(self.h, _ptr(out), n))
        return out

    def dither_map(self):
        out = np.empty(512, dtype=self.dtype)
        _check(self.lib.bfir_conv_dither_map(self.h, _ptr(out)))
        return out

    def dither_ptr(self, channel):
        return _check(self.lib.bfir_conv_dither_ptr(self.h, channel))


class Equalizer:
    """``class equalizer`` (reference brutefir/equalizer.hpp:66-115) on the GPU, without the WAV cache."""

    def __init__(self, block_length, n_blocks, realsize, sampling_rate):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        rc = self.lib.bfir_eq_create(ctypes.byref(self.h), block_length, n_blocks, realsize, sampling_rate)
        if rc != OK:
            self.h = ctypes.c_void_p()
            raise BfirError(rc, last_error())
        self.taps = self.lib.bfir_eq_taps(self.h)
        self.realsize, self.dtype = realsize, real_dtype(realsize)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.bfir_eq_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _arr(v):
        return (ctypes.c_double * len(v))(*[float(x) for x in v])

    def generate(self, freq, mag, phase):
        """equalizer::generate: returns the taps/2-sample filter (host array)."""
        out = np.empty(self.taps // 2, dtype=self.dtype)
        _check(self.lib.bfir_eq_render(self.h, len(freq), self._arr(freq), self._arr(mag), self._arr(phase), _ptr(out)))
        return out

    def generate_device(self, freq, mag, phase):
        """same, but returns the device pointer of the rendered filter (for Brutefir.set_coeff_device)"""
        p = self.lib.bfir_eq_render_device(self.h, len(freq), self._arr(freq), self._arr(mag), self._arr(phase))
        if not p:
            raise BfirError(ERR_CUDA, last_error())
        return p
