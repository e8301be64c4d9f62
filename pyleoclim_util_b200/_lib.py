"""ctypes binding of lib/libwwz_b200.so (the C ABI declared in include/wwz_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libwwz_b200.so")

DENSE = 1
STANDARDIZE = 2
FAITHFUL = 4
NO_RECURRENCE = 8
EXACT_WINDOW = 16
UAR1 = 32
FOSTER = 64
FREQ_HZ = 128
PREPARED = 256

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)
_i64 = ctypes.c_int64
_u32 = ctypes.c_uint32
_u64 = ctypes.c_uint64
_vp = ctypes.c_void_p
_dbl = ctypes.c_double

_lib = None


class WwzB200Error(RuntimeError):
    pass


_SIGS = {
    "wwzb200_version": (ctypes.c_int, []),
    "wwzb200_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "wwzb200_set_device": (ctypes.c_int, [ctypes.c_int]),
    "wwzb200_get_device": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "wwzb200_sm_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "wwzb200_release": (ctypes.c_int, []),
    "wwzb200_alloc_pinned": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "wwzb200_free_pinned": (ctypes.c_int, [_vp]),
    "wwzb200_copy_to_host": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "wwzb200_copy_to_host_async": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    "wwzb200_stream_synchronize": (ctypes.c_int, [_vp]),
    "wwzb200_host_register": (ctypes.c_int, [_vp, ctypes.c_size_t]),
    "wwzb200_host_unregister": (ctypes.c_int, [_vp]),
    "wwzb200_kirchner": (ctypes.c_int, [_dp, _dp, _i64, _dp, _i64, _dp, _i64, _dbl, _i64, _u32,
                                        _dp, _dp, _dp, _dp, _dp, _dp]),
    "wwzb200_kirchner_psd": (ctypes.c_int, [_dp, _dp, _i64, _dp, _i64, _dp, _i64, _dbl, _i64, _i64, _u32,
                                            _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "wwzb200_wwa2psd": (ctypes.c_int, [_dp, _dp, _i64, _i64, _dbl, _i64, _i64, _dp]),
    "wwzb200_kirchner_shared_axis": (ctypes.c_int, [_dp, _dp, _i64, _i64, _dp, _i64, _dp, _i64, _dbl, _i64, _i64,
                                                    _u32, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "wwzb200_kirchner_ragged": (ctypes.c_int, [_dp, _dp, _ip, _dp, _ip, _dp, _ip, _ip, _i64, _dbl, _i64, _i64,
                                               _u32, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "wwzb200_kirchner_members": (ctypes.c_int, [_dp, _dp, _i64, _i64, _i64, _dp, _i64, _dp, _i64, _dbl, _i64, _i64, _u32,
                                                _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "wwzb200_kirchner_members_dev": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _dbl, _i64, _i64,
                                                    _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wwzb200_ar1_surrogates": (ctypes.c_int, [_dp, _i64, _dbl, _dbl, _dbl, _i64, _u64, _dp]),
    "wwzb200_uar1_surrogates": (ctypes.c_int, [_dp, _i64, _dbl, _dbl, _dbl, _i64, _u64, _dp]),
    "wwzb200_quantiles": (ctypes.c_int, [_dp, _i64, _i64, _dp, _i64, _dp]),
    "wwzb200_ar1_signif": (ctypes.c_int, [_dp, _i64, _dbl, _dbl, _dbl, _i64, _u64, _dp, _i64, _dp, _i64, _dbl, _i64,
                                          _i64, _u32, _dp, _i64, _dp, _dp, _dp]),
    "wwzb200_kirchner_dev": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _dbl, _i64, _i64, _u32,
                                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "wwzb200_wwa2psd_dev": (ctypes.c_int, [_vp, _vp, _i64, _i64, _dbl, _i64, _i64, _vp, _vp]),
    "wwzb200_wwa2psd_parts_dev": (ctypes.c_int, [_vp, _vp, _i64, _i64, _dbl, _i64, _i64, _vp, _vp]),
    "wwzb200_kirchner_shared_axis_dev": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _dbl, _i64,
                                                        _i64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wwzb200_ar1_surrogates_dev": (ctypes.c_int, [_vp, _i64, _dbl, _dbl, _dbl, _i64, _u64, _vp, _vp]),
    "wwzb200_uar1_surrogates_dev": (ctypes.c_int, [_vp, _i64, _dbl, _dbl, _dbl, _i64, _u64, _vp, _vp]),
    "wwzb200_surrogates_block_dev": (ctypes.c_int, [_vp, _i64, _dbl, _dbl, _dbl, _i64, _i64, _u64, _u32, _vp, _vp]),
    "wwzb200_series_pad": (ctypes.c_int, [_i64, ctypes.POINTER(_i64)]),
    "wwzb200_shared_axis_cells_dev": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _dbl, _i64, _u32,
                                                     _vp, _vp, _vp]),
    "wwzb200_quantiles_cells_dev": (ctypes.c_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp]),
    "wwzb200_quantiles_dev": (ctypes.c_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "wwzb200_shared_axis_prepare_dev": (ctypes.c_int, [_vp, _i64, _i64, _vp, _i64, _vp, _i64, _dbl, _i64, _u32, _vp,
                                                       ctypes.POINTER(ctypes.c_int), _vp]),
    "wwzb200_shared_axis_cells_scatter_dev": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _dbl, _i64, _u32,
                                                             _vp, ctypes.POINTER(_vp), _i64, _i64, _i64, _i64, _vp]),
    "wwzb200_ipc_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(_vp), ctypes.c_char_p]),
    "wwzb200_ipc_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "wwzb200_ipc_close": (ctypes.c_int, [_vp]),
    "wwzb200_ipc_free": (ctypes.c_int, [_vp]),
    "wwzb200_wtc": (ctypes.c_int, [_dp, _dp, _dp, _dp, _i64, _i64, _i64, _dp, _dbl, _dbl, _i64, _dp, _dp, _dp]),
    "wwzb200_wtc_dev": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _dbl, _dbl, _i64, _vp, _vp, _vp, _vp]),
    "wwzb200_quantiles_segments_dev": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp]),
    "wwzb200_measure_fp64_peak": (ctypes.c_int, [_dp, _dp]),
    "wwzb200_launch_count": (_i64, []),
    "wwzb200_profile": (ctypes.c_int, [ctypes.c_int]),
    "wwzb200_profile_read": (ctypes.c_int, [_dp, ctypes.POINTER(_i64), _dp]),
    "wwzb200_profile_walked": (ctypes.c_int, [_dp]),
}


def lib():
    """Load the shared library (raises if it has not been built: there is no CPU path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WwzB200Error(
                "%s not found; build it with `python -m pyleoclim_util_b200.build` "
                "(pyleoclim_util_b200 has no CPU fallback)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        handle.wwzb200_last_error.restype = ctypes.c_char_p
        handle.wwzb200_last_error.argtypes = []
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)      # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().wwzb200_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("wwz_b200: " + msg)
        raise WwzB200Error("wwz_b200 error %d: %s" % (rc, msg))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)


def device_count():
    n = ctypes.c_int()
    check(lib().wwzb200_device_count(ctypes.byref(n)))
    return n.value


def set_device(i):
    check(lib().wwzb200_set_device(int(i)))


def sm_count():
    n = ctypes.c_int()
    check(lib().wwzb200_sm_count(ctypes.byref(n)))
    return n.value


def launch_count():
    return int(lib().wwzb200_launch_count())


def measure_fp64_peak():
    """FP64 FMA peak of the current device (TFLOP/s, kernel ms) from the DFMA-chain microbenchmark."""
    tf, ms = ctypes.c_double(), ctypes.c_double()
    check(lib().wwzb200_measure_fp64_peak(ctypes.byref(tf), ctypes.byref(ms)))
    return tf.value, ms.value


def profile(enable):
    check(lib().wwzb200_profile(1 if enable else 0))


def profile_read():
    """(summed cell-kernel ms, launches, cell-points) recorded since profile(True)."""
    ms, n, cp = ctypes.c_double(), _i64(), ctypes.c_double()
    check(lib().wwzb200_profile_read(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(cp)))
    return ms.value, int(n.value), cp.value


def profile_walked():
    """(warp, point) pairs evaluated by the tau-recurrence kernel since profile(True)."""
    v = ctypes.c_double()
    check(lib().wwzb200_profile_walked(ctypes.byref(v)))
    return v.value


def _free_pinned(addr):
    if _lib is not None:
        _lib.wwzb200_free_pinned(_vp(addr))


class _PinnedPool:
    """Recycles page-locked output buffers.  A result array handed to the caller is backed by pinned
    memory (the device-to-host copy is then a plain DMA: ~0.5 ms instead of ~10 ms for the 24 MB of a
    1000 x 500 grid); when the caller drops the array its buffer returns to the pool, so steady-state
    loops neither re-allocate nor alias live results."""

    def __init__(self, cap_bytes=2 << 30, max_item=512 << 20):
        self.free = {}
        self.pooled = 0
        self.cap, self.max_item = cap_bytes, max_item

    def take(self, nbytes):
        lst = self.free.get(nbytes)
        if lst:
            self.pooled -= nbytes
            return lst.pop()
        p = _vp()
        check(lib().wwzb200_alloc_pinned(max(nbytes, 1), ctypes.byref(p)))
        return p.value

    def give(self, addr, nbytes):
        if self.pooled + nbytes <= self.cap:
            self.free.setdefault(nbytes, []).append(addr)
            self.pooled += nbytes
        else:
            _free_pinned(addr)

    def clear(self):
        for lst in self.free.values():
            for addr in lst:
                _free_pinned(addr)
        self.free.clear()
        self.pooled = 0


_pool = _PinnedPool()


def pinned_empty(shape, dtype=np.float64):
    """NumPy array backed by page-locked host memory; the buffer goes back to a pool when the array
    (and every view of it) has been collected."""
    shape = tuple(int(s) for s in np.atleast_1d(shape))
    dt = np.dtype(dtype)
    count = int(np.prod(shape, dtype=np.int64))
    nbytes = max(count * dt.itemsize, 1)
    addr = _pool.take(nbytes)
    buf = (ctypes.c_char * nbytes).from_address(addr)
    weakref.finalize(buf, _pool.give, addr, nbytes)
    return np.frombuffer(buf, dtype=dt, count=count).reshape(shape)


def output_empty(shape):
    """Result buffer: pinned (pooled) for anything worth a DMA, plain NumPy for tiny or huge arrays."""
    nbytes = int(np.prod(shape, dtype=np.int64)) * 8
    if (64 << 10) <= nbytes <= _pool.max_item:
        return pinned_empty(shape)
    return np.empty(shape)
