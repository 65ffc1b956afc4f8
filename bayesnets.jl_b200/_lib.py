"""ctypes binding of libbn_cuda.so (include/bn_cuda.h).  No CPU fallback: if the CUDA library is missing
or no GPU is visible, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbn_cuda.so")

BN_OK, BN_BAD_ARG, BN_DIM_MISMATCH, BN_NOT_IN_FACTOR, BN_CUDA_ERR = 0, -1, -2, -3, -4
BN_ZERO_WEIGHT, BN_OOM, BN_UNSUPPORTED, BN_BOUNDS, BN_NCCL_ERR = -5, -6, -7, -8, -9
BN_MAX_CARD = 255
BN_MAX_EVIDENCE_STATE = 127
BN_COMM_ID_BYTES = 128
DTYPE_F64, DTYPE_I64 = 0, 1
REDUCE_SUM, REDUCE_MAX = 0, 1
OP_MUL, OP_DIV, OP_ADD, OP_SUB, OP_MAX, OP_MIN = range(6)


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch (src/Factors/factors_dims.jl:197-199)."""


class BNCudaError(RuntimeError):
    pass


_lib = None

u8p, i8p, i32p, i64p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int8), C.POINTER(C.c_int32),
                              C.POINTER(C.c_int64), C.POINTER(C.c_double))
h64 = C.c_uint64

_SIGS = {
    "bn_last_error": (C.c_char_p, []),
    "bn_version": (C.c_int32, []),
    "bn_device_count": (C.c_int32, [i32p]),
    "bn_set_stream": (C.c_int32, [C.c_void_p]),
    "bn_launch_count": (C.c_uint64, []),
    "bn_pool_trim": (C.c_int32, [C.POINTER(C.c_uint64)]),
    "bn_synchronize": (C.c_int32, []),
    "bn_dev_alloc": (C.c_int32, [C.c_int32, C.c_uint64, C.POINTER(C.c_void_p)]),
    "bn_dev_free": (C.c_int32, [C.c_int32, C.c_void_p]),
    "bn_dev_memset": (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint64]),
    "bn_dev_upload": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "bn_dev_download": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "bn_comm_init_all": (C.c_int32, [i32p, C.c_int32, C.POINTER(h64)]),
    "bn_comm_unique_id": (C.c_int32, [u8p]),
    "bn_comm_init_rank": (C.c_int32, [u8p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(h64)]),
    "bn_comm_info": (C.c_int32, [h64, i32p, i32p, i32p]),
    "bn_comm_destroy": (C.c_int32, [h64]),
    "bn_comm_shard": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bn_comm_allreduce": (C.c_int32, [h64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]),
    "bn_comm_allreduce_dev": (C.c_int32, [h64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]),
    "bn_comm_barrier": (C.c_int32, [h64]),
    "bn_net_create_comm": (C.c_int32, [C.c_int32, i32p, i32p, i32p, i64p, f64p, h64, C.POINTER(h64)]),
    "bn_spec_cache_stats": (C.c_int32, [C.POINTER(C.c_uint64)]),
    "bn_net_create": (C.c_int32, [C.c_int32, i32p, i32p, i32p, i64p, f64p, C.c_int32, C.POINTER(h64)]),
    "bn_net_destroy": (C.c_int32, [h64]),
    "bn_lw_infer": (C.c_int32, [h64, i8p, i32p, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64, f64p, f64p]),
    "bn_lw_infer_dev": (C.c_int32, [h64, i8p, i32p, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "bn_set_lw_mode": (C.c_int32, [C.c_int32, C.c_int32]),
    "bn_lw_spec_compile_count": (C.c_uint64, []),
    "bn_set_gibbs_mode": (C.c_int32, [C.c_int32]),
    "bn_lw_spec_source": (C.c_int32, [C.c_int32, i32p, i32p, i32p, i64p, f64p, i8p, i32p, C.c_int32, C.c_int32,
                                      C.c_char_p, C.c_int64, i64p, i64p, i32p]),
    "bn_gibbs_spec_source": (C.c_int32, [C.c_int32, i32p, i32p, i32p, i64p, f64p, i8p, i32p, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_char_p, C.c_int64, i64p, i64p]),
    "bn_sample": (C.c_int32, [h64, i8p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64, u8p, f64p, u8p]),
    "bn_sample_dev": (C.c_int32, [h64, i8p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "bn_gibbs_infer": (C.c_int32, [h64, i8p, i32p, C.c_int32, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64,
                                   C.c_int64, C.c_int32, C.c_uint64, u8p, C.c_int32, f64p, f64p]),
    "bn_gibbs_infer_dev": (C.c_int32, [h64, i8p, i32p, C.c_int32, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64,
                                       C.c_int64, C.c_int32, C.c_uint64, C.c_void_p]),
    "bn_gibbs_sample": (C.c_int32, [h64, i8p, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, i32p,
                                    C.c_uint64, u8p, u8p]),
    "bn_factor_upload": (C.c_int32, [C.c_int32, i32p, i32p, f64p, C.c_int32, C.POINTER(h64)]),
    "bn_factor_alloc": (C.c_int32, [C.c_int32, i32p, i32p, C.c_int32, C.POINTER(h64)]),
    "bn_factor_from_cpd": (C.c_int32, [h64, C.c_int32, C.POINTER(h64)]),
    "bn_factor_relabel": (C.c_int32, [h64, i32p]),
    "bn_factor_ndims": (C.c_int32, [h64, i32p]),
    "bn_factor_shape": (C.c_int32, [h64, i32p, i32p]),
    "bn_factor_length": (C.c_int32, [h64, i64p]),
    "bn_factor_download": (C.c_int32, [h64, f64p]),
    "bn_factor_dev_ptr": (C.c_int32, [h64, C.POINTER(C.c_void_p)]),
    "bn_factor_free": (C.c_int32, [h64]),
    "bn_factor_product": (C.c_int32, [h64, h64, C.POINTER(h64)]),
    "bn_factor_sum": (C.c_int32, [h64, i32p, C.c_int32, C.POINTER(h64)]),
    "bn_factor_slice": (C.c_int32, [h64, i32p, i32p, C.c_int32, C.POINTER(h64)]),
    "bn_factor_setindex": (C.c_int32, [h64, i32p, i32p, C.c_int32, f64p, C.c_int64]),
    "bn_factor_permutedims": (C.c_int32, [h64, i32p, C.c_int32, C.POINTER(h64)]),
    "bn_factor_push": (C.c_int32, [h64, C.c_int32, C.c_int32, C.POINTER(h64)]),
    "bn_factor_normalize": (C.c_int32, [h64, C.c_int32]),
    "bn_factor_norm": (C.c_int32, [h64, C.c_int32, f64p]),
    "bn_factor_div_scalar": (C.c_int32, [h64, C.c_double]),
    "bn_factor_eliminate": (C.c_int32, [C.POINTER(h64), C.c_int32, i32p, C.c_int32, C.POINTER(h64)]),
    "bn_factor_join": (C.c_int32, [h64, h64, C.c_int32, C.POINTER(h64)]),
    "bn_factor_reduce": (C.c_int32, [h64, i32p, C.c_int32, C.c_int32, C.POINTER(h64)]),
    "bn_factor_normalize_dims": (C.c_int32, [h64, i32p, C.c_int32, C.c_int32]),
    "bn_factor_broadcast": (C.c_int32, [h64, i32p, C.c_int32, f64p, i32p, C.c_int32]),
    "bn_exact_infer": (C.c_int32, [h64, i8p, i32p, C.c_int32, C.c_int32, i32p, i32p, f64p]),
    "bn_statistics": (C.c_int32, [h64, u8p, C.c_uint64, i64p]),
    "bn_statistics_dev": (C.c_int32, [h64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "bn_logpdf": (C.c_int32, [h64, u8p, C.c_uint64, f64p]),
    "bn_logpdf_dev": (C.c_int32, [h64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "bn_logpdf_counts_dev": (C.c_int32, [h64, C.c_void_p, C.c_void_p]),
    "bn_loopy_infer": (C.c_int32, [h64, i8p, i32p, C.c_uint64, C.c_int32, C.c_double, C.c_int32, C.c_int32, f64p, i32p]),
    "bn_loopy_infer_dev": (C.c_int32, [h64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.c_double, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p]),
}

EXPORTS = tuple(_SIGS)


def lib() -> C.CDLL:
    """Load libbn_cuda.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BNCudaError(
                f"{LIB_PATH} is missing: build it with `python bayesnets.jl_b200/build.py` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    """Map bn_status to the exception the reference would have thrown (include/bn_cuda.h)."""
    if rc == BN_OK:
        return
    msg = lib().bn_last_error().decode("utf-8", "replace")
    if rc == BN_DIM_MISMATCH:
        raise DimensionMismatch(msg)
    if rc in (BN_BAD_ARG, BN_NOT_IN_FACTOR):
        raise ValueError(msg)  # ArgumentError
    if rc == BN_BOUNDS:
        raise IndexError(msg)  # BoundsError
    if rc == BN_ZERO_WEIGHT:
        raise RuntimeError(msg)
    if rc == BN_OOM:
        raise MemoryError(msg)
    if rc == BN_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == BN_NCCL_ERR:
        raise BNCudaError(f"NCCL: {msg}")
    raise BNCudaError(f"libbn_cuda status {rc}: {msg}")


def ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def set_stream(cuda_stream: int | None) -> None:
    """Enqueue subsequent calls from this thread on `cuda_stream` (e.g. torch.cuda.current_stream().cuda_stream)."""
    check(lib().bn_set_stream(C.c_void_p(cuda_stream or 0)))


_LW_MODES = {"auto": 0, "generic": 1, "specialized": 2}


def set_lw_mode(mode: str = "auto", prune: bool = True) -> None:
    """Pick the LW kernel for this thread: "auto" (network-specialised NVRTC kernel when it can be built, generic
    interpreter otherwise), "generic", or "specialized" (raise if unavailable).  `prune` lets the specialised kernel
    skip barren nodes (never changes a count bit; see csrc/specialize.cu)."""
    check(lib().bn_set_lw_mode(C.c_int32(_LW_MODES[mode]), C.c_int32(1 if prune else 0)))


def set_gibbs_mode(mode: str = "auto") -> None:
    """Pick the Gibbs inference kernel for this thread: "auto", "generic" or "specialized" (csrc/specialize.cu)."""
    check(lib().bn_set_gibbs_mode(C.c_int32(_LW_MODES[mode])))


def launch_count() -> int:
    return int(lib().bn_launch_count())


def pool_trim() -> int:
    """Give the device memory parked by freed factors / temporaries back to the driver; returns the bytes released."""
    n = C.c_uint64(0)
    check(lib().bn_pool_trim(C.byref(n)))
    return int(n.value)


class DeviceBuffer:
    """A plain device allocation (bn_dev_alloc) for the *_dev entry points: what the host mirror uses where a caller
    without a CUDA array library needs device-resident results.  Freed by close() / garbage collection."""

    def __init__(self, nbytes: int, device: int = 0, zero: bool = False):
        import weakref
        p = C.c_void_p(0)
        check(lib().bn_dev_alloc(C.c_int32(device), C.c_uint64(int(nbytes)), C.byref(p)))
        self.ptr, self.nbytes, self.device = int(p.value or 0), int(nbytes), int(device)
        self._fin = weakref.finalize(self, DeviceBuffer._free, self.device, self.ptr)
        if zero:
            self.zero_()

    @staticmethod
    def _free(device, p):
        try:
            lib().bn_dev_free(C.c_int32(device), C.c_void_p(p))
        except Exception:
            pass

    def close(self):
        self._fin()

    def zero_(self):
        check(lib().bn_dev_memset(C.c_void_p(self.ptr), C.c_int32(0), C.c_uint64(self.nbytes)))
        return self

    def upload(self, arr):
        import numpy as np
        a = np.ascontiguousarray(arr)
        if a.nbytes > self.nbytes:
            raise ValueError(f"upload of {a.nbytes} bytes into a {self.nbytes}-byte device buffer")
        check(lib().bn_dev_upload(C.c_void_p(self.ptr), C.c_void_p(a.ctypes.data), C.c_uint64(a.nbytes)))
        return self

    def download(self, dtype, count: int | None = None):
        """Copy back as a numpy array (synchronises the calling thread's current stream)."""
        import numpy as np
        dt = np.dtype(dtype)
        n = self.nbytes // dt.itemsize if count is None else int(count)
        out = np.empty(n, dt)
        check(lib().bn_dev_download(C.c_void_p(out.ctypes.data), C.c_void_p(self.ptr), C.c_uint64(n * dt.itemsize)))
        return out


def synchronize() -> None:
    """Wait for everything this thread enqueued on its current stream (the *_dev entry points are asynchronous)."""
    check(lib().bn_synchronize())


def spec_cache_stats() -> dict:
    """Counters of the network-specialised kernel cache since load: NVRTC compiles (of which on a background thread),
    cubins found in the on-disk cache."""
    v = (C.c_uint64 * 4)()
    check(lib().bn_spec_cache_stats(v))
    return {"nvrtc_compiles": int(v[0]), "background_compiles": int(v[1]), "disk_hits": int(v[2]), "kernels_loaded": int(v[3])}


def device_count() -> int:
    n = C.c_int32(0)
    check(lib().bn_device_count(C.byref(n)))
    return n.value


def device_count_or_zero() -> int:
    """device_count() that reports 0 instead of raising on a machine without a CUDA driver / device."""
    try:
        return device_count()
    except BNCudaError:
        return 0
