"""ctypes binding of libmlmultipoles.so (the C ABI declared in include/mlmultipoles.h).

There is no CPU fallback: if the CUDA library is missing or no CUDA device is present,
every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLM_LIB") or os.path.join(_HERE, "libmlmultipoles.so")   # MLM_LIB: experiments only
# the same sources with -DMLM_STATS (event counters; measurement only: bench.py, tools/) -- never used by the API
STATS_LIB_PATH = os.path.join(_HERE, "libmlmultipoles_stats.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)
_D, _I, _L, _P = ctypes.c_double, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p

# name -> argument types (all return int unless noted); mirrors include/mlmultipoles.h
SIGNATURES = {
    "mlm_create": [_I, ctypes.POINTER(_P)],
    "mlm_destroy": [_P],
    "mlm_launch_count": [_P, ctypes.POINTER(_L)],
    "mlm_synchronize": [_P],
    "mlm_auto_strip": [_P, _L, _L],
    "mlm_host_register": [_P, _P, _L],
    "mlm_host_unregister": [_P, _P],
    "mlm_memcpy2d_async": [_P, _P, _L, _P, _L, _L, _L, _P],
    "mlm_alloc_host": [_P, _L, ctypes.POINTER(_P)],
    "mlm_free_host": [_P, _P],
    "mlm_device_alloc": [_P, _L, ctypes.POINTER(_P)],
    "mlm_device_free": [_P, _P],
    "mlm_ipc_export": [_P, _P, ctypes.c_char_p],
    "mlm_ipc_open": [_P, ctypes.c_char_p, ctypes.POINTER(_P)],
    "mlm_ipc_close": [_P, _P],
    "mlm_memcpy_async": [_P, _P, _P, _L, _P],
    "mlm_points": [_P, _D, _D, _D, _D, _P, _L, _I, _P, _P, _P, _P, _P, _P],
    "mlm_points_unordered": [_P, _D, _D, _D, _D, _P, _L, _I, _P, _P, _P, _P, _P, _P],
    "mlm_points_host": [_P, _D, _D, _D, _D, _P, _L, _I, _P, _P, _P, _P, _P],
    "mlm_map": [_P, _D, _D, _D, _D, _D, _D, _D, _D, _L, _L, _L, _I, _I, _P, _P, _P, _P, _P, _P],
    "mlm_map_deal": [_P, _D, _D, _D, _D, _D, _D, _D, _D, _L, _L, _L, _L, _I, ctypes.POINTER(ctypes.c_int32), _I, _I,
                     _P, _P, _P, _P, _P, _P],
    "mlm_map_strips": [_P, _D, _D, _D, _D, _D, _D, _D, _D, _L, _L, _L, _L, _L, _I, _I, _P, _P, _P, _P, _P, _P],
    "mlm_map_host": [_P, _D, _D, _D, _D, _D, _D, _D, _D, _L, _L, _L, _I, _I, _P, _P, _P, _P, _P],
    "mlm_auto_segment": [_P, _L, _L],
    "mlm_models": [_P, _P, _P, _P, _P, _P, _P, _L, _D, _D, _L, _I, _I, _P, _P, _P, _P, _P, _P],
    "mlm_models_host": [_P, _P, _P, _P, _P, _P, _P, _L, _D, _D, _L, _I, _I, _P, _P, _P, _P, _P],
    "mlm_models_times": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _L, _I, _I, _P, _P, _P, _P, _P, _P],
    "mlm_models_times_host": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _L, _I, _I, _P, _P, _P, _P, _P],
    "mlm_models_chi2_times": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _L, _I, _I, _P, _P, _P, _P],
    "mlm_models_chi2_times_host": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _L, _I, _I, _P, _P, _P],
    "mlm_models_chi2": [_P, _P, _P, _P, _P, _P, _P, _L, _D, _D, _L, _I, _I, _P, _P, _P, _P],
    "mlm_models_chi2_host": [_P, _P, _P, _P, _P, _P, _P, _L, _D, _D, _L, _I, _I, _P, _P, _P],
    "mlm_from_wk": [_P, _P, _L, _I, _D, _D, _P, _P],
    "mlm_from_wk_host": [_P, _P, _L, _I, _D, _D, _P],
    "mlm_solve": [_P, _D, _D, _P, _L, _P, _P],
    "mlm_solve_host": [_P, _D, _D, _P, _L, _P],
    "mlm_akl": [_P, _P, _P, _P, _L, _P, _P],
    "mlm_akl_host": [_P, _P, _P, _P, _L, _P],
    "mlm_fp64_peak": [_P, ctypes.POINTER(_D)],
    "mlm_fp64_probe": [_P, _I, _I, ctypes.POINTER(_D)],
    "mlm_store_probe": [_P, _I, _L, _L, _I, _P, _P, _P, _P, _P, _P],
    "mlm_kernel_info": [_P, ctypes.POINTER(ctypes.c_int32), _I],
    "mlm_stats_read": [_P, ctypes.POINTER(ctypes.c_uint64), _I, _I],
}
STRING_FUNCS = {"mlm_last_error": [_P], "mlm_version": [], "mlm_stats_names": []}

_libs: dict = {}
_lock = threading.Lock()
_contexts: dict = {}


class MlmError(RuntimeError):
    pass


def load_library(path: str | None = None):
    """dlopen libmlmultipoles.so (or the library at `path`) and declare every prototype.  Raises if it is not built."""
    path = LIB_PATH if path is None else path
    with _lock:
        if path in _libs:
            return _libs[path]
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is not built. Run `python -m microlensing_b200.build` (needs nvcc). "
                "microlensing_b200 has no CPU fallback.")
        lib = ctypes.CDLL(path)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        for name, argtypes in STRING_FUNCS.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_char_p
        _libs[path] = lib
        return lib


class Context:
    """One mlm_context (device, two streams, scratch) per (process, device)."""

    def __init__(self, device: int, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        self.device = device
        h = _P()
        rc = self.lib.mlm_create(device, ctypes.byref(h))
        if rc != 0:
            msg = self.lib.mlm_last_error(None).decode()
            raise MlmError(f"mlm_create(device={device}) failed ({rc}): {msg} -- a CUDA device is required, "
                           "there is no CPU fallback")
        self.handle = h
        self._finalizer = weakref.finalize(self, self.lib.mlm_destroy, h)

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.mlm_last_error(self.handle).decode()
            if rc < 0:
                raise ValueError(f"{what}: {msg}")
            raise MlmError(f"{what} failed (cudaError {rc}): {msg}")

    def launch_count(self) -> int:
        n = _L(0)
        self.check(self.lib.mlm_launch_count(self.handle, ctypes.byref(n)), "mlm_launch_count")
        return int(n.value)

    def synchronize(self):
        self.check(self.lib.mlm_synchronize(self.handle), "mlm_synchronize")

    def fp64_peak_tflops(self) -> float:
        v = _D(0.0)
        self.check(self.lib.mlm_fp64_peak(self.handle, ctypes.byref(v)), "mlm_fp64_peak")
        return float(v.value)

    def auto_strip(self, nx: int, nrows: int) -> int:
        """Strip length the map entry points choose for strip=0 on a launch of nx x nrows pixels."""
        v = int(self.lib.mlm_auto_strip(self.handle, int(nx), int(nrows)))
        if v <= 0:
            raise ValueError("mlm_auto_strip: bad argument")
        return v

    def auto_segment(self, M: int, T: int) -> int:
        """Epochs per tracking segment the model-grid entry points choose for seg=0 on a grid of M x T."""
        v = int(self.lib.mlm_auto_segment(self.handle, int(M), int(T)))
        if v <= 0:
            raise ValueError("mlm_auto_segment: bad argument")
        return v

    def host_register(self, address: int, nbytes: int):
        self.check(self.lib.mlm_host_register(self.handle, _P(address), int(nbytes)), "mlm_host_register")

    def host_unregister(self, address: int):
        self.check(self.lib.mlm_host_unregister(self.handle, _P(address)), "mlm_host_unregister")

    def memcpy2d_async(self, dst: int, dpitch: int, src: int, spitch: int, width: int, height: int, stream: int | None):
        self.check(self.lib.mlm_memcpy2d_async(self.handle, _P(dst), int(dpitch), _P(src), int(spitch), int(width),
                                               int(height), _P(stream)), "mlm_memcpy2d_async")

    def fp64_probe(self, mode: int, blocks_per_sm: int) -> float:
        v = _D(0.0)
        self.check(self.lib.mlm_fp64_probe(self.handle, mode, blocks_per_sm, ctypes.byref(v)), "mlm_fp64_probe")
        return float(v.value)

    def stats_read(self, reset: bool = True) -> dict:
        """Event counters of the -DMLM_STATS build: {event: (lanes, warp-level occurrences)}."""
        names = self.lib.mlm_stats_names().decode().split()
        a = (ctypes.c_uint64 * (2 * len(names)))()
        self.check(self.lib.mlm_stats_read(self.handle, a, 2 * len(names), int(reset)), "mlm_stats_read")
        return {n: (int(a[2 * i]), int(a[2 * i + 1])) for i, n in enumerate(names)}

    def kernel_info(self) -> dict:
        a = (ctypes.c_int32 * 4)()
        self.check(self.lib.mlm_kernel_info(self.handle, a, 4), "mlm_kernel_info")
        return {"k_points4_regs": a[0], "k_points4_local_bytes": a[1], "k_map4_regs": a[2], "k_map4_local_bytes": a[3]}

    # ---- device buffers shareable across the processes of one box (CUDA IPC) -----------------
    def device_alloc(self, nbytes: int) -> int:
        p = _P()
        self.check(self.lib.mlm_device_alloc(self.handle, int(nbytes), ctypes.byref(p)), "mlm_device_alloc")
        return int(p.value)

    def device_free(self, p: int):
        self.check(self.lib.mlm_device_free(self.handle, _P(p)), "mlm_device_free")

    def ipc_export(self, p: int) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self.check(self.lib.mlm_ipc_export(self.handle, _P(p), buf), "mlm_ipc_export")
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        p = _P()
        self.check(self.lib.mlm_ipc_open(self.handle, ctypes.c_char_p(handle), ctypes.byref(p)), "mlm_ipc_open")
        return int(p.value)

    def memcpy_async(self, dst: int, src: int, nbytes: int, stream: int | None):
        self.check(self.lib.mlm_memcpy_async(self.handle, _P(dst), _P(src), int(nbytes), _P(stream)), "mlm_memcpy_async")

    def ipc_close(self, p: int):
        self.check(self.lib.mlm_ipc_close(self.handle, _P(p)), "mlm_ipc_close")

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array over page-locked host memory (freed when the array and its views die)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        nbytes = max(n * dtype.itemsize, 1)
        p = _P()
        self.check(self.lib.mlm_alloc_host(self.handle, nbytes, ctypes.byref(p)), "mlm_alloc_host")
        buf = (ctypes.c_char * nbytes).from_address(p.value)
        weakref.finalize(buf, self.lib.mlm_free_host, self.handle, p)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def default_device() -> int:
    for key in ("MLM_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


def context(device: int | None = None, stats: bool = False) -> Context:
    """The process-wide context of `device`.  stats=True: a context of the counting build (measurement only)."""
    if device is None:
        device = default_device()
    key = (device, stats)
    with _lock:
        ctx = _contexts.get(key)
    if ctx is None:
        ctx = Context(device, STATS_LIB_PATH if stats else None)
        with _lock:
            _contexts.setdefault(key, ctx)
            ctx = _contexts[key]
    return ctx


def ptr(a) -> _P:
    """Address of a C-contiguous numpy array, a raw int device pointer, or None."""
    if a is None:
        return _P(None)
    if isinstance(a, np.ndarray):
        if not a.flags.c_contiguous:
            raise ValueError("array must be C-contiguous")
        return _P(a.ctypes.data)
    if isinstance(a, int):
        return _P(a)
    if hasattr(a, "data_ptr"):          # torch tensor (device or host); plumbing only
        return _P(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")
