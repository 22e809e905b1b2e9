"""Minimal device-memory helper over the CUDA runtime the product library is linked against (ctypes; no torch), for GPU tests that only
need malloc / memcpy around C-ABI calls."""
import ctypes as C

import numpy as np

from bundlefusion_b200 import _capi as capi

_rt = None


def runtime():
    global _rt
    if _rt is None:
        capi.lib()                                   # pulls libcudart.so.12 into the process via the library's own RUNPATH
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        rt.cudaFree.argtypes = [C.c_void_p]
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
        rt.cudaGetDeviceCount.argtypes = [C.POINTER(C.c_int)]
        rt.cudaDeviceSynchronize.argtypes = []
        _rt = rt
    return _rt


def device_count() -> int:
    n = C.c_int(0)
    try:
        rc = runtime().cudaGetDeviceCount(C.byref(n))
    except OSError:
        return 0
    return n.value if rc == 0 else 0


class DevBuf:
    """A device allocation holding a copy of a numpy array."""

    def __init__(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self.shape, self.dtype, self.nbytes = arr.shape, arr.dtype, max(arr.nbytes, 1)
        p = C.c_void_p()
        rc = runtime().cudaMalloc(C.byref(p), self.nbytes)
        if rc != 0:
            raise RuntimeError(f"cudaMalloc failed: {rc}")
        self.ptr = p.value
        if arr.nbytes:
            rc = runtime().cudaMemcpy(self.ptr, arr.ctypes.data, arr.nbytes, 1)
            if rc != 0:
                raise RuntimeError(f"cudaMemcpy H2D failed: {rc}")

    def get(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        if out.nbytes:
            rc = runtime().cudaMemcpy(out.ctypes.data, self.ptr, out.nbytes, 2)        # synchronises with the default stream
            if rc != 0:
                raise RuntimeError(f"cudaMemcpy D2H failed: {rc}")
        return out

    def __del__(self):
        if getattr(self, "ptr", None) and _rt is not None:
            _rt.cudaFree(self.ptr)
            self.ptr = None
