"""Host-side handle of the `.sens` reader / writer (include/bf_sens.h; csrc/sens_io.cu): the recorded-sequence container the reference's SensorDataReader
plays back (external/mLib ext-depthcamera/sensorData.h, FL/SensorDataReader.cpp).  Host code only -- no CUDA device is needed to read or write a file."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi

COLOR_RAW, COLOR_PNG, COLOR_JPEG = 0, 1, 2
DEPTH_RAW_USHORT, DEPTH_ZLIB_USHORT, DEPTH_OCCI_USHORT = 0, 1, 2


class BFSensHeader(C.Structure):
    _fields_ = [("version", C.c_uint32), ("sensorName", C.c_char * 256),
                ("colorIntrinsic", C.c_float * 16), ("colorExtrinsic", C.c_float * 16), ("depthIntrinsic", C.c_float * 16), ("depthExtrinsic", C.c_float * 16),
                ("colorCompression", C.c_int32), ("depthCompression", C.c_int32),
                ("colorWidth", C.c_uint32), ("colorHeight", C.c_uint32), ("depthWidth", C.c_uint32), ("depthHeight", C.c_uint32),
                ("depthShift", C.c_float), ("numFrames", C.c_uint64), ("numIMUFrames", C.c_uint64)]


def _bind():
    L = capi.lib()
    if getattr(L, "_sens_bound", False):
        return L
    vp = C.c_void_p
    L.bfSensOpen.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(BFSensHeader)]
    L.bfSensReadFrame.argtypes = [vp, C.c_uint64, vp, vp, vp, vp]
    L.bfSensReadFrameRaw.argtypes = [vp, C.c_uint64, vp, vp]
    L.bfSensClose.argtypes = [vp]; L.bfSensClose.restype = None
    L.bfSensCreate.argtypes = [C.c_char_p, C.POINTER(BFSensHeader), C.POINTER(vp)]
    L.bfSensAppendFrame.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_uint64]
    L.bfSensFinish.argtypes = [vp]
    L.bfSensDecodeJpeg.argtypes = [vp, C.c_size_t, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.bfSensDecodePng.argtypes = [vp, C.c_size_t, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.bfSensErrorString.argtypes = [C.c_int]; L.bfSensErrorString.restype = C.c_char_p
    L._sens_bound = True
    return L


def _check(L, rc, what):
    if rc != 0:
        raise RuntimeError(f"bundlefusion_b200: {what}: {L.bfSensErrorString(rc).decode()}")


def _decode(fn, what, data: bytes) -> np.ndarray:
    L = _bind()
    w, h = C.c_uint32(0), C.c_uint32(0)
    buf = np.frombuffer(data, np.uint8)
    _check(L, fn(L)(buf.ctypes.data, len(data), None, C.byref(w), C.byref(h)), what)
    out = np.zeros((h.value, w.value, 3), np.uint8)
    _check(L, fn(L)(buf.ctypes.data, len(data), out.ctypes.data, C.byref(w), C.byref(h)), what)
    return out


def decode_jpeg(data: bytes) -> np.ndarray:
    return _decode(lambda L: L.bfSensDecodeJpeg, "bfSensDecodeJpeg", data)


def decode_png(data: bytes) -> np.ndarray:
    return _decode(lambda L: L.bfSensDecodePng, "bfSensDecodePng", data)


class SensorDataReader:
    """what FL/SensorDataReader.cpp gives the frame loop: ``frame(i)`` -> depth float32 [H, W] in metres (-inf invalid), colour uint8 [H, W, 4], pose, time stamps"""

    def __init__(self, path: str):
        self.L = _bind()
        self.header = BFSensHeader()
        h = C.c_void_p()
        _check(self.L, self.L.bfSensOpen(path.encode(), C.byref(h), C.byref(self.header)), f"bfSensOpen({path})")
        self._h = h

    def __len__(self):
        return int(self.header.numFrames)

    def frame(self, i: int, pinned: bool = False):
        hd = self.header
        if pinned:
            import torch
            depth = torch.empty((hd.depthHeight, hd.depthWidth), dtype=torch.float32).pin_memory(); color = torch.empty((hd.colorHeight, hd.colorWidth, 4), dtype=torch.uint8).pin_memory()
            dp, cp = depth.data_ptr(), color.data_ptr()
        else:
            depth = np.zeros((hd.depthHeight, hd.depthWidth), np.float32); color = np.zeros((hd.colorHeight, hd.colorWidth, 4), np.uint8)
            dp, cp = depth.ctypes.data, color.ctypes.data
        pose = np.zeros((4, 4), np.float32); ts = np.zeros(2, np.uint64)
        _check(self.L, self.L.bfSensReadFrame(self._h, i, dp, cp, pose.ctypes.data, ts.ctypes.data), f"bfSensReadFrame({i})")
        return depth, color, pose, ts

    def frame_pose(self, i: int) -> np.ndarray:
        """the recorded camera-to-world pose of frame i (no pixel decoding)"""
        pose = np.zeros((4, 4), np.float32)
        _check(self.L, self.L.bfSensReadFrame(self._h, i, None, None, pose.ctypes.data, None), f"bfSensReadFrame({i})")
        return pose

    def frame_raw(self, i: int):
        hd = self.header
        depth = np.zeros((hd.depthHeight, hd.depthWidth), np.uint16); color = np.zeros((hd.colorHeight, hd.colorWidth, 3), np.uint8)
        _check(self.L, self.L.bfSensReadFrameRaw(self._h, i, depth.ctypes.data, color.ctypes.data), f"bfSensReadFrameRaw({i})")
        return depth, color

    def close(self):
        if getattr(self, "_h", None):
            self.L.bfSensClose(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SensorDataWriter:
    def __init__(self, path: str, width: int, height: int, intrinsic, depth_shift: float = 1000.0, zlib_depth: bool = True, sensor_name: str = "bundlefusion_b200"):
        self.L = _bind()
        hd = BFSensHeader()
        hd.version = 4; hd.sensorName = sensor_name.encode()[:255]
        K = np.ascontiguousarray(intrinsic, np.float32).reshape(16); I = np.eye(4, dtype=np.float32).reshape(16)
        for k in range(16):
            hd.colorIntrinsic[k] = hd.depthIntrinsic[k] = float(K[k]); hd.colorExtrinsic[k] = hd.depthExtrinsic[k] = float(I[k])
        hd.colorCompression = COLOR_RAW; hd.depthCompression = DEPTH_ZLIB_USHORT if zlib_depth else DEPTH_RAW_USHORT
        hd.colorWidth = hd.depthWidth = width; hd.colorHeight = hd.depthHeight = height; hd.depthShift = depth_shift
        self.header = hd
        h = C.c_void_p()
        _check(self.L, self.L.bfSensCreate(path.encode(), C.byref(hd), C.byref(h)), f"bfSensCreate({path})")
        self._h = h

    def append(self, depth_ushort: np.ndarray, color_rgb: np.ndarray, pose=None, ts_color: int = 0, ts_depth: int = 0):
        d = np.ascontiguousarray(depth_ushort, np.uint16); c = np.ascontiguousarray(color_rgb, np.uint8)
        p = np.ascontiguousarray(pose, np.float32) if pose is not None else None
        _check(self.L, self.L.bfSensAppendFrame(self._h, d.ctypes.data, c.ctypes.data, p.ctypes.data if p is not None else None, ts_color, ts_depth), "bfSensAppendFrame")

    def finish(self):
        if self._h:
            _check(self.L, self.L.bfSensFinish(self._h), "bfSensFinish")
            self._h = None
