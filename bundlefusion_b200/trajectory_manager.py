"""Host-side mirror of the reference's ``TrajectoryManager`` (FL/TrajectoryManager.{h,cpp}) over the C-ABI of include/bf_bundler.h.

The state machine itself is host C++ inside libbundlefusion_b200.so (csrc/trajectory_host.cu); this class only forwards, with the
reference's method names and argument meaning.  It needs no GPU (the reference class does no device work apart from one cudaMemcpy of
the trajectory, which is the caller's here).
"""
import ctypes as C

import numpy as np

from . import _capi

INTEGRATED, NOT_INTEGRATED_NO_TRANSFORM, NOT_INTEGRATED_WITH_TRANSFORM, INVALID, REINTEGRATION = range(5)      # TrajectoryManager.h:8-14

_bound = False


def _lib():
    global _bound
    L = _capi.lib()
    if not _bound:
        vp, u, f = C.c_void_p, C.c_uint, C.c_float
        L.bfTrajectoryCreate.argtypes = [u, u, f]; L.bfTrajectoryCreate.restype = vp
        L.bfTrajectoryDestroy.argtypes = [vp]; L.bfTrajectoryDestroy.restype = None
        L.bfTrajectoryAddFrame.argtypes = [vp, C.c_int, vp, u]; L.bfTrajectoryAddFrame.restype = None
        L.bfTrajectoryUpdateOptimizedTransform.argtypes = [vp, vp, u]; L.bfTrajectoryUpdateOptimizedTransform.restype = None
        L.bfTrajectoryGenerateUpdateLists.argtypes = [vp]; L.bfTrajectoryGenerateUpdateLists.restype = None
        L.bfTrajectoryConfirmIntegration.argtypes = [vp, u]; L.bfTrajectoryConfirmIntegration.restype = None
        L.bfTrajectoryGetTopFromReIntegrateList.argtypes = [vp, vp, vp, vp]; L.bfTrajectoryGetTopFromReIntegrateList.restype = C.c_int
        L.bfTrajectoryGetTopFromIntegrateList.argtypes = [vp, vp, vp]; L.bfTrajectoryGetTopFromIntegrateList.restype = C.c_int
        L.bfTrajectoryGetTopFromDeIntegrateList.argtypes = [vp, vp, vp]; L.bfTrajectoryGetTopFromDeIntegrateList.restype = C.c_int
        for n in ("bfTrajectoryGetNumOptimizedFrames", "bfTrajectoryGetNumAddedFrames", "bfTrajectoryGetNumActiveOperations"):
            getattr(L, n).argtypes = [vp]; getattr(L, n).restype = u
        L.bfTrajectoryGetFrameType.argtypes = [vp, u]; L.bfTrajectoryGetFrameType.restype = C.c_int
        L.bfTrajectoryGetFrameDist.argtypes = [vp, u]; L.bfTrajectoryGetFrameDist.restype = f
        L.bfTrajectoryGetOptimizedTransforms.argtypes = [vp, vp]; L.bfTrajectoryGetOptimizedTransforms.restype = u
        _bound = True
    return L


class TrajectoryManager:
    """``TrajectoryManager(numMaxImage)``; s_topNActive / s_minPoseDistSqrt (TrajectoryManager.cpp:14-15, GlobalAppState) are explicit."""

    def __init__(self, numMaxImage: int, topNActive: int = 10, minPoseDistSqrt: float = 0.0):
        self._L = _lib()
        self.numMaxImage = int(numMaxImage)
        self._h = self._L.bfTrajectoryCreate(self.numMaxImage, int(topNActive), float(minPoseDistSqrt))
        if not self._h:
            raise RuntimeError("bfTrajectoryCreate failed")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.bfTrajectoryDestroy(self._h)
            self._h = None

    @staticmethod
    def _m(T):
        T = np.ascontiguousarray(T, np.float32).reshape(16)
        return T

    def addFrame(self, what: int, transform, idx: int):
        if not 0 <= idx < self.numMaxImage:
            raise IndexError(idx)
        T = self._m(transform)
        self._L.bfTrajectoryAddFrame(self._h, int(what), T.ctypes.data, idx)

    def updateOptimizedTransform(self, trajectory, numFrames: int):
        """`trajectory`: host array [>= min(numFrames, added)][4][4] (the reference copies it from the device itself)."""
        T = np.ascontiguousarray(trajectory, np.float32)
        if T.size < 16 * min(numFrames, self.getNumAddedFrames()):
            raise ValueError("trajectory shorter than numFrames")
        self._L.bfTrajectoryUpdateOptimizedTransform(self._h, T.ctypes.data, int(numFrames))

    def generateUpdateLists(self):
        self._L.bfTrajectoryGenerateUpdateLists(self._h)

    def confirmIntegration(self, frameIdx: int):
        self._L.bfTrajectoryConfirmIntegration(self._h, int(frameIdx))

    def getTopFromReIntegrateList(self):
        """(oldTransform, newTransform, frameIdx) or None."""
        o, n, i = np.zeros(16, np.float32), np.zeros(16, np.float32), C.c_uint(0)
        if not self._L.bfTrajectoryGetTopFromReIntegrateList(self._h, o.ctypes.data, n.ctypes.data, C.addressof(i)):
            return None
        return o.reshape(4, 4), n.reshape(4, 4), i.value

    def getTopFromIntegrateList(self):
        t, i = np.zeros(16, np.float32), C.c_uint(0)
        if not self._L.bfTrajectoryGetTopFromIntegrateList(self._h, t.ctypes.data, C.addressof(i)):
            return None
        return t.reshape(4, 4), i.value

    def getTopFromDeIntegrateList(self):
        t, i = np.zeros(16, np.float32), C.c_uint(0)
        if not self._L.bfTrajectoryGetTopFromDeIntegrateList(self._h, t.ctypes.data, C.addressof(i)):
            return None
        return t.reshape(4, 4), i.value

    def getNumOptimizedFrames(self) -> int:
        return self._L.bfTrajectoryGetNumOptimizedFrames(self._h)

    def getNumAddedFrames(self) -> int:
        return self._L.bfTrajectoryGetNumAddedFrames(self._h)

    def getNumActiveOperations(self) -> int:
        return self._L.bfTrajectoryGetNumActiveOperations(self._h)

    def getOptimizedTransforms(self) -> np.ndarray:
        n = min(self.getNumAddedFrames(), self.getNumOptimizedFrames())
        out = np.zeros((max(n, 1), 4, 4), np.float32)
        n = self._L.bfTrajectoryGetOptimizedTransforms(self._h, out.ctypes.data)
        return out[:n]

    # introspection (tests)
    def frameType(self, idx: int) -> int:
        return self._L.bfTrajectoryGetFrameType(self._h, int(idx))

    def frameDist(self, idx: int) -> float:
        return self._L.bfTrajectoryGetFrameDist(self._h, int(idx))
