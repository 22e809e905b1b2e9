"""Host-side handle of the frame-loop object (include/bf_frameloop.h; csrc/frame_loop.cu): one ``step`` per sensor frame runs what
FriedLiver's frame callback runs (FL/DepthSensing/DepthSensing.cpp:966-1129: ingest, OnlineBundler::processInput, reintegrate(),
integration of the current frame, OnlineBundler::process).  The sequencing is C++ inside the library; this file only carries pointers."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import BFHashParams


class BFFrameLoopParams(C.Structure):
    _fields_ = [("depthWidth", C.c_uint32), ("depthHeight", C.c_uint32), ("colorWidth", C.c_uint32), ("colorHeight", C.c_uint32),
                ("integrationWidth", C.c_uint32), ("integrationHeight", C.c_uint32), ("siftWidth", C.c_uint32), ("siftHeight", C.c_uint32),
                ("depthIntrinsics", C.c_float * 16), ("colorIntrinsics", C.c_float * 16),
                ("submapSize", C.c_uint32), ("maxNumImages", C.c_uint32), ("maxNumKeysPerImage", C.c_uint32), ("maxNumFrames", C.c_uint32), ("maxGlobalResiduals", C.c_uint32),
                ("numLocalNonLinIterations", C.c_uint32), ("numLocalLinIterations", C.c_uint32), ("numGlobalNonLinIterations", C.c_uint32), ("numGlobalLinIterations", C.c_uint32),
                ("numOptPerResidualRemoval", C.c_uint32),
                ("sensorDepthMin", C.c_float), ("sensorDepthMax", C.c_float), ("minKeyScale", C.c_float), ("featureCountThreshold", C.c_int32),
                ("siftMatchThresh", C.c_float), ("siftMatchRatioMaxLocal", C.c_float), ("siftMatchRatioMaxGlobal", C.c_float),
                ("minNumMatchesLocal", C.c_uint32), ("minNumMatchesGlobal", C.c_uint32),
                ("maxKabschResidual2", C.c_float), ("surfAreaPcaThresh", C.c_float),
                ("projCorrDistThres", C.c_float), ("projCorrNormalThres", C.c_float), ("projCorrColorThresh", C.c_float),
                ("verifySiftErrThresh", C.c_float), ("verifySiftCorrThresh", C.c_float), ("verifyOptErrThresh", C.c_float), ("verifyOptCorrThresh", C.c_float),
                ("optMaxResThresh", C.c_float),
                ("useLocalVerify", C.c_int32), ("useLocalDense", C.c_int32), ("useComprehensiveFrameInvalidation", C.c_int32),
                ("downsampledWidth", C.c_uint32), ("downsampledHeight", C.c_uint32),
                ("colorDownSigma", C.c_float), ("depthDownSigmaD", C.c_float), ("depthDownSigmaR", C.c_float),
                ("erodeSIFTdepth", C.c_int32), ("depthFilter", C.c_int32), ("depthSigmaD", C.c_float), ("depthSigmaR", C.c_float),
                ("maxFrameFixes", C.c_uint32), ("topNActive", C.c_uint32), ("minPoseDistSqrt", C.c_float),
                ("reconstructionEnabled", C.c_int32),
                ("hash", BFHashParams),
                ("renderDepthMin", C.c_float), ("renderDepthMax", C.c_float)]


class BFFrameLoopStatus(C.Structure):
    _fields_ = [("frame", C.c_uint32), ("validTransform", C.c_int32), ("globalTrackingLost", C.c_int32), ("transform", C.c_float * 16),
                ("numKeyPoints", C.c_uint32), ("lastMatchedFrame", C.c_int32), ("numLocalCorrespondences", C.c_uint32), ("numReintegrated", C.c_uint32),
                ("localSolved", C.c_int32), ("localValid", C.c_int32), ("numKeyframes", C.c_uint32), ("numGlobalCorrespondences", C.c_uint32),
                ("globalSolved", C.c_int32), ("globalRemoved", C.c_int32), ("numOptimizedFrames", C.c_uint32)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "transform"}
        d["transform"] = np.array(list(self.transform), np.float32).reshape(4, 4)
        return d


def default_params(width: int, height: int) -> BFFrameLoopParams:
    L = _bind(capi.lib())
    p = BFFrameLoopParams()
    L.bfFrameLoopDefaultParams(C.byref(p), width, height)
    return p


def _bind(L):
    if getattr(L, "_frameloop_bound", False):
        return L
    vp = C.c_void_p
    L.bfFrameLoopDefaultParams.argtypes = [C.POINTER(BFFrameLoopParams), C.c_uint32, C.c_uint32]
    L.bfFrameLoopDefaultParams.restype = None
    L.bfFrameLoopCreate.argtypes = [C.POINTER(BFFrameLoopParams), C.POINTER(vp)]
    L.bfFrameLoopDestroy.argtypes = [vp]
    L.bfFrameLoopDestroy.restype = None
    L.bfFrameLoopStep.argtypes = [vp, vp, vp, C.c_int, C.POINTER(BFFrameLoopStatus)]
    L.bfFrameLoopStepAhead.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.POINTER(BFFrameLoopStatus)]
    L.bfFrameLoopStepPastEnd.argtypes = [vp, C.POINTER(BFFrameLoopStatus)]
    L.bfFrameLoopGetTrajectory.argtypes = [vp, vp, C.c_uint]
    L.bfFrameLoopGetTrajectory.restype = C.c_uint
    L.bfFrameLoopGetHashData.argtypes = [vp]
    L.bfFrameLoopGetHashData.restype = C.POINTER(capi.BFHashDataStruct)
    L.bfFrameLoopGetHashParams.argtypes = [vp]
    L.bfFrameLoopGetHashParams.restype = C.POINTER(BFHashParams)
    L.bfFrameLoopGetCounters.argtypes = [vp, C.c_ulonglong * 8]
    L.bfFrameLoopGetCounters.restype = None
    L.bfFrameLoopSetOverlap.argtypes = [vp, C.c_int]
    L.bfFrameLoopJoin.argtypes = [vp]
    L.bfFrameLoopSetProfiling.argtypes = [vp, C.c_int]
    L.bfFrameLoopGetStageTimes.argtypes = [vp, C.c_double * 8]
    L.bfFrameLoopGetStageTimes.restype = C.c_ulonglong
    L._frameloop_bound = True
    return L


class FrameLoop:
    """One frame loop on one device.  ``step(depth, color)``: torch cuda tensors (float32 [H,W], uint8 [H,W,4]) or pinned / plain host tensors."""

    def __init__(self, params: BFFrameLoopParams, device="cuda:0"):
        import torch
        self._torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FrameLoop needs a CUDA device (no CPU fallback)")
        self.lib = _bind(capi.lib())
        self.params = params
        torch.cuda.set_device(self.device)
        self.lib.bfSetStream(C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        h = C.c_void_p()
        capi.check(self.lib.bfFrameLoopCreate(C.byref(params), C.byref(h)), "bfFrameLoopCreate")
        self._h = h

    def _bind_stream(self):
        t = self._torch
        t.cuda.set_device(self.device)
        self.lib.bfSetStream(C.c_void_p(t.cuda.current_stream(self.device).cuda_stream))

    def step(self, depth, color, next_depth=None, next_color=None) -> BFFrameLoopStatus:
        """one frame; with (next_depth, next_color) -- the frame the NEXT call will pass -- that frame's upload, ingest, SIFT detection and dense cache
        are queued on the loop's feature stream during this call (bfFrameLoopStepAhead; same results)"""
        self._bind_stream()
        st = BFFrameLoopStatus()
        on_host = 0 if depth.is_cuda else 1
        if next_depth is None:
            capi.check(self.lib.bfFrameLoopStep(self._h, C.c_void_p(depth.data_ptr()), C.c_void_p(color.data_ptr()), on_host, C.byref(st)), "bfFrameLoopStep")
        else:
            if next_depth.is_cuda != depth.is_cuda:
                raise ValueError("the announced frame must live where the current one does (both device or both host)")
            capi.check(self.lib.bfFrameLoopStepAhead(self._h, C.c_void_p(depth.data_ptr()), C.c_void_p(color.data_ptr()), C.c_void_p(next_depth.data_ptr()),
                                                     C.c_void_p(next_color.data_ptr()), on_host, C.byref(st)), "bfFrameLoopStepAhead")
        return st

    def step_past_end(self) -> BFFrameLoopStatus:
        self._bind_stream()
        st = BFFrameLoopStatus()
        capi.check(self.lib.bfFrameLoopStepPastEnd(self._h, C.byref(st)), "bfFrameLoopStepPastEnd")
        return st

    def trajectory(self, n: int) -> np.ndarray:
        out = np.zeros((n, 4, 4), np.float32)
        m = self.lib.bfFrameLoopGetTrajectory(self._h, out.ctypes.data, n)
        return out[:m]

    def counters(self) -> dict:
        c = (C.c_ulonglong * 8)()
        self.lib.bfFrameLoopGetCounters(self._h, c)
        return dict(zip(("frames", "integrations", "reintegrations", "local_solves", "global_solves", "global_pcg_iters", "host_syncs", "keyframes"), [int(x) for x in c]))

    STAGES = ("upload_ingest", "sift_detect", "dense_cache_and_count", "match_filters_sift_pose", "tsdf_reintegrate_integrate", "local_solve", "fuse_keyframe_match", "global_solve_trajectory")

    def set_overlap(self, enable: bool) -> bool:
        """TSDF work on the loop's second stream (the reference's reconstruction thread); returns the previous setting"""
        return bool(self.lib.bfFrameLoopSetOverlap(self._h, 1 if enable else 0))

    def join(self) -> None:
        self._bind_stream()
        capi.check(self.lib.bfFrameLoopJoin(self._h), "bfFrameLoopJoin")

    def set_profiling(self, enable: bool) -> None:
        capi.check(self.lib.bfFrameLoopSetProfiling(self._h, 1 if enable else 0), "bfFrameLoopSetProfiling")

    def stage_times(self) -> dict:
        """mean milliseconds per step and stage (device time line) since set_profiling(True)"""
        t = (C.c_double * 8)()
        n = int(self.lib.bfFrameLoopGetStageTimes(self._h, t))
        return {"steps": n, **{k: round(t[i] / max(1, n), 4) for i, k in enumerate(self.STAGES)}}

    def heap_free(self) -> int:
        out = C.c_uint(0)
        capi.check(self.lib.bfTsdfGetHeapFreeCount(self.lib.bfFrameLoopGetHashData(self._h), C.byref(out)), "bfTsdfGetHeapFreeCount")
        return out.value

    def close(self):
        if getattr(self, "_h", None):
            self.lib.bfFrameLoopDestroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
