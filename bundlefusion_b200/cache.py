"""Host-side mirror of ``CUDACache`` (FL/CUDACache.{h,cpp}): same constructor arguments, ``storeFrame`` / ``getCacheFramesGPU`` /
``getWidth`` / ``getHeight`` / ``getIntrinsics`` semantics; buffers allocated as ``CUDACache::alloc`` does and handed to the C-ABI
(include/bf_cache.h) as raw device pointers.  The object can be passed to ``CUDASolverBundling.solve(cudaCache=...)``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import BFCacheParams, BFCUDACachedFrame

_FIELDS = (("d_depthDownsampled", 1, "float32"), ("d_cameraposDownsampled", 4, "float32"), ("d_intensityDownsampled", 1, "float32"),
           ("d_intensityDerivsDownsampled", 2, "float32"), ("d_normalsDownsampledUCHAR4", 4, "uint8"), ("d_normalsDownsampled", 4, "float32"))


class CUDACache:
    def __init__(self, widthDepthInput: int, heightDepthInput: int, widthDownSampled: int, heightDownSampled: int, maxNumImages: int,
                 inputIntrinsics, device="cuda:0", colorDownSigma: float = 2.5, depthDownSigmaD: float = 1.0, depthDownSigmaR: float = 0.05):
        import torch
        self._torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CUDACache needs a CUDA device (no CPU fallback)")
        self.lib = capi.lib()
        self.width, self.height, self.m_maxNumImages = widthDownSampled, heightDownSampled, maxNumImages
        K = np.asarray(inputIntrinsics, np.float32).reshape(4, 4)
        Kc = K.copy()                                                   # cpp:20-24
        Kc[0, 0] *= np.float32(widthDownSampled) / np.float32(widthDepthInput)
        Kc[1, 1] *= np.float32(heightDownSampled) / np.float32(heightDepthInput)
        Kc[0, 2] *= np.float32(widthDownSampled - 1) / np.float32(widthDepthInput - 1)
        Kc[1, 2] *= np.float32(heightDownSampled - 1) / np.float32(heightDepthInput - 1)
        self.m_intrinsics = Kc
        self.intrinsics = (float(Kc[0, 0]), float(Kc[1, 1]), float(Kc[0, 2]), float(Kc[1, 2]))
        p = BFCacheParams()
        p.inputDepthWidth, p.inputDepthHeight, p.width, p.height = widthDepthInput, heightDepthInput, widthDownSampled, heightDownSampled
        Kinv = intrinsics_inverse(K)
        for k in range(16):
            p.inputIntrinsicsInv[k] = float(Kinv.reshape(-1)[k])
        p.filterIntensitySigma, p.filterDepthSigmaD, p.filterDepthSigmaR = colorDownSigma, depthDownSigmaD, depthDownSigmaR
        self.params = p
        n = widthDownSampled * heightDownSampled
        self._bufs, self._frames = [], (BFCUDACachedFrame * maxNumImages)()
        for k in range(maxNumImages):
            fr = {}
            for name, ch, dt in _FIELDS:
                t = torch.zeros(n * ch, dtype=getattr(torch, dt), device=self.device)
                fr[name] = t
                setattr(self._frames[k], name, t.data_ptr())
            self._bufs.append(fr)
        raw = np.frombuffer(bytes(self._frames), dtype=np.uint8).copy()
        self.d_cache = torch.from_numpy(raw).to(self.device)            # m_cache on the device (cpp: alloc)
        self.m_currentFrame = 0

    def storeFrame(self, d_depth, inputDepthWidth, inputDepthHeight, d_color, inputColorWidth, inputColorHeight):
        """CUDACache::storeFrame (cpp:45-86).  d_depth float32 [H,W] (-inf invalid), d_color uint8 [H,W,4].  Asynchronous."""
        t = self._torch
        t.cuda.set_device(self.device)
        self.lib.bfSetStream(C.c_void_p(t.cuda.current_stream(self.device).cuda_stream))
        if self.m_currentFrame >= self.m_maxNumImages:
            raise RuntimeError("CUDACache is full")
        p = self.params
        assert (inputDepthWidth, inputDepthHeight) == (p.inputDepthWidth, p.inputDepthHeight)
        p.inputColorWidth, p.inputColorHeight = inputColorWidth, inputColorHeight
        capi.check(self.lib.bfCacheStoreFrame(C.byref(p), d_depth.data_ptr(), d_color.data_ptr(), C.byref(self._frames[self.m_currentFrame])), "bfCacheStoreFrame")
        self.m_currentFrame += 1

    def getCacheFramesGPU(self):
        return self.d_cache

    def getWidth(self): return self.width
    def getHeight(self): return self.height
    def getIntrinsics(self): return self.m_intrinsics
    def getNumFrames(self): return self.m_currentFrame

    def download(self, k: int) -> dict:
        """Frame k as host arrays keyed like synth.make_cache_frame."""
        w, h = self.width, self.height
        b = self._bufs[k]
        g = lambda name, ch: b[name].cpu().numpy().reshape((h, w, ch) if ch > 1 else (h, w))
        return {"depth": g("d_depthDownsampled", 1), "campos": g("d_cameraposDownsampled", 4), "normals": g("d_normalsDownsampled", 4),
                "normalsU": g("d_normalsDownsampledUCHAR4", 4), "intensity": g("d_intensityDownsampled", 1), "intensityDerivs": g("d_intensityDerivsDownsampled", 2)}


def intrinsics_inverse(K) -> np.ndarray:
    """mat4f::getInverse of an intrinsics matrix, as FL/CUDACache.cpp:25, 38 and FL/Bundler.cpp:25 form it: the general 4x4 inverse (bfMat4Inverse -- the reference's formula bit for
    bit).  The closed form 1 / fx, -mx / fx is NOT the same float for every calibration (fy / (fx * fy) rounds differently from 1 / fx)."""
    from .scene_rep import mat4_inverse_f32
    return mat4_inverse_f32(np.asarray(K, np.float32).reshape(4, 4))
