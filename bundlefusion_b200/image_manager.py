"""Host-side mirror of the device part of ``CUDAImageManager::process`` (FL/CUDAImageManager.cpp:22-158): raw sensor depth / colour on
the device in, integration-resolution depth / colour out, through the C-ABI of include/bf_ingest.h."""
from __future__ import annotations

import ctypes as C

from . import _capi as capi
from ._capi import BFIngestParams


class CUDAImageManager:
    def __init__(self, widthIntegration: int, heightIntegration: int, device="cuda:0", erodeSIFTdepth: bool = True, depthFilter: bool = True,
                 depthSigmaD: float = 2.0, depthSigmaR: float = 0.05):
        import torch
        self._torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CUDAImageManager needs a CUDA device (no CPU fallback)")
        self.lib = capi.lib()
        self.m_widthIntegration, self.m_heightIntegration = widthIntegration, heightIntegration
        self._erode, self._filter, self._sD, self._sR = erodeSIFTdepth, depthFilter, depthSigmaD, depthSigmaR
        self.m_currFrame = 0

    def process(self, d_depthRaw, d_colorRaw):
        """d_depthRaw float32 [H,W] (-inf invalid), d_colorRaw uint8 [CH,CW,4] cuda tensors -> (depth [hi,wi], colour [hi,wi,4]).  Asynchronous."""
        t = self._torch
        t.cuda.set_device(self.device)
        self.lib.bfSetStream(C.c_void_p(t.cuda.current_stream(self.device).cuda_stream))
        p = BFIngestParams()
        p.depthHeight, p.depthWidth = d_depthRaw.shape[:2]
        p.colorHeight, p.colorWidth = d_colorRaw.shape[:2]
        p.widthIntegration, p.heightIntegration = self.m_widthIntegration, self.m_heightIntegration
        p.erodeIterations, p.erodeStructureSize, p.erodeDThresh, p.erodeFracReq = (2 if self._erode else 0), 3, 0.05, 0.3
        p.depthSigmaD, p.depthSigmaR = (self._sD if self._filter else 0.0), self._sR
        dout = t.full((self.m_heightIntegration, self.m_widthIntegration), float("-inf"), dtype=t.float32, device=self.device)
        cout = t.zeros((self.m_heightIntegration, self.m_widthIntegration, 4), dtype=t.uint8, device=self.device)
        capi.check(self.lib.bfIngestFrame(C.byref(p), d_depthRaw.data_ptr(), d_colorRaw.data_ptr(), dout.data_ptr(), cout.data_ptr()), "bfIngestFrame")
        self.m_currFrame += 1
        return dout, cout
