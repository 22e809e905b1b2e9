"""Host-side mirror of the reference's ``CUDARayCastSDF`` (FL/DepthSensing/CUDARayCastSDF.{h,cpp}) over the C-ABI of include/bf_raycast.h: the fused model
seen from a pose as depth / camera-space position / normal / colour images.  Buffers are allocated here as ``RayCastData::allocate`` does
(FL/DepthSensing/RayCastSDFUtil.h:56-61) and handed to the library as raw device pointers; the interval images the reference gets back from Direct3D 11
are two more device images filled by the library (bfRayCastSplat)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import BFRayCastData, BFRayCastParams


def ray_cast_params(width: int, height: int, fx: float, fy: float, mx: float, my: float, min_depth: float = 0.1, max_depth: float = 4.0, truncation: float = 0.06,
                    ray_increment_factor: float = 0.8, thres_sample_dist_factor: float = 50.5, thres_dist_factor: float = 50.0, use_gradients: bool = False,
                    num_sdf_blocks: int = 200000) -> BFRayCastParams:
    """CUDARayCastSDF::parametersFromGlobalAppState (h:24-50) with the values of zParametersDefault.txt:35-36, 53-56"""
    p = BFRayCastParams()
    p.m_width, p.m_height = width, height
    p.fx, p.fy, p.mx, p.my = fx, fy, mx, my
    p.m_minDepth, p.m_maxDepth = min_depth, max_depth
    p.m_rayIncrement = np.float32(ray_increment_factor) * np.float32(truncation)
    p.m_thresSampleDist = np.float32(thres_sample_dist_factor) * np.float32(p.m_rayIncrement)
    p.m_thresDist = np.float32(thres_dist_factor) * np.float32(p.m_rayIncrement)
    p.m_useGradients = 1 if use_gradients else 0
    p.m_maxNumVertices = num_sdf_blocks * 6
    p.m_splatMinimum = 1
    return p


class CUDARayCastSDF:
    def __init__(self, params: BFRayCastParams, device="cuda:0"):
        import torch
        self._torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CUDARayCastSDF needs a CUDA device (no CPU fallback)")
        self.lib = capi.lib()
        self.m_params = BFRayCastParams()
        C.memmove(C.byref(self.m_params), C.byref(params), C.sizeof(BFRayCastParams))
        H, W = params.m_height, params.m_width
        kw = dict(device=self.device, dtype=torch.float32)
        self.d_depth = torch.empty(H, W, **kw); self.d_depth4 = torch.empty(H, W, 4, **kw)
        self.d_normals = torch.empty(H, W, 4, **kw); self.d_colors = torch.empty(H, W, 4, **kw)
        self.d_rayMin = torch.empty(H, W, **kw); self.d_rayMax = torch.empty(H, W, **kw)
        d = BFRayCastData()
        d.d_depth, d.d_depth4, d.d_normals, d.d_colors = self.d_depth.data_ptr(), self.d_depth4.data_ptr(), self.d_normals.data_ptr(), self.d_colors.data_ptr()
        d.d_vertexBuffer = None
        d.d_rayIntervalSplatMin, d.d_rayIntervalSplatMax = self.d_rayMin.data_ptr(), self.d_rayMax.data_ptr()
        self.m_data = d
        L = self.lib
        vp = C.c_void_p
        L.bfRayCastRenderPose.argtypes = [vp, vp, vp, vp, vp, vp]
        L.bfRayCastSplat.argtypes = [vp, vp, vp, vp, vp]
        L.bfRayCastRender.argtypes = [vp, vp, vp, vp]
        L.bfRayCastComputeNormals.argtypes = [vp, C.c_uint, C.c_uint]

    def _bind_stream(self):
        t = self._torch
        t.cuda.set_device(self.device)
        self.lib.bfSetStream(C.c_void_p(t.cuda.current_stream(self.device).cuda_stream))

    def render(self, hashData, hashParams, cam, lastRigidTransform) -> None:
        """CUDARayCastSDF::render (cpp:42-73): interval splat from the hash's last compactified list, ray march, normals.  Nothing returns to the host."""
        self._bind_stream()
        T = np.ascontiguousarray(lastRigidTransform, np.float32).reshape(16)
        capi.check(self.lib.bfRayCastRenderPose(C.byref(hashData), C.byref(hashParams), C.byref(cam), C.byref(self.m_data), C.byref(self.m_params), T.ctypes.data), "bfRayCastRenderPose")

    def getRayCastData(self) -> BFRayCastData:
        return self.m_data

    def getRayCastParams(self) -> BFRayCastParams:
        return self.m_params

    def updateRayCastMinMax(self, depthMin: float, depthMax: float) -> None:
        self.m_params.m_minDepth, self.m_params.m_maxDepth = depthMin, depthMax

    def download(self) -> dict:
        self._torch.cuda.synchronize(self.device)
        return {"depth": self.d_depth.cpu().numpy(), "depth4": self.d_depth4.cpu().numpy(), "normals": self.d_normals.cpu().numpy(), "colors": self.d_colors.cpu().numpy(),
                "ray_min": self.d_rayMin.cpu().numpy(), "ray_max": self.d_rayMax.cpu().numpy()}
