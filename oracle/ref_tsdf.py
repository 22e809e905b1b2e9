"""Driver for oracle/_ref/libref_tsdf*.so -- the REFERENCE's own TSDF kernels (compat-patched, see oracle/build_ref.py) run on the
GPU through their own extern "C" stubs, sequenced exactly as the reference host class does
(FL/DepthSensing/CUDASceneRepHashSDF.h:65-155, 328-391).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from bundlefusion_b200._capi import (BF_HASH_BUCKET_SIZE, BF_SDF_BLOCK_VOXELS, BFDepthCameraData, BFDepthCameraParams, BFHashDataStruct,
                                     BFHashParams)
from bundlefusion_b200.scene_rep import set_pose

_HERE = os.path.dirname(os.path.abspath(__file__))


def available(fast: bool = True) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_tsdf_fast.so" if fast else "libref_tsdf.so"))


class ReferenceSceneRepHashSDF:
    def __init__(self, params: BFHashParams, device="cuda:0", fast_math: bool = True):
        import torch
        self._torch = torch
        self.device = torch.device(device)
        self.L = C.CDLL(os.path.join(_HERE, "_ref", "libref_tsdf_fast.so" if fast_math else "libref_tsdf.so"))
        P = C.POINTER
        L = self.L
        L.updateConstantHashParams.argtypes = [P(BFHashParams)]
        L.updateConstantDepthCameraParams.argtypes = [P(BFDepthCameraParams)]
        L.bindInputDepthColorTextures.argtypes = [P(BFDepthCameraData), C.c_uint, C.c_uint]
        for n in ("resetCUDA", "resetHashBucketMutexCUDA", "garbageCollectIdentifyCUDA", "garbageCollectFreeCUDA"):
            getattr(L, n).argtypes = [P(BFHashDataStruct), P(BFHashParams)]; getattr(L, n).restype = None
        L.allocCUDA.argtypes = [P(BFHashDataStruct), P(BFHashParams), P(BFDepthCameraData), P(BFDepthCameraParams), C.c_void_p]
        L.compactifyHashAllInOneCUDA.argtypes = [P(BFHashDataStruct), P(BFHashParams)]; L.compactifyHashAllInOneCUDA.restype = C.c_uint
        for n in ("integrateDepthMapCUDA", "deIntegrateDepthMapCUDA"):
            getattr(L, n).argtypes = [P(BFHashDataStruct), P(BFHashParams), P(BFDepthCameraData), P(BFDepthCameraParams)]; getattr(L, n).restype = None
        self.hp = BFHashParams()
        C.memmove(C.byref(self.hp), C.byref(params), C.sizeof(BFHashParams))
        hp = self.hp
        n_entries, n_blocks = hp.m_hashNumBuckets * BF_HASH_BUCKET_SIZE, hp.m_numSDFBlocks
        kw = dict(device=self.device)
        torch.cuda.set_device(self.device)
        self.d_heap = torch.empty(n_blocks, dtype=torch.int32, **kw)
        self.d_heapCounter = torch.zeros(1, dtype=torch.int32, **kw)
        self.d_hash = torch.empty(n_entries * 8, dtype=torch.int32, **kw)
        self.d_hashDecision = torch.zeros(n_entries, dtype=torch.int32, **kw)
        self.d_hashDecisionPrefix = torch.zeros(n_entries, dtype=torch.int32, **kw)
        self.d_hashCompactified = torch.empty(n_entries * 8, dtype=torch.int32, **kw)
        self.d_hashCompactifiedCounter = torch.zeros(1, dtype=torch.int32, **kw)
        self.d_SDFBlocks = torch.empty(n_blocks * BF_SDF_BLOCK_VOXELS * 3, dtype=torch.int32, **kw)
        self.d_hashBucketMutex = torch.empty(hp.m_hashNumBuckets, dtype=torch.int32, **kw)
        hd = BFHashDataStruct()
        for name in ("d_heap", "d_heapCounter", "d_hashDecision", "d_hashDecisionPrefix", "d_hash", "d_hashCompactified", "d_hashCompactifiedCounter",
                     "d_SDFBlocks", "d_hashBucketMutex"):
            setattr(hd, name, getattr(self, name).data_ptr())
        hd.m_bIsOnGPU = 1
        self.hd = hd
        self.alloc_rounds = 0
        self.reset()

    def reset(self):                                           # h:147-155
        set_pose(self.hp, np.eye(4, dtype=np.float32))
        self.hp.m_numOccupiedBlocks = 0
        self.L.updateConstantHashParams(C.byref(self.hp))
        self.L.resetCUDA(C.byref(self.hd), C.byref(self.hp))
        self._torch.cuda.synchronize()

    def getHeapFreeCount(self) -> int:                         # h:168-172
        return (int(self.d_heapCounter.cpu().numpy().view(np.uint32)[0]) + 1) & 0xFFFFFFFF

    def _begin(self, T, depth, color, cam):
        dd = BFDepthCameraData(); dd.d_depthData = depth.data_ptr(); dd.d_colorData = color.data_ptr() if color is not None else None
        self.L.updateConstantDepthCameraParams(C.byref(cam))
        self.L.bindInputDepthColorTextures(C.byref(dd), cam.m_imageWidth, cam.m_imageHeight)     # h:61-63
        set_pose(self.hp, T)                                                                     # h:128-134
        self.L.updateConstantHashParams(C.byref(self.hp))
        return dd

    def _compactify(self):                                     # h:355-391
        self.hp.m_numOccupiedBlocks = self.L.compactifyHashAllInOneCUDA(C.byref(self.hd), C.byref(self.hp))
        self.L.updateConstantHashParams(C.byref(self.hp))

    def integrate(self, T, depth, color, cam):                 # h:65-83
        dd = self._begin(T, depth, color, cam)
        prev = self.getHeapFreeCount()                         # h:328-352 (alloc loop)
        while True:
            self.L.resetHashBucketMutexCUDA(C.byref(self.hd), C.byref(self.hp))
            self.L.allocCUDA(C.byref(self.hd), C.byref(self.hp), C.byref(dd), C.byref(cam), None)
            self.alloc_rounds += 1
            cur = self.getHeapFreeCount()
            if cur == prev:
                break
            prev = cur
        self._compactify()
        self.L.integrateDepthMapCUDA(C.byref(self.hd), C.byref(self.hp), C.byref(dd), C.byref(cam))

    def deIntegrate(self, T, depth, color, cam):               # h:85-108
        dd = self._begin(T, depth, color, cam)
        self._compactify()
        self.L.deIntegrateDepthMapCUDA(C.byref(self.hd), C.byref(self.hp), C.byref(dd), C.byref(cam))

    def garbageCollect(self):                                  # h:110-126
        if self.hp.m_numOccupiedBlocks > 0:
            self.L.garbageCollectIdentifyCUDA(C.byref(self.hd), C.byref(self.hp))
            self.L.resetHashBucketMutexCUDA(C.byref(self.hd), C.byref(self.hp))
            self.L.garbageCollectFreeCUDA(C.byref(self.hd), C.byref(self.hp))

    def download(self) -> dict:
        self._torch.cuda.synchronize(self.device)
        n_entries = self.hp.m_hashNumBuckets * BF_HASH_BUCKET_SIZE
        return {"hash": self.d_hash.cpu().numpy().reshape(n_entries, 8), "compactified": self.d_hashCompactified.cpu().numpy().reshape(n_entries, 8),
                "compactified_count": int(self.hp.m_numOccupiedBlocks), "heap": self.d_heap.cpu().numpy().view(np.uint32),
                "heap_counter": int(self.d_heapCounter.cpu().numpy().view(np.uint32)[0]), "voxels": self.d_SDFBlocks.cpu().numpy().reshape(-1, BF_SDF_BLOCK_VOXELS, 3),
                "decision": self.d_hashDecision.cpu().numpy(), "mutex": self.d_hashBucketMutex.cpu().numpy()}
