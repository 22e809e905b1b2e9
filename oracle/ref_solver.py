"""Driver for oracle/_ref/libref_solver*.so -- the REFERENCE's own bundle-adjustment solver kernels (FL/Solver/SolverBundling.cu,
FL/SBA.cu; compat-patched, see oracle/build_ref.py) run on the GPU through their own extern "C" stubs, with buffers allocated and
calls sequenced exactly as the reference host class does (FL/Solver/CUDASolverBundling.cpp:42-86, 187-298).
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from bundlefusion_b200._capi import BFSolverInput, BFSolverParameters, BFSolverState, BFSolverStateAnalysis

_HERE = os.path.dirname(os.path.abspath(__file__))


def _path(fast):
    return os.path.join(_HERE, "_ref", "libref_solver_fast.so" if fast else "libref_solver.so")


def available(fast: bool = True) -> bool:
    return os.path.exists(_path(fast))


class ReferenceSolverBundling:
    def __init__(self, maxNumberOfImages: int, maxNumResiduals: int, device="cuda:0", fast_math: bool = True):
        import torch
        self._torch = torch
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        L = self.L = C.CDLL(_path(fast_math))
        P = C.POINTER
        L.solveBundlingStub.argtypes = [P(BFSolverInput), P(BFSolverState), P(BFSolverParameters), P(BFSolverStateAnalysis), C.c_void_p, C.c_void_p]
        L.solveBundlingStub.restype = None
        L.buildVariablesToCorrespondencesTableCUDA.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        L.buildVariablesToCorrespondencesTableCUDA.restype = None
        L.evalMaxResidual.argtypes = [P(BFSolverInput), P(BFSolverState), P(BFSolverStateAnalysis), P(BFSolverParameters), C.c_void_p]
        L.evalMaxResidual.restype = None
        L.countHighResiduals.argtypes = [P(BFSolverInput), P(BFSolverState), P(BFSolverParameters), C.c_void_p]
        L.countHighResiduals.restype = C.c_int
        L.convertLiePosesToMatricesCU.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
        L.convertLiePosesToMatricesCU.restype = None
        L.convertMatricesToPosesCU.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        L.convertMatricesToPosesCU.restype = None
        L.convertPosesToMatricesCU.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
        L.convertPosesToMatricesCU.restype = None
        N, R = maxNumberOfImages, maxNumResiduals
        self.N, self.R = N, R
        self.m_maxCorrPerImage = int(min(max(R // N, 1000), 4000))                      # cpp:39
        f = lambda n: torch.zeros(max(int(n), 1), dtype=torch.float32, device=self.device)
        i32 = lambda n: torch.zeros(max(int(n), 1), dtype=torch.int32, device=self.device)
        b = self._bufs = {}
        for name in ("d_deltaRot", "d_deltaTrans", "d_rRot", "d_rTrans", "d_zRot", "d_zTrans", "d_pRot", "d_pTrans", "d_Ap_XRot", "d_Ap_XTrans",
                     "d_precondionerRot", "d_precondionerTrans"):
            b[name] = f(3 * N)
        b["d_Jp"] = f(3 * R)
        b["d_scanAlpha"] = f(2)
        b["d_rDotzOld"] = f(N)
        b["d_sumResidual"] = f(1)
        b["d_countHighResidual"] = i32(1)
        b["d_denseJtJ"] = f(36 * N * N)
        b["d_denseJtr"] = f(6 * N)
        npairs = N * (N - 1) // 2
        b["d_denseCorrCounts"] = f(npairs)
        b["d_xTransforms"] = f(16 * N)
        b["d_xTransformInverses"] = f(16 * N)
        b["d_denseOverlappingImages"] = i32(2 * npairs)
        b["d_numDenseOverlappingImages"] = i32(1)
        b["d_corrCount"] = i32(1)
        b["d_corrCountColor"] = i32(1)
        b["d_sumResidualColor"] = f(1)
        st = BFSolverState()
        for k, v in b.items():
            setattr(st, k, v.data_ptr())
        self.st = st
        self.d_variablesToCorrespondences = i32(N * self.m_maxCorrPerImage)
        self.d_numEntriesPerRow = i32(N)
        nblk = (R + 511) // 512
        self.d_maxResidual, self.d_maxResidualIndex = f(nblk), i32(nblk)
        self.h_maxResidual, self.h_maxResidualIndex = np.zeros(nblk, np.float32), np.zeros(nblk, np.int32)
        an = BFSolverStateAnalysis()
        an.d_maxResidual, an.d_maxResidualIndex = self.d_maxResidual.data_ptr(), self.d_maxResidualIndex.data_ptr()
        an.h_maxResidual = self.h_maxResidual.ctypes.data
        an.h_maxResidualIndex = self.h_maxResidualIndex.ctypes.data
        self.an = an

    def solve(self, d_corr, nCorr, d_valid, nImages, nNonLin, nLin, wS, wD=None, wC=None, d_rot=None, d_trans=None, cudaCache=None,
              usePairwiseDense=True, record_convergence=False):
        """CUDASolverBundling::solve (cpp:187-284) on the default stream (the reference uses no other); synchronous in effect."""
        t = self._torch
        nNonLin = min(nNonLin, len(wS))
        wD = wD if wD is not None else [0.0] * len(wS)
        wC = wC if wC is not None else [0.0] * len(wS)
        arrs = [np.ascontiguousarray(w, np.float32) for w in (wS, wD, wC)]
        self.st.d_xRot, self.st.d_xTrans = d_rot.data_ptr(), d_trans.data_ptr()
        p = BFSolverParameters()
        p.nNonLinearIterations, p.nLinIterations = nNonLin, nLin
        p.verifyOptDistThresh, p.verifyOptPercentThresh, p.highResidualThresh = 0.02, 0.05, float("inf")
        p.denseDistThresh, p.denseNormalThresh, p.denseColorThresh, p.denseColorGradientMin = 0.15, 0.97, 0.1, 0.005
        p.denseDepthMin, p.denseDepthMax, p.denseOverlapCheckSubsampleFactor = 0.5, 4.0, 4
        p.weightSparse, p.weightDenseDepth, p.weightDenseColor = float(arrs[0][0]), float(arrs[1][0]), float(arrs[2][0])
        p.useDense = 1 if (p.weightDenseDepth > 0 or p.weightDenseColor > 0) else 0
        p.useDenseDepthAllPairwise = 1 if usePairwiseDense else 0
        si = BFSolverInput()
        si.d_correspondences = d_corr.data_ptr()
        si.d_variablesToCorrespondences = self.d_variablesToCorrespondences.data_ptr()
        si.d_numEntriesPerRow = self.d_numEntriesPerRow.data_ptr()
        si.numberOfCorrespondences, si.numberOfImages = nCorr, nImages
        si.maxNumberOfImages, si.maxCorrPerImage = self.N, self.m_maxCorrPerImage
        si.maxNumDenseImPairs = self.N * (self.N - 1) // 2
        si.d_validImages = d_valid.data_ptr()
        fp = C.POINTER(C.c_float)
        si.weightsSparse, si.weightsDenseDepth, si.weightsDenseColor = (a.ctypes.data_as(fp) for a in arrs)
        if cudaCache is not None:
            si.d_cacheFrames = cudaCache.getCacheFramesGPU().data_ptr()
            si.denseDepthWidth, si.denseDepthHeight = cudaCache.width, cudaCache.height
            for k in range(4):
                si.intrinsics[k] = cudaCache.intrinsics[k]
        else:
            si.d_cacheFrames = None
            for k in range(4):
                si.intrinsics[k] = float("-inf")
        t.cuda.synchronize()
        self.d_numEntriesPerRow.zero_()                                                  # cpp:288
        t.cuda.synchronize()
        if nCorr > 0:
            self.L.buildVariablesToCorrespondencesTableCUDA(d_corr.data_ptr(), nCorr, self.m_maxCorrPerImage,
                                                            self.d_variablesToCorrespondences.data_ptr(), self.d_numEntriesPerRow.data_ptr(), None)
        conv = np.full(nNonLin + 1, -1.0, np.float32) if record_convergence else None
        self.L.solveBundlingStub(C.byref(si), C.byref(self.st), C.byref(p), C.byref(self.an), conv.ctypes.data if conv is not None else None, None)
        t.cuda.synchronize()
        self._last = (si, p, arrs)
        return conv

    def max_residual(self):
        """computeMaxResidual (cpp:313-329): device block maxima + the host reduction."""
        si, p, _ = self._last
        self.L.evalMaxResidual(C.byref(si), C.byref(self.st), C.byref(self.an), C.byref(p), None)
        self._torch.cuda.synchronize()
        n = (si.numberOfCorrespondences + 511) // 512
        r = self.d_maxResidual.cpu().numpy()[:n]
        idx = self.d_maxResidualIndex.cpu().numpy()[:n]
        k = int(np.argmax(r))
        return float(r[k]), int(idx[k])
