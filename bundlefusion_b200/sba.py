"""Host-side mirror of the reference's ``SBA`` (FL/SBA.{h,cpp}): the weight schedules of the local and the global bundle adjustment and
the sequence ``align`` runs around one solve -- matrices -> Lie poses, solve, optional verification, removal of the worst image pair
(``removeMaxResidualCUDA``: invalidate its correspondences, drop frames that lost all of theirs), Lie poses -> matrices.  All arithmetic is
in the library; this class only sequences C-ABI calls, through an ``ops`` object so that the control flow can be tested without a GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi


class DeviceOps:
    """The device-side steps ``SBA.align`` needs, on torch CUDA tensors."""

    def __init__(self):
        self.lib = capi.lib()
        L, vp, u = self.lib, C.c_void_p, C.c_uint
        L.convertMatricesToPosesCU.argtypes = [vp, u, vp, vp, vp]; L.convertMatricesToPosesCU.restype = None      # FL/SBA.cpp:12-15
        L.convertPosesToMatricesCU.argtypes = [vp, vp, u, vp, vp]; L.convertPosesToMatricesCU.restype = None

    def matrices_to_poses(self, d_transforms, n, d_rot, d_trans, d_valid):
        self.lib.convertMatricesToPosesCU(d_transforms.data_ptr(), n, d_rot.data_ptr(), d_trans.data_ptr(), d_valid.data_ptr())

    def poses_to_matrices(self, d_rot, d_trans, n, d_transforms, d_valid):
        self.lib.convertPosesToMatricesCU(d_rot.data_ptr(), d_trans.data_ptr(), n, d_transforms.data_ptr(), d_valid.data_ptr())

    def entry_images(self, d_correspondences, index):
        """(imgIdx_i, imgIdx_j) of one EntryJ: the 8-byte read the reference does in CUDASolverBundling::getMaxResidual (cpp:436-440)."""
        e = d_correspondences.view(-1)[32 * index:32 * index + 8].cpu().numpy().view(np.uint32)
        return int(e[0]), int(e[1])

    def invalidate_pair(self, d_correspondences, num, i, j):
        capi.check(self.lib.bfSiftInvalidateImageToImage(d_correspondences.data_ptr(), num, i, j), "bfSiftInvalidateImageToImage")

    def check_invalid_frames(self, d_numEntriesPerRow, d_valid, numImages, d_correspondences, num, comprehensive):
        capi.check(self.lib.bfSiftCheckForInvalidFrames(d_numEntriesPerRow.data_ptr(), d_valid.data_ptr(), numImages, d_correspondences.data_ptr(), num,
                                                        1 if comprehensive else 0), "bfSiftCheckForInvalidFrames")


class SBA:
    def __init__(self, solver, numLocalNonLinIterations: int = 2, numGlobalNonLinIterations: int = 3, useLocalDense: bool = True,
                 useGlobalDenseOpt: bool = False, useComprehensiveFrameInvalidation: bool = False, ops=None):
        """Weight schedules: FL/SBA.cpp:26-47.  ``solver``: a CUDASolverBundling mirror (solve / getMaxResidual / shouldRemove / useVerification /
        getVarToCorrNumEntriesPerRow)."""
        n = max(numGlobalNonLinIterations, numLocalNonLinIterations)
        self.m_solver = solver
        self.ops = ops if ops is not None else DeviceOps()
        self.m_localWeightsSparse = [1.0] * n
        self.m_localWeightsDenseDepth = [i + 1.0 for i in range(n)]
        self.m_localWeightsDenseColor = [0.0] * n
        self.m_globalWeightsSparse = [1.0] * n
        self.m_globalWeightsDenseDepth = [1.0] * n
        for i in range(2, n):
            self.m_globalWeightsDenseDepth[i] = float(i)
        self.m_globalWeightsDenseColor = [0.1] * n
        self.m_bUseLocalDense, self.m_bUseGlobalDenseOpt = useLocalDense, useGlobalDenseOpt
        self.m_bUseComprehensiveFrameInvalidation = useComprehensiveFrameInvalidation
        self.m_maxResidual, self.m_bVerify = -1.0, False
        self.removed_pairs = []

    def weights(self, isLocal: bool, cache):
        """(cache to use, sparse, dense depth, dense colour) -- FL/SBA.cpp:64-95."""
        if isLocal:
            if self.m_bUseLocalDense:
                return cache, self.m_localWeightsSparse, self.m_localWeightsDenseDepth, self.m_localWeightsDenseColor
            z = [0.0] * len(self.m_localWeightsDenseDepth)
            return None, self.m_localWeightsSparse, z, z
        if not self.m_bUseGlobalDenseOpt:
            z = [0.0] * len(self.m_globalWeightsDenseDepth)
            return None, self.m_globalWeightsSparse, z, z
        return cache, self.m_globalWeightsSparse, self.m_globalWeightsDenseDepth, self.m_globalWeightsDenseColor

    def align(self, d_correspondences, numCorrespondences: int, d_validImages, numImages: int, d_transforms, d_xRot, d_xTrans, maxNumIters: int,
              numPCGits: int, useVerify: bool, isLocal: bool, isStart: bool, isEnd: bool, curFrame: int, cudaCache=None, revalidateIdx: int = -1) -> bool:
        """SBA::align + alignCUDA (FL/SBA.cpp:53-132).  Returns whether an image pair was removed (the caller then solves again)."""
        self.m_bVerify, self.m_maxResidual = False, -1.0
        cache, wS, wD, wC = self.weights(isLocal, cudaCache)
        self.ops.matrices_to_poses(d_transforms, numImages, d_xRot, d_xTrans, d_validImages)
        self.m_solver.solve(d_correspondences, numCorrespondences, d_validImages, numImages, maxNumIters, numPCGits, wS, wD, wC,
                            d_rotationAnglesUnknowns=d_xRot, d_translationUnknowns=d_xTrans, rebuildJT=isStart, findMaxResidual=isEnd, cudaCache=cache,
                            usePairwiseDense=True)
        removed = False
        if isEnd and wS[0] > 0:
            frame = curFrame if revalidateIdx == -1 else revalidateIdx
            removed = self.removeMaxResidual(d_correspondences, numCorrespondences, d_validImages, numImages, frame)
        if useVerify:
            self.m_bVerify = self.m_solver.useVerification(d_correspondences, numCorrespondences) if wS[0] > 0 else True
        self.ops.poses_to_matrices(d_xRot, d_xTrans, numImages, d_transforms, d_validImages)
        return removed

    def removeMaxResidual(self, d_correspondences, numCorrespondences, d_validImages, numImages, curFrame) -> bool:
        """SBA::removeMaxResidualCUDA (FL/SBA.cpp:165-203) with CUDASolverBundling::getMaxResidual's decision (cpp:429-452)."""
        maxRes, index = self.m_solver.getMaxResidual()
        self.m_maxResidual = maxRes
        if index < 0:
            return False
        i, j = self.ops.entry_images(d_correspondences, index)
        if not self.m_solver.shouldRemove(i, j):
            return False
        self.ops.invalidate_pair(d_correspondences, numCorrespondences, i, j)
        self.removed_pairs.append((i, j, maxRes))
        self.ops.check_invalid_frames(self.m_solver.getVarToCorrNumEntriesPerRow(), d_validImages, numImages, d_correspondences, numCorrespondences,
                                      self.m_bUseComprehensiveFrameInvalidation)
        return True
