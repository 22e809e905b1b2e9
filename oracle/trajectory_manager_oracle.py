"""TEST INFRASTRUCTURE ONLY (see oracle/README or DESIGN.md section 2): a plain-Python restatement of the reference's
``TrajectoryManager`` state machine, used by tests/ as the checker of csrc/trajectory_host.cu.  Never imported by the product.

Follows FL/TrajectoryManager.cpp line by line (FL = /root/reference/FriedLiver/Source):
  constructor :7-21, addFrame :23-31, updateOptimizedTransform :33-43, generateUpdateLists :45-108, confirmIntegration :110-114,
  getTopFromReIntegrateList :116-136, getTopFromIntegrateList :138-153, getTopFromDeIntegrateList :155-166,
  getNumActiveOperations :193-199, invalidateFrame :201-210, getOptimizedTransforms FL/TrajectoryManager.h:49-67.
The pose distance uses the C oracle's SE(3) logarithm (oracle/solver_oracle.c orc_matrix_to_pose, itself pinned against the
reference's convertMatricesToPosesCU).  PINNED against the reference's own class: FL/TrajectoryManager.{h,cpp} + the Lie pose maps of
FL/PoseHelper.h compiled by g++ against minimal mLib types (oracle/build_ref.py build_trajectory_host), driven through application-like
sessions; its answers are committed as tests/golden/trajectory_manager_reference.npz and this restatement and the library's C++ replay them
exactly (tests/test_trajectory_manager_reference.py).
"""
import numpy as np

from . import oracle as orc

INTEGRATED, NOT_INTEGRATED_NO_TRANSFORM, NOT_INTEGRATED_WITH_TRANSFORM, INVALID, REINTEGRATION = range(5)
NEG_INF = np.float32(-np.inf)


class Frame:
    def __init__(self):
        self.type = NOT_INTEGRATED_NO_TRANSFORM
        self.frameIdx = 0xFFFFFFFF
        self.integrated = np.full((4, 4), NEG_INF, np.float32)
        self.dist = np.float32(0)


class TrajectoryManagerOracle:
    def __init__(self, numMaxImage, topNActive, minPoseDistSqrt):
        self.optimized = np.zeros((numMaxImage, 4, 4), np.float32)
        self.frames = [Frame() for _ in range(numMaxImage)]
        self.sorted = []
        self.numAdded = 0
        self.numOptimized = 0
        self.toDeIntegrate, self.toIntegrate, self.toReIntegrate = [], [], []
        self.topN, self.minPoseDist, self.rescale = topNActive, np.float32(minPoseDistSqrt), np.float32(2.0)

    def addFrame(self, what, transform, idx):
        f = self.frames[idx]
        f.type, f.frameIdx = what, idx
        f.integrated = np.array(transform, np.float32).reshape(4, 4).copy()
        self.optimized[idx] = f.integrated
        self.sorted.append(f)
        self.numAdded += 1

    def updateOptimizedTransform(self, trajectory, numFrames):
        self.numOptimized = numFrames
        n = min(numFrames, self.numAdded)
        self.optimized[:n] = np.asarray(trajectory, np.float32).reshape(-1, 4, 4)[:n]

    def _invalidate(self, idx):
        f = self.frames[idx]
        if f.type == INVALID:
            return
        before, f.type = f.type, INVALID
        if before == INTEGRATED:
            self.toDeIntegrate.append(f)

    def generateUpdateLists(self):
        n = min(self.numOptimized, self.numAdded)
        for i in range(n):
            f = self.frames[i]
            if self.optimized[i, 0, 0] == NEG_INF:
                self._invalidate(i)
                continue
            if f.type in (NOT_INTEGRATED_NO_TRANSFORM, INVALID):
                f.type = NOT_INTEGRATED_WITH_TRANSFORM
                self.toIntegrate.append(f)
            with np.errstate(all="ignore"):
                ro, to = orc.matrix_to_pose(self.optimized[i])
                ri, ti = orc.matrix_to_pose(f.integrated)
                dt = ti * self.rescale - to * self.rescale      # components 0..2 of the host Pose are the translation (PoseHelper.h:355-358)
                dr = ri - ro
                d = np.float32(0)              # point6d operator| (mLib, un-vendored): six products summed left to right
                for v in (dt[0], dt[1], dt[2], dr[0], dr[1], dr[2]):
                    d = np.float32(d + np.float32(v * v))
                f.dist = d
        head = self.sorted[:n]
        # comparator (cpp:84-93): Integrated first; among Integrated, larger dist first; everything else equivalent -> stable key sort
        # (NaN distances -- frames integrated with an invalid pose -- are ordered last among the integrated ones: the C++ side's documented choice)
        head.sort(key=lambda f: (0, -float(f.dist) if not np.isnan(f.dist) else np.inf) if f.type == INTEGRATED else (1, 0.0))
        self.sorted[:n] = head
        i = len(self.toReIntegrate)
        while i < self.topN and i < n:
            f = self.sorted[i]
            if f.dist > self.minPoseDist and f.type == INTEGRATED:
                f.type = REINTEGRATION
                self.toReIntegrate.append(f)
            else:
                break
            i += 1

    def confirmIntegration(self, idx):
        self.frames[idx].type = INTEGRATED

    def getTopFromReIntegrateList(self):
        if not self.toReIntegrate:
            return None
        while self.toReIntegrate:
            f = self.toReIntegrate.pop(0)
            new = self.optimized[f.frameIdx].copy()
            old = f.integrated.copy()
            idx = f.frameIdx
            if new[0, 0] != NEG_INF:
                f.integrated = new.copy()
                break
        return old, new, idx

    def getTopFromIntegrateList(self):
        if not self.toIntegrate:
            return None
        f = self.toIntegrate.pop(0)
        t = self.optimized[f.frameIdx].copy()
        f.integrated = t.copy()
        return t, f.frameIdx

    def getTopFromDeIntegrateList(self):
        if not self.toDeIntegrate:
            return None
        f = self.toDeIntegrate.pop(0)
        return f.integrated.copy(), f.frameIdx

    def getNumActiveOperations(self):
        return len(self.toDeIntegrate) + len(self.toIntegrate) + len(self.toReIntegrate)

    def getOptimizedTransforms(self):
        n = min(self.numAdded, self.numOptimized)
        out = self.optimized[:n].copy()
        for i in range(n):
            if self.frames[i].type == INVALID:
                out[i] = NEG_INF
        return out
