"""Host logic of the SBA mirror (bundlefusion_b200/sba.py vs FL/SBA.cpp): weight schedules and the sequence of one align() call, with the
device steps and the solver replaced by recorders.  Runs without a GPU."""
from bundlefusion_b200.sba import SBA


class FakeSolver:
    def __init__(self, maxRes=(0.2, 5), thresh=0.08, verify=True):
        self.calls, self.maxRes, self.thresh, self.verify = [], maxRes, thresh, verify

    def solve(self, corr, n, valid, numImages, nNonLin, nLin, wS, wD, wC, **kw):
        self.calls.append(("solve", n, numImages, nNonLin, nLin, list(wS), list(wD), list(wC), kw["rebuildJT"], kw["findMaxResidual"], kw["cudaCache"], kw["usePairwiseDense"]))

    def getMaxResidual(self):
        return self.maxRes

    def shouldRemove(self, i, j):
        return (not (i == 0 and j < 10)) and self.maxRes[0] > self.thresh

    def useVerification(self, corr, n):
        self.calls.append(("verify", n)); return self.verify

    def getVarToCorrNumEntriesPerRow(self):
        return "rows"


class FakeOps:
    def __init__(self, pair=(3, 7)):
        self.calls, self.pair = [], pair

    def matrices_to_poses(self, *a): self.calls.append("m2p")
    def poses_to_matrices(self, *a): self.calls.append("p2m")
    def entry_images(self, corr, index): self.calls.append(("entry", index)); return self.pair
    def invalidate_pair(self, corr, n, i, j): self.calls.append(("invalidate", n, i, j))
    def check_invalid_frames(self, rows, valid, n, corr, num, comp): self.calls.append(("check", rows, n, num, comp))


def test_weight_schedules():
    s = SBA(FakeSolver(), 2, 3, ops=FakeOps())
    assert s.m_localWeightsSparse == [1, 1, 1] and s.m_localWeightsDenseDepth == [1, 2, 3] and s.m_localWeightsDenseColor == [0, 0, 0]        # SBA.cpp:28-32
    assert s.m_globalWeightsSparse == [1, 1, 1] and s.m_globalWeightsDenseDepth == [1, 1, 2] and s.m_globalWeightsDenseColor == [0.1] * 3      # :35-38
    assert s.weights(True, "cache") == ("cache", [1, 1, 1], [1, 2, 3], [0, 0, 0])
    assert s.weights(False, "cache") == (None, [1, 1, 1], [0, 0, 0], [0, 0, 0])                 # global: sparse only until the end-of-scan optimisation
    s.m_bUseGlobalDenseOpt = True
    assert s.weights(False, "cache") == ("cache", [1, 1, 1], [1, 1, 2], [0.1] * 3)
    s.m_bUseLocalDense = False
    assert s.weights(True, "cache") == (None, [1, 1, 1], [0, 0, 0], [0, 0, 0])
    big = SBA(FakeSolver(), 6, 4, ops=FakeOps())
    assert big.m_globalWeightsDenseDepth == [1, 1, 2, 3, 4, 5] and len(big.m_localWeightsDenseDepth) == 6


def test_align_sequence_with_removal():
    solver, ops = FakeSolver(maxRes=(0.2, 5)), FakeOps(pair=(3, 7))
    s = SBA(solver, ops=ops, useComprehensiveFrameInvalidation=True)
    removed = s.align("corr", 100, "valid", 11, "T", "rot", "trans", 2, 100, True, True, True, True, curFrame=10, cudaCache="cache")
    assert removed and s.m_maxResidual == 0.2 and s.m_bVerify is True and s.removed_pairs == [(3, 7, 0.2)]
    assert ops.calls == ["m2p", ("entry", 5), ("invalidate", 100, 3, 7), ("check", "rows", 11, 100, True), "p2m"]
    assert solver.calls[0] == ("solve", 100, 11, 2, 100, [1, 1, 1], [1, 2, 3], [0, 0, 0], True, True, "cache", True) and solver.calls[1] == ("verify", 100)


def test_align_without_removal():
    # residual below the threshold
    solver, ops = FakeSolver(maxRes=(0.05, 5)), FakeOps()
    assert SBA(solver, ops=ops).align("c", 10, "v", 4, "T", "r", "t", 3, 150, False, False, False, True, curFrame=3) is False
    assert ops.calls == ["m2p", ("entry", 5), "p2m"] and len(solver.calls) == 1 and solver.calls[0][8:11] == (False, True, None)
    # the pair (0, j < 10) is never removed (CUDASolverBundling.cpp:445)
    solver, ops = FakeSolver(maxRes=(0.5, 2)), FakeOps(pair=(0, 4))
    assert SBA(solver, ops=ops).align("c", 10, "v", 4, "T", "r", "t", 3, 150, False, False, True, True, curFrame=3) is False
    assert ("invalidate", 10, 0, 4) not in ops.calls
    # not the last solve of the frame: no max-residual search at all
    solver, ops = FakeSolver(), FakeOps()
    assert SBA(solver, ops=ops).align("c", 10, "v", 4, "T", "r", "t", 3, 150, False, False, True, False, curFrame=3) is False
    assert ops.calls == ["m2p", "p2m"] and solver.calls[0][9] is False
    # no valid maximum (no correspondences)
    solver, ops = FakeSolver(maxRes=(0.0, -1)), FakeOps()
    assert SBA(solver, ops=ops).align("c", 0, "v", 4, "T", "r", "t", 3, 150, False, False, True, True, curFrame=3) is False
    assert ops.calls == ["m2p", "p2m"]
