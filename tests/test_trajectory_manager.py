"""Host-side TrajectoryManager (csrc/trajectory_host.cu behind include/bf_bundler.h) against the plain-Python restatement of
FL/TrajectoryManager.cpp (oracle/trajectory_manager_oracle.py).  Pure host logic: runs without a GPU."""
import numpy as np
import pytest

from bundlefusion_b200 import trajectory_manager as tmod
from bundlefusion_b200.trajectory_manager import TrajectoryManager
from oracle import oracle as orc
from oracle.trajectory_manager_oracle import TrajectoryManagerOracle

NEG = np.full((4, 4), -np.inf, np.float32)


def rand_pose(rng, rot=0.3, trans=0.5):
    return orc.pose_to_matrix(rng.normal(0, rot, 3).astype(np.float32), rng.normal(0, trans, 3).astype(np.float32))


def same_state(tm, ref, n):
    assert tm.getNumAddedFrames() == ref.numAdded and tm.getNumOptimizedFrames() == ref.numOptimized
    assert tm.getNumActiveOperations() == ref.getNumActiveOperations()
    for i in range(n):
        assert tm.frameType(i) == ref.frames[i].type, i
    a, b = tm.getOptimizedTransforms(), ref.getOptimizedTransforms()
    assert a.shape == b.shape and np.array_equal(a, b)


def same_pop(a, b):
    if b is None:
        assert a is None
        return
    assert a is not None and len(a) == len(b)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y))


def test_constructor_and_empty_lists():
    tm = TrajectoryManager(16, topNActive=3, minPoseDistSqrt=0.01)
    assert tm.getNumAddedFrames() == 0 and tm.getNumOptimizedFrames() == 0 and tm.getNumActiveOperations() == 0
    assert tm.getTopFromIntegrateList() is None and tm.getTopFromDeIntegrateList() is None and tm.getTopFromReIntegrateList() is None
    assert all(tm.frameType(i) == tmod.NOT_INTEGRATED_NO_TRANSFORM for i in range(16))
    tm.generateUpdateLists()                                  # nothing added: no-op
    assert tm.getNumActiveOperations() == 0 and tm.getOptimizedTransforms().shape == (0, 4, 4)
    with pytest.raises(IndexError):
        tm.addFrame(tmod.INTEGRATED, np.eye(4), 16)


def test_reintegration_selection_by_pose_distance():
    """Five integrated frames; the optimiser moves #1 a lot, #3 a little, #4 below the threshold: the list is (#1, #3), capped by topN."""
    rng = np.random.default_rng(1)
    poses = [rand_pose(rng) for _ in range(5)]
    tm = TrajectoryManager(8, topNActive=2, minPoseDistSqrt=1e-3)
    for i, T in enumerate(poses):
        tm.addFrame(tmod.INTEGRATED, T, i)
    opt = np.stack(poses).copy()
    opt[1, :3, 3] += np.float32(0.5)
    opt[3, :3, 3] += np.float32(0.1)
    opt[4, 0, 3] += np.float32(0.01)                          # dist 4e-4 < 1e-3
    opt[2, 1, 3] += np.float32(0.05)                          # dist 1e-2: third largest, cut by topN = 2
    tm.updateOptimizedTransform(opt, 5)
    tm.generateUpdateLists()
    assert [tm.frameType(i) for i in range(5)] == [tmod.INTEGRATED, tmod.REINTEGRATION, tmod.INTEGRATED, tmod.REINTEGRATION, tmod.INTEGRATED]
    # the distance is taken between Lie-algebra poses (t_lie = V^-1 t), so it is |dt|^2 only up to the rotation's V^-1: a few percent here
    for i, want in ((1, 4 * 0.75), (3, 4 * 0.03), (4, 4 * 1e-4), (2, 4 * 2.5e-3)):          # the factor 2 multiplies the translation part of the pose
        assert abs(tm.frameDist(i) - want) < 0.15 * want, (i, tm.frameDist(i))
    assert tm.frameDist(0) == 0.0
    assert tm.getNumActiveOperations() == 2
    old, new, idx = tm.getTopFromReIntegrateList()
    assert idx == 1 and np.array_equal(old, poses[1]) and np.array_equal(new, opt[1])
    tm.confirmIntegration(1)
    tm.generateUpdateLists()                                  # #1 is now integrated with its optimised pose (dist 0)
    # Reference quirk kept (cpp:97): the refill loop starts at sorted[len(list)], not sorted[0] -- with #3 still queued it looks at the
    # SECOND-largest integrated frame (#4, below the threshold) and stops, so #2 is not picked yet.
    assert tm.frameDist(1) == 0.0 and tm.frameType(2) == tmod.INTEGRATED and tm.getNumActiveOperations() == 1
    assert tm.getTopFromReIntegrateList()[2] == 3
    tm.confirmIntegration(3)
    tm.generateUpdateLists()                                  # list empty -> starts at sorted[0] = #2
    assert tm.frameType(2) == tmod.REINTEGRATION and tm.getTopFromReIntegrateList()[2] == 2
    assert tm.getTopFromReIntegrateList() is None


def test_invalidation_and_revalidation():
    rng = np.random.default_rng(2)
    poses = np.stack([rand_pose(rng) for _ in range(4)])
    tm = TrajectoryManager(4, topNActive=4, minPoseDistSqrt=0.0)
    tm.addFrame(tmod.INTEGRATED, poses[0], 0)
    tm.addFrame(tmod.INTEGRATED, poses[1], 1)
    tm.addFrame(tmod.NOT_INTEGRATED_NO_TRANSFORM, NEG, 2)     # a frame that could not be aligned (Bundler marks it -inf)
    tm.addFrame(tmod.INTEGRATED, poses[3], 3)
    opt = poses.copy(); opt[1] = -np.inf; opt[2] = -np.inf
    tm.updateOptimizedTransform(opt, 4)
    tm.generateUpdateLists()
    assert tm.frameType(1) == tmod.INVALID and tm.frameType(2) == tmod.INVALID
    T, idx = tm.getTopFromDeIntegrateList()                   # only the frame that WAS integrated is de-integrated, with the pose it went in with
    assert idx == 1 and np.array_equal(T, poses[1]) and tm.getTopFromDeIntegrateList() is None
    out = tm.getOptimizedTransforms()
    assert np.all(np.isneginf(out[1])) and np.all(np.isneginf(out[2])) and np.array_equal(out[0], poses[0])
    opt[2] = poses[2]                                         # the optimiser recovers frame 2
    tm.updateOptimizedTransform(opt, 4)
    tm.generateUpdateLists()
    assert tm.frameType(2) == tmod.NOT_INTEGRATED_WITH_TRANSFORM
    T, idx = tm.getTopFromIntegrateList()
    assert idx == 2 and np.array_equal(T, poses[2])
    tm.confirmIntegration(2)
    assert tm.frameType(2) == tmod.INTEGRATED and tm.getNumActiveOperations() == 0


def test_reintegrate_list_skips_frames_invalidated_while_queued():
    """getTopFromReIntegrateList pops entries whose optimised pose went invalid meanwhile and returns the first valid one (cpp:121-133)."""
    rng = np.random.default_rng(3)
    poses = np.stack([rand_pose(rng) for _ in range(3)])
    tm = TrajectoryManager(3, topNActive=3, minPoseDistSqrt=0.0)
    ref = TrajectoryManagerOracle(3, 3, 0.0)
    for m in (tm, ref):
        for i in range(3):
            m.addFrame(tmod.INTEGRATED, poses[i], i)
    opt = poses.copy(); opt[:, 0, 3] += np.float32([0.3, 0.2, 0.1])
    for m in (tm, ref):
        m.updateOptimizedTransform(opt, 3); m.generateUpdateLists()
    opt2 = opt.copy(); opt2[0] = -np.inf
    for m in (tm, ref):
        m.updateOptimizedTransform(opt2, 3)
    a, b = tm.getTopFromReIntegrateList(), ref.getTopFromReIntegrateList()
    same_pop(a, b)
    assert a[2] == 1 and tm.getNumActiveOperations() == 1
    same_state(tm, ref, 3)


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4])
def test_random_sessions_match_the_restatement(seed):
    """A scanning session in miniature: frames arrive, the optimiser rewrites the trajectory (sometimes marking frames invalid, sometimes
    for fewer / more frames than were added), the fusion thread pops one operation of each kind per frame and confirms it."""
    rng = np.random.default_rng(100 + seed)
    N, topN, minD = 40, int(rng.integers(1, 6)), float(rng.choice([0.0, 1e-4, 1e-2]))
    tm, ref = TrajectoryManager(N, topN, minD), TrajectoryManagerOracle(N, topN, minD)
    traj = np.zeros((N, 4, 4), np.float32)
    added, nOpt = 0, 0
    for step in range(3 * N):
        op = rng.random()
        if added < N and (op < 0.35 or added == 0):
            valid = rng.random() < 0.85
            T = rand_pose(rng) if valid else NEG.copy()
            kind = tmod.INTEGRATED if valid else tmod.NOT_INTEGRATED_NO_TRANSFORM
            tm.addFrame(kind, T, added); ref.addFrame(kind, T, added)
            traj[added] = T
            added += 1
        elif op < 0.6:
            # the optimised-frame count only grows in the application (with a shrinking count the reference sorts on stale distances)
            n = nOpt = int(min(N, max(nOpt, added + rng.integers(-2, 2))))
            for i in range(min(n, added)):
                r = rng.random()
                if r < 0.1:
                    traj[i] = -np.inf
                elif r < 0.5 or np.isneginf(traj[i, 0, 0]):
                    base = traj[i] if not np.isneginf(traj[i, 0, 0]) else rand_pose(rng)
                    traj[i] = (rand_pose(rng, 0.02, 0.05) @ base).astype(np.float32)
                    traj[i, 3] = [0, 0, 0, 1]
            tm.updateOptimizedTransform(traj, n); ref.updateOptimizedTransform(traj, n)
            tm.generateUpdateLists(); ref.generateUpdateLists()
            for i in range(min(n, added)):
                if ref.frames[i].type in (tmod.INTEGRATED, tmod.REINTEGRATION) and np.isfinite(ref.frames[i].dist):
                    assert tm.frameDist(i) == ref.frames[i].dist, (i, tm.frameDist(i), ref.frames[i].dist)
        else:
            a, b = tm.getTopFromDeIntegrateList(), ref.getTopFromDeIntegrateList()
            same_pop(a, b)
            a, b = tm.getTopFromReIntegrateList(), ref.getTopFromReIntegrateList()
            same_pop(a, b)
            if b is not None and not np.isneginf(b[1][0, 0]):
                tm.confirmIntegration(b[2]); ref.confirmIntegration(b[2])
            a, b = tm.getTopFromIntegrateList(), ref.getTopFromIntegrateList()
            same_pop(a, b)
            if b is not None:
                tm.confirmIntegration(b[1]); ref.confirmIntegration(b[1])
        same_state(tm, ref, added)
