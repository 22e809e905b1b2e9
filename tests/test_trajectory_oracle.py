"""Known-answer tests pinning oracle/trajectory_oracle.c (row a22, FL/OnlineBundler.cu:6-140) against float64 numpy products."""
import numpy as np

from bundlefusion_b200 import synth
from oracle import oracle as orc

F = np.float32


def poses(n, seed):
    rng = np.random.default_rng(seed)
    return np.stack([synth.se3_exp(rng.standard_normal(3) * 0.3, rng.standard_normal(3)) for _ in range(n)]).astype(F)


def test_update_trajectory_and_init_next_global():
    G, per = 5, 11
    glob, loc = poses(G + 1, 1), poses(G * per, 2)
    inval = np.ones(G * (per - 1), np.int32); inval[[3, 17]] = 0
    out = orc.update_trajectory(glob, loc, per, inval)
    for k in range(len(inval)):
        if inval[k] == 0:
            assert np.all(np.isneginf(out[k]))
        else:
            g, l = k // (per - 1), k % (per - 1)
            np.testing.assert_allclose(out[k], glob[g].astype(np.float64) @ loc[g * per + l].astype(np.float64), atol=3e-6)
    g2 = orc.init_next_global(glob, 3, 2, loc, 9, per)
    np.testing.assert_allclose(g2[3], glob[2].astype(np.float64) @ loc[3 * per - (per - 9)].astype(np.float64), atol=3e-6)
    np.testing.assert_array_equal(g2[:3], glob[:3])


def test_compute_sift_transform_branches():
    n_all, cur, cur_all = 40, 7, 27                                    # frame 7 of its chunk, 27 overall
    sift, comp = poses(n_all, 3), poses(n_all, 4)
    finv = poses(cur, 5)
    nf = np.zeros(cur, np.int32); nf[[2, 4]] = 30                       # most recent matched frame of the chunk: 4
    prev = cur_all - (cur - 4)
    T = sift[prev].astype(np.float64) @ finv[4].astype(np.float64)
    for last_valid, expect in ((0, T), (prev + 5, comp[prev].astype(np.float64) @ finv[4].astype(np.float64)),
                               (prev - 3, comp[prev - 3].astype(np.float64) @ np.linalg.inv(sift[prev - 3].astype(np.float64)) @ sift[prev].astype(np.float64) @ finv[4].astype(np.float64))):
        traj, out = orc.compute_sift_transform(finv, nf, comp, last_valid, sift, cur_all, cur)
        np.testing.assert_allclose(traj[cur_all], T, atol=3e-6)
        np.testing.assert_allclose(out, expect, atol=2e-5)
        np.testing.assert_array_equal(np.delete(traj, cur_all, 0), np.delete(sift, cur_all, 0))
    # no matched frame in the chunk: nothing is written
    traj, out = orc.compute_sift_transform(finv, np.zeros(cur, np.int32), comp, 5, sift, cur_all, cur)
    np.testing.assert_array_equal(traj, sift); assert np.all(out == 0)


def test_select_reintegration_rule():
    """dist = |(2 t, w)_integrated - (2 t, w)_optimised|^2 on the SE(3) logs (the host's Pose is (translation, rotation) and the factor 2 lands on
    components 0..2: FL/PoseHelper.h:355-358, FL/TrajectoryManager.cpp:67-74); top-N integrated frames above the threshold, descending."""
    n = 50
    integ = poses(n, 7)
    rng = np.random.default_rng(8)
    opt = integ.copy()
    moved = {3: 0.05, 10: 0.2, 11: 0.03, 20: 0.1, 30: 0.4, 41: 0.0005}
    for k, m in moved.items():
        opt[k] = (synth.se3_exp(np.zeros(3), np.array([m, 0, 0])) @ integ[k].astype(np.float64)).astype(F)
    state = np.ones(n, np.int32); state[30] = 0                       # frame 30 moved most but is not integrated
    opt[20, 0, 0] = -np.inf                                          # frame 20 lost its transform
    dist, lst = orc.select_reintegration(opt, integ, state, 3, 0.0004)
    assert lst.tolist() == [10, 3, 11]
    dist, lst = orc.select_reintegration(opt, integ, state, 10, 0.0004)
    assert lst.tolist() == [10, 3, 11]                               # 41 moved 0.5 mm: dist 1e-6 < threshold; unmoved frames: 0
    assert dist[30] == -1 and dist[20] == -1
    # a left-multiplied pure translation changes t by exactly that translation (the rotation part of the log is unchanged)
    np.testing.assert_allclose(dist[10], 4 * 0.2 ** 2, rtol=2e-3)
    np.testing.assert_allclose(dist[5], 0.0, atol=1e-9)
