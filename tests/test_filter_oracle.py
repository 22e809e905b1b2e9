"""Known-answer tests pinning oracle/filter_oracle.c (row a19, Kabsch match filter): on planted inlier / outlier matches the filter keeps
only inliers, returns at most 25 of them, and its transform agrees with a float64 numpy Kabsch (np.linalg.svd) of the kept set and with
the ground truth; degenerate and too-small sets are rejected."""
import numpy as np

from bundlefusion_b200 import synth
from oracle import oracle as orc


def kabsch64(src, tgt):
    p0, q0 = src.mean(0), tgt.mean(0)
    Hm = (src - p0).T @ (tgt - q0) / len(src)
    U, S, Vt = np.linalg.svd(Hm)
    D = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = q0 - R @ p0
    return T


def points3d(pb, idx):
    k = pb["keys"][idx].astype(np.float64)
    Ki = pb["Kinv"].astype(np.float64)
    v = np.c_[k[:, 3] * k[:, 0], k[:, 3] * k[:, 1], k[:, 3]]
    return v @ Ki[:3, :3].T + Ki[:3, 3]


def test_filter_keeps_inliers_and_estimates_the_transform():
    pb = synth.make_filter_problem(n_pairs=6, n_inliers=40, n_outliers=12, noise=0.002, seed=1)
    nf, fd, fi, T, Ti = orc.sift_filter_matches(pb["cur"], 0, pb["P"], pb["keys"], pb["num"], pb["dists"], pb["idxs"], pb["Kinv"])
    n = pb["n"]
    assert nf[pb["cur"]] == 0
    for p in range(pb["P"] - 1):
        c = int(nf[p])
        assert 5 <= c <= 25
        kept = fi[p, :c]
        local = kept[:, 0] - p * n
        assert np.all(local < pb["n_inliers"]), "an outlier survived"
        assert np.array_equal(kept[:, 1] - pb["cur"] * n, local)
        assert np.all(fi[p, c:] == 0xFFFFFFFF) and np.all(fd[p, c:] == 999.0)
        src, tgt = points3d(pb, kept[:, 0]), points3d(pb, kept[:, 1])
        T64 = kabsch64(src, tgt)
        np.testing.assert_allclose(T[p], T64, atol=2e-4)
        np.testing.assert_allclose(T[p], pb["T_gt"][p], atol=2e-2)
        np.testing.assert_allclose(Ti[p] @ T[p], np.eye(4), atol=1e-5)
        res = np.sum((src @ T[p][:3, :3].T.astype(np.float64) + T[p][:3, 3] - tgt) ** 2, 1)
        assert res.max() < 0.0004 and np.all(np.diff(res) >= -1e-7)          # stored in ascending residual order
        # every kept match's distance is one of the raw distances of that pair (the arrays travel together through the sorts)
        raw = {(int(a), int(b)): float(d) for (a, b), d in zip(pb["idxs"][p, :n], pb["dists"][p, :n])}
        for (a, b), d in zip(kept, fd[p, :c]):
            assert raw[(int(a), int(b))] == float(d)


def test_filter_rejects_bad_sets():
    pb = synth.make_filter_problem(n_pairs=3, n_inliers=40, n_outliers=0, noise=0.001, seed=2)
    num = pb["num"].copy(); num[0] = 4                                # fewer raw matches than minNumMatches
    idxs = pb["idxs"].copy(); keys = pb["keys"].copy()
    n = pb["n"]
    rng = np.random.default_rng(3)                                    # pair 1: target side scrambled -> no rigid transform fits
    idxs[1, :n, 1] = pb["cur"] * n + rng.permutation(n)
    keys[2 * n:3 * n, 1] = 240.0 + 0.01 * np.arange(n)                # pair 2: source points (almost) collinear -> condition number test
    keys[2 * n:3 * n, 3] = 1.5
    nf, fd, fi, T, Ti = orc.sift_filter_matches(pb["cur"], 0, pb["P"], keys, num, pb["dists"], idxs, pb["Kinv"])
    assert nf[0] == 0 and nf[1] == 0 and nf[2] == 0
    assert np.all(fd[:3] == 999.0)


def test_add_residuals_builds_the_solver_input():
    pb = synth.make_filter_problem(n_pairs=4, n_inliers=30, n_outliers=5, seed=5)
    nf, fd, fi, T, Ti = orc.sift_filter_matches(pb["cur"], 0, pb["P"], pb["keys"], pb["num"], pb["dists"], pb["idxs"], pb["Kinv"])
    ent, eidx = orc.sift_add_residuals(pb["cur"], 0, pb["P"], nf, fi, pb["keys"], pb["Kinv"])
    assert len(ent) == int(nf[:pb["cur"]].sum())
    k = 0
    for p in range(pb["P"] - 1):                                       # ascending pair order, filtered order within a pair
        for m in range(int(nf[p])):
            assert ent["i"][k] == p and ent["j"][k] == pb["cur"]
            np.testing.assert_array_equal(eidx[k], fi[p, m])
            np.testing.assert_allclose(ent["pi"][k], points3d(pb, fi[p, m:m + 1, 0])[0], atol=1e-5)
            np.testing.assert_allclose(ent["pj"][k], points3d(pb, fi[p, m:m + 1, 1])[0], atol=1e-5)
            # the filter's transform maps the source point onto the target point within the residual bound
            r = T[p][:3, :3].astype(np.float64) @ ent["pi"][k] + T[p][:3, 3] - ent["pj"][k]
            assert r @ r < 0.0004
            k += 1
