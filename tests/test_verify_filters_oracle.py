"""Oracle of the surface-area and dense-verification match filters (oracle/filter_oracle.c, row a19) against independent float64
restatements written from the reference's description (FL/SiftGPU/cuda_surfaceArea.h, cuda_SVD.h jacobi, SIFTImageManager.cu:413-585)
and against known answers."""
import numpy as np

from bundlefusion_b200 import synth
from oracle import oracle as orc


# ---- float64 restatement of the surface-area measure -------------------------------------------------------------------------
def jacobi_rows(A):
    """Cyclic Jacobi (Numerical Recipes) on a symmetric 3x3; returns (eigenvalues, rotation matrix V whose COLUMNS are eigenvectors)."""
    a = np.array(A, np.float64); n = 3
    v = np.eye(n); d = np.diag(a).copy(); b = d.copy(); z = np.zeros(n)
    for sweep in range(1, 51):
        sm = abs(a[0, 1]) + abs(a[0, 2]) + abs(a[1, 2])
        if sm == 0.0:
            return d, v
        tresh = 0.2 * sm / 9 if sweep < 4 else 0.0
        for p in range(n - 1):
            for q in range(p + 1, n):
                g = 100.0 * abs(a[p, q])
                if sweep > 4 and abs(d[p]) + g == abs(d[p]) and abs(d[q]) + g == abs(d[q]):
                    a[p, q] = 0.0
                elif abs(a[p, q]) > tresh:
                    h = d[q] - d[p]
                    if abs(h) + g == abs(h):
                        t = a[p, q] / h
                    else:
                        theta = 0.5 * h / a[p, q]
                        t = 1.0 / (abs(theta) + np.sqrt(1.0 + theta * theta))
                        if theta < 0:
                            t = -t
                    c = 1.0 / np.sqrt(1 + t * t); s = t * c; tau = s / (1.0 + c)
                    h = t * a[p, q]
                    z[p] -= h; z[q] += h; d[p] -= h; d[q] += h
                    a[p, q] = 0.0

                    def rot(m, i, j, k, l):
                        g_, h_ = m[i, j], m[k, l]
                        m[i, j] = g_ - s * (h_ + g_ * tau); m[k, l] = h_ + s * (g_ - h_ * tau)
                    for j in range(p):
                        rot(a, j, p, j, q)
                    for j in range(p + 1, q):
                        rot(a, p, j, j, q)
                    for j in range(q + 1, n):
                        rot(a, p, j, q, j)
                    for j in range(n):
                        rot(v, j, p, j, q)
        b += z; d = b.copy(); z[:] = 0
    return None, None


def area_f64(pts):
    """pts [n,3] camera-space key points of one image."""
    n = len(pts)
    mean = pts.mean(0)
    V = (pts - mean).T @ (pts - mean) / n
    d, v = jacobi_rows(V)
    if d is None:
        return 0.0
    ev = [v[i].copy() for i in range(3)]                      # the reference takes ROWS of the rotation matrix (cuda_SVD.h:94-99)
    d = list(d)
    for i in range(3):                                        # selection by |eigenvalue| with row exchange (:101-118)
        j = max(range(i, 3), key=lambda k: (abs(d[k]), -k))
        if abs(d[j]) > 0 and j != i:
            d[i], d[j] = d[j], d[i]; ev[i], ev[j] = ev[j], ev[i]
    s = (pts - ((pts - mean) @ ev[2])[:, None] * ev[2]) - mean
    q = np.c_[s @ ev[0], s @ ev[1]]
    m2 = q.mean(0); c = (q - m2).T @ (q - m2) / n
    disc = 0.5 * np.sqrt((c[0, 0] - c[1, 1]) ** 2 + 4 * c[0, 1] ** 2)
    ax = []
    for lam in ((c[0, 0] + c[1, 1]) / 2 + disc, (c[0, 0] + c[1, 1]) / 2 - disc):
        w = np.array([-c[0, 1], c[0, 0] - lam])
        with np.errstate(invalid="ignore", divide="ignore"):
            ax.append(w / np.linalg.norm(w))
    o = np.c_[q @ ax[0], q @ ax[1]]
    e = o.max(0) - o.min(0)
    if e[0] < 1e-5 or e[1] < 1e-5:
        return 0.0
    return float(e[0] * e[1])


def cam_points(pb, p, which):
    n = int(pb["num"][p])
    k = pb["keys"][pb["fidx"][p, :n, which]].astype(np.float64)
    Ki = pb["Kinv"].astype(np.float64)
    v = np.c_[k[:, 3] * k[:, 0], k[:, 3] * k[:, 1], k[:, 3]]
    return v @ Ki[:3, :3].T + Ki[:3, 3]


def test_surface_area_matches_the_float64_restatement():
    for seed in range(4):
        pb = synth.make_area_problem(seed)
        nf, areas = orc.sift_filter_surface_area(pb["cur"], 0, pb["P"], pb["keys"], pb["num"], pb["fidx"], pb["Kinv"], 0.032)
        for p in range(pb["P"] - 1):
            if pb["num"][p] == 0:
                assert np.all(areas[p] == -1.0)                # not visited
                continue
            for which in (0, 1):
                want = area_f64(cam_points(pb, p, which))
                assert abs(areas[p, which] - want) <= 2e-3 * max(want, 1e-4) + 1e-7, (seed, p, which, areas[p, which], want)


def test_surface_area_decisions():
    pb = synth.make_area_problem(1)
    nf, areas = orc.sift_filter_surface_area(pb["cur"], 0, pb["P"], pb["keys"], pb["num"], pb["fidx"], pb["Kinv"], 0.032)   # GlobalBundlingState: s_surfAreaPcaThresh
    assert list(nf[:8]) == [20, 0, 8, 0, 25, 3, 0, 0]
    assert areas[0].min() > 0.2 and areas[4].min() > 0.2       # metres^2, key points spread over a 1..3 m deep view
    assert areas[1].max() < 1e-3                               # a 4-pixel patch at 1.5 m
    assert areas[2, 0] < 1e-3 and areas[2, 1] > 0.032          # small in ONE image only: kept
    assert areas[6].max() == 0.0 and areas[7].max() < 1e-6     # two points / a line: an extent below 1e-5 gives area 0
    assert np.all(areas[pb["cur"]] == -1.0)
    # the start offset skips pairs, the current frame is never touched
    nf2, ar2 = orc.sift_filter_surface_area(pb["cur"], 2, pb["P"], pb["keys"], pb["num"], pb["fidx"], pb["Kinv"], 0.032)
    assert nf2[1] == 6 and np.all(ar2[:2] == -1.0) and list(nf2[2:8]) == [8, 0, 25, 3, 0, 0]
    # a huge threshold removes everything that was visited and has finite areas
    nf3, _ = orc.sift_filter_surface_area(pb["cur"], 0, pb["P"], pb["keys"], pb["num"], pb["fidx"], pb["Kinv"], 1e9)
    assert not nf3[:8].any()


def test_surface_area_axis_aligned_covariance_degenerates_to_zero():
    """Four corners of a rectangle with identity intrinsics: the 3-D covariance is exactly diagonal (Jacobi returns the identity), the
    2-D one too, and its first eigenvector is 0/0.  The NaN coordinates then lose every fminf / fmaxf against the idle lanes'
    +-FLT_MAX (warpReduceMin / Max over 32 lanes, at most 25 of them busy), the extent is -inf and the area 0: the reference
    rejects such a pair whatever its true extent, and so does the restatement."""
    keys = np.array([[2, 1, 1, 1], [-2, 1, 1, 1], [2, -1, 1, 1], [-2, -1, 1, 1]] * 2, np.float32)
    num = np.array([4, 0], np.int32); fidx = np.full((2, 25, 2), 0xFFFFFFFF, np.uint32)
    fidx[0, :4, 0] = np.arange(4); fidx[0, :4, 1] = 4 + np.arange(4)
    nf, areas = orc.sift_filter_surface_area(1, 0, 2, keys, num, fidx, np.eye(4, dtype=np.float32), 0.032)
    assert np.all(areas[0] == 0.0) and nf[0] == 0
    # nudged off the axes the same rectangle measures its 4 x 2 extent
    keys2 = keys.copy(); keys2[0, 0] += 0.25; keys2[4, 0] += 0.25
    nf, areas = orc.sift_filter_surface_area(1, 0, 2, keys2, num, fidx, np.eye(4, dtype=np.float32), 0.032)
    assert nf[0] == 4 and np.all(np.abs(areas[0] - 8.0) < 1.0)


# ---- dense verification --------------------------------------------------------------------------------------------------------
def proj_error_f64(T, K, fin, fmodel, W, H, distThresh, normalThresh, dMin, dMax):
    p = fin["campos"].reshape(-1, 4).astype(np.float64); n = fin["normals"].reshape(-1, 4).astype(np.float64); d = fin["depth"].reshape(-1).astype(np.float64)
    ok = (p[:, 0] != -np.inf) & (n[:, 0] != -np.inf) & (d >= dMin) & (d <= dMax)
    out = np.zeros((len(p), 3))
    T = T.astype(np.float64); K = K.astype(np.float64)
    mp = fmodel["campos"].reshape(-1, 4).astype(np.float64); mn = fmodel["normals"].reshape(-1, 4).astype(np.float64); md = fmodel["depth"].reshape(-1).astype(np.float64)
    for i in np.nonzero(ok)[0]:
        pt = T @ p[i]; nt = T[:, :3] @ n[i, :3]
        t = K[:3, :3] @ pt[:3] + K[:3, 3]
        sx, sy = int(np.floor(abs(t[0] / t[2]) + 0.5) * np.sign(t[0] / t[2])), int(np.floor(abs(t[1] / t[2]) + 0.5) * np.sign(t[1] / t[2]))
        if not (0 <= sx < W and 0 <= sy < H):
            continue
        m = sy * W + sx
        if mp[m, 0] == -np.inf or mn[m, 0] == -np.inf:
            continue
        dist = np.linalg.norm(pt - mp[m]); dN = nt[:3] @ mn[m, :3]
        if not (dMin <= md[m] <= dMax):
            continue
        bad = (pt[2] < md[m]) and dist > distThresh
        if (dN >= normalThresh and dist <= distThresh) or bad:
            w = max(0.0, 0.5 * ((1 - dist / distThresh) + (1 - (pt[2] - dMin) / (dMax - dMin))))
            out[i] = (dist, w, 1.0)
    return out


VERIFY = dict(distThresh=0.15, normalThresh=0.97, colorThresh=0.1, errThresh=0.075, corrThresh=0.02, dMin=0.1, dMax=4.0)   # GlobalBundlingState: s_verifyOpt*


def test_dense_verify_matches_float64_and_separates_good_from_bad_transforms():
    pb = synth.make_dense_verify_problem()
    P, cur = pb["P"], pb["cur"]
    num = np.full(P, 7, np.int32)
    nf, stats = orc.sift_filter_dense_verify(cur, 0, P, pb["W"], pb["H"], pb["K"], num, pb["T"], pb["caches"], **VERIFY)
    for p in range(P - 1):
        a = proj_error_f64(pb["T"][p], pb["K"], pb["caches"][p], pb["caches"][cur], pb["W"], pb["H"], VERIFY["distThresh"], VERIFY["normalThresh"], VERIFY["dMin"], VERIFY["dMax"])
        b = proj_error_f64(np.linalg.inv(pb["T"][p].astype(np.float64)), pb["K"], pb["caches"][cur], pb["caches"][p], pb["W"], pb["H"], VERIFY["distThresh"], VERIFY["normalThresh"], VERIFY["dMin"], VERIFY["dMax"])
        # the block total as the reference's kernel forms it -- not a plain sum, see tests/test_manager_reference_emulated.py: reference_reduction
        from tests.test_manager_reference_emulated import reference_reduction
        tot = reference_reduction((a + b).astype(np.float32), pb["W"], pb["H"]).astype(np.float64)
        corr = 0.5 * tot[2] / (pb["W"] * pb["H"])
        # a handful of pixels sit on a rounding / threshold edge and may fall differently in float32
        assert abs(stats[p, 1] - corr) < 0.01, (p, stats[p], corr)
        if tot[1] > 0:
            assert abs(stats[p, 0] - tot[0] / tot[1]) < 0.05 * max(tot[0] / tot[1], 0.01), (p, stats[p], tot)
    good, bad = [p for p in range(P - 1) if p % 2 == 0], [p for p in range(P - 1) if p % 2 == 1]
    assert all(nf[p] == 7 for p in good) and all(nf[p] == 0 for p in bad), (nf, stats)
    assert all(stats[p, 0] < 0.03 and stats[p, 1] > 0.3 for p in good), stats
    assert nf[cur] == 7 and np.all(stats[cur] == -1.0)


def test_dense_verify_skips_and_edge_cases():
    pb = synth.make_dense_verify_problem(n_prev=3)
    P, cur = pb["P"], pb["cur"]
    num = np.array([5, 0, 5, 5], np.int32)
    nf, stats = orc.sift_filter_dense_verify(cur, 1, P, pb["W"], pb["H"], pb["K"], num, pb["T"], pb["caches"], **VERIFY)
    assert np.all(stats[0] == -1.0) and nf[0] == 5             # before startFrame
    assert np.all(stats[1] == -1.0) and nf[1] == 0             # no filtered matches: not evaluated
    assert nf[2] == 5 and stats[2, 1] > 0.3
    # a frame without any valid pixel: no correspondences, err = 0/0 = NaN -> rejected (SIFTImageManager.cu:577)
    empty = {k: np.full_like(v, -np.inf) for k, v in pb["caches"][2].items() if k in ("depth", "campos", "normals")}
    caches = list(pb["caches"]); caches[2] = empty
    nf, stats = orc.sift_filter_dense_verify(cur, 0, P, pb["W"], pb["H"], pb["K"], num, pb["T"], caches, **VERIFY)
    assert nf[2] == 0 and np.isnan(stats[2, 0]) and stats[2, 1] == 0.0
    # a transform that throws everything behind the camera: projections land outside / NaN pixel rule, pair rejected
    T = pb["T"].copy(); T[0] = np.diag([1, 1, -1, 1]).astype(np.float32) @ T[0]
    nf, _ = orc.sift_filter_dense_verify(cur, 0, P, pb["W"], pb["H"], pb["K"], num, T, pb["caches"], **VERIFY)
    assert nf[0] == 0
