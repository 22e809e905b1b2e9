"""Oracle of SIFT detection / description (oracle/sift_detect_oracle.c, row a17): known answers on synthetic blobs, the reference's
gating rules (depth, minimum scale, feature-count limit), and the invariances the algorithm has by construction (translation by the
coarsest grid step: bit-identical; rotation by 90 degrees: same key points, matching descriptors)."""
import numpy as np
import pytest

from oracle import oracle as orc

H, W = 480, 640


def blob_image(blobs, base=0.5):
    yy, xx = np.mgrid[0:H, 0:W]
    I = np.full((H, W), base, np.float64)
    for (cx, cy, s, a) in blobs:
        I += a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    return I.astype(np.float32)


def texture_image(seed):
    """Smooth random texture with structure at every octave: band-limited anisotropic noise at four scales."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    I = np.zeros((H, W))
    for s, aniso in ((2.5, (1.0, 1.7)), (5.0, (1.6, 1.0)), (10.0, (1.0, 1.5)), (20.0, (1.4, 1.0))):
        n = gaussian_filter(rng.standard_normal((H, W)), (s * aniso[0], s * aniso[1]), mode="wrap")
        I += n / n.std()
    return np.clip(0.5 + 0.11 * I, 0, 1).astype(np.float32)


FLAT = np.full((H, W), 1.5, np.float32)
BLOBS = [(100, 120, 4.0, 0.4), (300, 200, 6.0, -0.3), (500, 350, 9.0, 0.35), (200, 400, 3.0, 0.3), (420, 100, 14.0, -0.35)]


def test_filter_bank_follows_the_sigma_schedule():
    sig, wid, taps = orc.sift_filter_bank()
    k = 2.0 ** (1.0 / 3.0); s0 = 1.6 * k
    assert abs(sig[0] - np.sqrt(1.6 ** 2 - 0.5 ** 2)) < 1e-5                       # initial smoothing: from the camera's 0.5 to sigma0 k^-1 = 1.6
    # level i of an octave carries sigma0 k^(i-1): the incremental blur from level i-1 is sqrt of the difference of squares ...
    want = [np.sqrt((s0 * k ** i) ** 2 - (s0 * k ** (i - 1)) ** 2) for i in range(0, 5)]
    assert np.allclose(sig[1:], want, rtol=1e-5)
    assert list(wid) == [13, 11, 13, 17, 21, 25]                                   # 2 ceil(4 sigma - 0.5) + 1
    for i in range(6):
        t = taps[i, :wid[i]]
        assert abs(t.sum() - 1.0) < 1e-6 and np.array_equal(t, t[::-1]) and np.all(taps[i, wid[i]:] == 0) and t.argmax() == wid[i] // 2


def test_blobs_are_found_at_their_position_and_scale():
    kp, des, lc = orc.sift_detect(blob_image(BLOBS), FLAT, minKeyScale=0.0)
    assert lc.sum() == len(kp) and len(kp) >= len(BLOBS)
    for (cx, cy, s, a) in BLOBS:
        d = np.hypot(kp[:, 0] - 0.5 - cx, kp[:, 1] - 0.5 - cy)
        near = kp[d < 4.0]
        assert len(near) >= 1, (cx, cy)
        # no sub-scale refinement: the scale is the level's sigma (1.6 k^(j+1) 2^octave); a blob of std s peaks in the DoG near sigma ~ s
        assert np.all((near[:, 2] > 0.6 * s) & (near[:, 2] < 1.4 * s)), (s, near[:, 2])
        assert np.all(near[:, 3] == 1.5)
    # every key point sits on one of the blobs
    dmin = np.min([np.hypot(kp[:, 0] - 0.5 - cx, kp[:, 1] - 0.5 - cy) for (cx, cy, _, _) in BLOBS], axis=0)
    assert np.all(dmin < 4.0)
    # descriptors: SiftGPU convention, unit vector scaled by 512 and rounded, entries clamped near 0.2 * 512 before the second normalisation
    n2 = (des.astype(np.float64) ** 2).sum(1)
    assert np.all(np.abs(np.sqrt(n2) - 512) < 6)


def test_depth_gate_and_min_scale_and_count_limit():
    I = blob_image(BLOBS)
    D = FLAT.copy()
    D[:, :250] = -np.inf                       # no depth on the left: blobs at x = 100 and 200 disappear
    D[300:, 400:] = 4.5                        # beyond depthMax = 3: the blob at (500, 350) disappears
    kp, _, _ = orc.sift_detect(I, D, minKeyScale=0.0)
    assert len(kp) > 0 and np.all(kp[:, 0] > 250) and not np.any((kp[:, 0] > 400) & (kp[:, 1] > 300))
    # minimum scale (s_minKeyScale): only the coarse blobs stay
    kp2, _, _ = orc.sift_detect(I, FLAT, minKeyScale=5.0)
    assert len(kp2) > 0 and np.all(kp2[:, 2] >= 5.0)
    kp0, _, lc0 = orc.sift_detect(I, FLAT, minKeyScale=0.0)
    assert len(kp2) < len(kp0)
    # feature-count threshold: the lowest levels are dropped until at most `threshold` features besides the next level remain (SiftPyramid.cpp:245-254)
    T = texture_image(5)
    kpa, _, lca = orc.sift_detect(T, FLAT, minKeyScale=0.0, featureCountThreshold=100000)
    kpb, _, lcb = orc.sift_detect(T, FLAT, minKeyScale=0.0, featureCountThreshold=60)
    assert len(kpa) > 150 and len(kpb) < len(kpa)
    first = int(np.nonzero(lcb)[0][0])
    assert np.all(lcb[:first] == 0) and np.array_equal(lcb[first:], lca[first:])       # whole levels are dropped from the fine end, the rest is untouched
    assert lcb.sum() - lcb[first] <= 60 < lca[first - 1:].sum() if first > 0 else True


def test_unsupported_sizes_are_rejected():
    with pytest.raises(ValueError):
        orc.sift_detect(np.zeros((480, 600), np.float32), np.zeros((480, 600), np.float32))
    kp, des, lc = orc.sift_detect(np.full((H, W), 0.3, np.float32), FLAT)              # a constant image has no extrema
    assert len(kp) == 0 and lc.sum() == 0


def test_translation_by_the_coarsest_grid_step_is_exact():
    """Shifting the image by a multiple of 8 pixels (one pixel of the coarsest octave) shifts every key point by exactly that much and
    leaves its descriptor bit-identical wherever the wrapped-around border content is out of reach of the filters and sampling windows
    (checked for the two fine octaves, whose reach is ~150 pixels); coarser key points move with it and keep nearly the same descriptor."""
    T = texture_image(7)
    S = np.roll(T, (16, 24), axis=(0, 1))
    a = orc.sift_detect(T, FLAT, minKeyScale=0.0, featureCountThreshold=100000, maxKeyPoints=4096)
    b = orc.sift_detect(S, FLAT, minKeyScale=0.0, featureCountThreshold=100000, maxKeyPoints=4096)
    index_b = {}
    for j, (x, y, s, _) in enumerate(b[0]):
        index_b.setdefault((float(x), float(y), float(s)), []).append(j)
    margin, exact, approx, total = 170, 0, 0, 0
    for i, (x, y, s, _) in enumerate(a[0]):
        if not (margin < x < W - margin - 24 and margin < y < H - margin - 16):
            continue
        total += 1
        js = index_b.get((float(x + 24), float(y + 16), float(s)), [])
        if not js:
            continue
        da = a[1][i].astype(np.float64)
        dist = min(np.arccos(np.clip(da @ b[1][j].astype(np.float64) / (np.linalg.norm(da) * np.linalg.norm(b[1][j].astype(np.float64))), -1, 1)) for j in js)
        approx += dist < 0.05
        if s < 7.0:
            assert any(np.array_equal(a[1][i], b[1][j]) for j in js), (x, y, s)
            exact += 1
    assert exact >= 10 and approx >= 0.95 * total, (exact, approx, total)


def test_rotation_by_90_degrees_gives_matching_descriptors():
    T = texture_image(11)[:, 80:560]            # a square 480 x 480 crop ... padded back to a supported size below
    pad = lambda A: np.pad(A, ((0, 0), (0, 160)), constant_values=0.5)
    a = orc.sift_detect(pad(T), FLAT, minKeyScale=0.0, featureCountThreshold=100000, maxKeyPoints=4096)
    R = np.rot90(T).copy()                      # counter-clockwise: (x, y) -> (y, 479 - x)
    b = orc.sift_detect(pad(R), FLAT, minKeyScale=0.0, featureCountThreshold=100000, maxKeyPoints=4096)
    inner = lambda k: (k[:, 0] > 70) & (k[:, 0] < 410) & (k[:, 1] > 70) & (k[:, 1] < 410)
    ia, ib = np.nonzero(inner(a[0]))[0], np.nonzero(inner(b[0]))[0]
    assert len(ia) > 60 and len(ib) > 60
    pa = a[0][ia]; pb = b[0][ib]
    # the pixel-centre convention (x + 0.5) makes the map (x, y) -> (y, 480 - x); down-sampling keeps even samples, so coarse octaves see a
    # one-pixel-shifted lattice after the flip: compare positions with a tolerance of one coarse pixel and descriptors by distance
    mapped = np.c_[pa[:, 1], 480.0 - pa[:, 0]]
    matched = 0; close = 0
    for q, (m, s) in enumerate(zip(mapped, pa[:, 2])):
        d = np.hypot(pb[:, 0] - m[0], pb[:, 1] - m[1])
        cand = np.nonzero((d <= max(1.5, 0.6 * s)) & (np.abs(pb[:, 2] - s) < 1e-3))[0]
        if len(cand) == 0:
            continue
        close += 1
        da = a[1][ia[q]].astype(np.float64)
        best = min(np.arccos(np.clip((da @ b[1][ib[c]].astype(np.float64)) / (np.linalg.norm(da) * np.linalg.norm(b[1][ib[c]].astype(np.float64))), -1, 1)) for c in cand)
        matched += best < 0.35
    assert close > 0.7 * len(pa), (close, len(pa))
    assert matched > 0.8 * close, (matched, close)


def test_detect_match_filter_chain_recovers_a_known_motion():
    """A fronto-parallel textured plane at 1.5 m seen twice, the second time shifted by (24, 16) pixels: detection -> descriptor matching ->
    distance sort -> Kabsch filter (rows a17 -> a18 -> a19, all oracle) must return the rigid motion (24, 16) z / f with identity rotation."""
    T = texture_image(3)
    S = np.roll(T, (16, 24), axis=(0, 1))
    ka, da, _ = orc.sift_detect(T, FLAT, minKeyScale=3.0)
    kb, db, _ = orc.sift_detect(S, FLAT, minKeyScale=3.0)
    assert 100 < len(ka) <= 1024 and 100 < len(kb) <= 1024
    idx, dist, count = orc.sift_match(da, db)
    assert count >= 40
    n = len(idx)
    num = np.array([n, 0], np.int32)
    dists = np.full((2, 128), 999.0, np.float32); idxs = np.full((2, 128, 2), 0xFFFFFFFF, np.uint32)
    dists[0, :n] = dist; idxs[0, :n, 0] = idx[:, 0]; idxs[0, :n, 1] = idx[:, 1] + len(ka)
    d2, i2 = orc.sift_sort_matches(1, 0, 2, num, dists, idxs)
    fx = 525.0
    K = np.array([[fx, 0, 319.5, 0], [0, fx, 239.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    nf, fd, fi, Tm, _ = orc.sift_filter_matches(1, 0, 2, np.concatenate([ka, kb]), num, d2, i2, np.linalg.inv(K).astype(np.float32))
    assert nf[0] >= 15
    want = np.eye(4); want[0, 3] = 24 * 1.5 / fx; want[1, 3] = 16 * 1.5 / fx
    assert np.abs(Tm[0] - want).max() < 2e-3, Tm[0]
    # every surviving match really is the same surface point
    a = ka[fi[0, :nf[0], 0]]; b = kb[fi[0, :nf[0], 1] - len(ka)]
    assert np.all(np.abs(b[:, 0] - a[:, 0] - 24) < 1e-3) and np.all(np.abs(b[:, 1] - a[:, 1] - 16) < 1e-3)
