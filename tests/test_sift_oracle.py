"""Known-answer tests pinning oracle/sift_oracle.c: exact integer dot products, the published SiftGPU
matching rule (mutual best, acos distance < distmax, ratio to the second best < ratiomax) restated brute-force in numpy, the
reference's tie-breaking, the 128-match cap and empty inputs."""
import numpy as np

from bundlefusion_b200 import synth
from oracle import oracle as orc

F = np.float32


def numpy_matches(d1, d2, distmax=0.7, ratiomax=0.8):
    """Brute-force rule, valid when no two dot products of a row / column tie for the maximum (checked by the caller)."""
    dot = d1.astype(np.int64) @ d2.astype(np.int64).T
    def best(m):       # per row of m: (argmax, max, second)
        order = np.argsort(-m, axis=1, kind="stable")
        mx = np.take_along_axis(m, order[:, :1], 1)[:, 0]
        nx = np.take_along_axis(m, order[:, 1:2], 1)[:, 0] if m.shape[1] > 1 else np.zeros(len(m), np.int64)
        return order[:, 0], mx, nx
    def accept(mx, nx):
        dist = np.arccos(np.minimum(mx.astype(F) * F(2.0 ** -18), F(1.0))).astype(F)
        distn = np.arccos(np.minimum(nx.astype(F) * F(2.0 ** -18), F(1.0))).astype(F)
        return (dist < F(distmax)) & (dist < distn * F(ratiomax)), dist
    ri, rmx, rnx = best(dot)
    rok, rdist = accept(rmx, rnx)
    ci, cmx, cnx = best(dot.T)
    cok, _ = accept(cmx, cnx)
    out = []
    for j in range(dot.shape[1]):
        if cok[j] and rok[ci[j]] and ri[ci[j]] == j:
            out.append((ci[j], j, rdist[ci[j]]))
    return out, dot


def test_dot_products_exact():
    d1, d2, _ = synth.make_sift_pair(200, 150, 80, seed=1)
    np.testing.assert_array_equal(orc.sift_multiply(d1, d2), d1.astype(np.int32) @ d2.astype(np.int32).T)
    # descriptors are normalised to 512: self products sit at 2^18 within quantisation
    self_dot = np.einsum("ik,ik->i", d1.astype(np.int64), d1.astype(np.int64))
    assert np.all(np.abs(self_dot - 262144) < 262144 * 0.02)


def test_planted_matches_equal_bruteforce_rule():
    for (n1, n2, nc, seed) in [(300, 260, 120, 2), (64, 64, 64, 3), (1000, 37, 30, 4), (5, 700, 5, 5)]:
        d1, d2, truth = synth.make_sift_pair(n1, n2, nc, noise=0.05, seed=seed)
        exp, dot = numpy_matches(d1, d2)
        # the brute-force rule is only the specification when maxima are unique
        srt = np.sort(dot, axis=1)
        assert n2 < 2 or np.all((srt[:, -1] > srt[:, -2]) | (srt[:, -1] == 0))
        idx, dist, count = orc.sift_match(d1, d2)
        assert count == len(exp) and count <= 128
        got = sorted((int(a), int(b)) for a, b in idx)
        assert got == sorted((int(a), int(b)) for a, b, _ in exp)
        dmap = {(int(a), int(b)): d for a, b, d in exp}
        for (a, b), d in zip(idx, dist):
            assert abs(dmap[(int(a), int(b))] - d) <= 2e-7
        # most planted correspondences are recovered, and nothing else
        tset = set(map(tuple, truth.tolist()))
        assert len(set(got) & tset) >= 0.85 * min(nc, 128) and len(set(got) - tset) <= max(2, 0.03 * len(got))


def test_tie_breaking_follows_the_reference_lanes():
    """Equal maxima: 32 strided lanes, then a tree that folds slot t + step into slot t with a strict '>', so the winner is the
    one with the smallest (bit-reversed lane, position), lane = position % 32; a duplicated maximum is its own second, so the
    ratio test only passes with ratiomax > 1."""
    dot = np.full((4, 70), 1000, np.int32)
    dot[0, [2, 33]] = 240000          # lane 2 (reversed 01000) vs lane 1 (reversed 10000) -> 2 wins
    dot[1, [5, 37, 69]] = 240000      # lanes 5, 5, 5 -> first position 5 wins
    dot[2, [64, 31]] = 240000         # lanes 0, 31 -> 64 wins
    dot[3, [1, 34]] = 240000          # lane 1 vs lane 2 -> 34 wins although it comes later
    res, dist = orc.sift_row_match(dot, 0.7, 1.5)
    assert res.tolist() == [2, 5, 64, 34]
    res, _ = orc.sift_row_match(dot, 0.7, 0.8)
    assert res.tolist() == [-1, -1, -1, -1]
    np.testing.assert_allclose(dist, np.arccos(F(240000) * F(2.0 ** -18)), rtol=0, atol=2e-7)


def test_cap_offsets_and_empty():
    d = synth.make_sift_descriptors(400, seed=9)
    idx, dist, count = orc.sift_match(d, d.copy(), offset=(1000, 2000))      # identical sets: every feature is a mutual best
    assert count == 400 and len(idx) == 128
    assert np.all(idx[:, 0] - 1000 == idx[:, 1] - 2000)
    assert np.all(np.diff(idx[:, 1].astype(np.int64)) > 0)                   # oracle order: ascending in image-2 feature
    assert np.all(dist < 0.12)                                               # acos of a self product: 0 up to the 8-bit quantisation of the norm
    assert orc.sift_match(d[:0], d)[2] == 0 and orc.sift_match(d, d[:0])[2] == 0
    z = np.zeros((10, 128), np.uint8)
    assert orc.sift_match(z, d[:10])[2] == 0                                 # all-zero dots never become candidates


def test_sort_matches_total_order():
    rng = np.random.default_rng(0)
    P = 6
    nm = np.array([0, 5, 128, 200, 1, 77], np.int32)                  # 200: the counter may exceed the 128 stored
    d = rng.random((P, 128)).astype(F); ix = rng.integers(0, 1000, (P, 128, 2)).astype(np.uint32)
    d[2, 10:20] = d[2, 10]                                           # equal distances: ordered by (image-2 feature, image-1 feature)
    sd, si = orc.sift_sort_matches(4, 1, P, nm, d, ix)               # pair 4 = curFrame is skipped, pair 0 is before startFrame
    for p in range(P):
        n = min(int(nm[p]), 128)
        if p in (0, 4):
            np.testing.assert_array_equal(sd[p], d[p]); continue
        order = np.lexsort((ix[p, :n, 0], ix[p, :n, 1], d[p, :n]))
        np.testing.assert_array_equal(sd[p, :n], d[p, :n][order])
        np.testing.assert_array_equal(si[p, :n], ix[p, :n][order])
        np.testing.assert_array_equal(sd[p, n:], d[p, n:])            # entries beyond the count are untouched
