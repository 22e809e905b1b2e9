"""Known-answer tests pinning oracle/cache_oracle.c (row a20, CUDACache::storeFrame): analytic plane, hand-computed filter taps, and
agreement with the independent numpy generator synth.make_cache_frame when the filters are off."""
import numpy as np

from bundlefusion_b200 import synth
from oracle import oracle as orc

F = np.float32


def K_of(W, H):
    fx = 525.0 * W / 640.0
    K = np.eye(4, dtype=F); K[0, 0] = K[1, 1] = fx; K[0, 2] = (W - 1) / 2.0; K[1, 2] = (H - 1) / 2.0
    return K


def test_filters_off_equals_numpy_generator():
    d, c, _ = synth.make_frame(120, 320, 240)
    o = orc.cache_store_frame(d, c, K_of(320, 240), colorDownSigma=0.0, depthDownSigmaD=0.0)
    g = synth.make_cache_frame(d, c)
    np.testing.assert_array_equal(o["depth"], g["depth"])
    np.testing.assert_array_equal(o["normalsU"][..., 3], 0)
    for k, tol in (("campos", 2e-6), ("normals", 1e-4)):      # the generator divides by fx, the reference multiplies by 1/fx; normals are
        fin = np.isfinite(g[k])                               # cross products of differences of neighbouring positions (cancellation)
        np.testing.assert_array_equal(np.isfinite(o[k]), fin)
        np.testing.assert_allclose(o[k][fin], g[k][fin], rtol=tol, atol=tol)
    assert np.abs(o["normalsU"].astype(int) - g["normalsU"].astype(int)).max() <= 1
    np.testing.assert_allclose(o["intensity"], g["intensity"], rtol=0, atol=2e-7)
    fin = np.isfinite(g["intensityDerivs"])
    np.testing.assert_array_equal(np.isfinite(o["intensityDerivs"]), fin)
    np.testing.assert_allclose(o["intensityDerivs"][fin], g["intensityDerivs"][fin], rtol=0, atol=1e-6)


def test_fronto_parallel_plane():
    W, H = 160, 120
    d = np.full((H, W), 1.5, F); c = np.zeros((H, W, 4), np.uint8); c[..., 0] = 100; c[..., 1] = 50; c[..., 2] = 200
    o = orc.cache_store_frame(d, c, K_of(W, H))
    np.testing.assert_allclose(o["depth"], 1.5, rtol=0, atol=3e-7)                       # a Gaussian of a constant is the constant (to rounding)
    inner = o["normals"][1:-1, 1:-1]
    np.testing.assert_allclose(inner[..., :3], np.broadcast_to(np.array([0, 0, 1], F), inner[..., :3].shape), atol=2e-4)   # -(d/dy x d/dx) normalised = +z
    assert np.all(inner[..., 3] == 0)
    assert np.abs(o["normalsU"][1:-1, 1:-1, :3].astype(int) - np.array([128, 128, 255])).max() <= 1
    I = (F(0.299) * 100 + F(0.587) * 50 + F(0.114) * 200) / F(255)
    np.testing.assert_allclose(o["intensity"], I, atol=1e-6)
    np.testing.assert_allclose(o["intensityDerivs"][1:-1, 1:-1], 0, atol=1e-6)
    assert np.all(np.isinf(o["intensityDerivs"][0]))
    K = K_of(W, H)
    xs = (np.arange(80, dtype=F) * F((W - 1) / 79.0) + F(0.5)).astype(int)
    np.testing.assert_allclose(o["campos"][30, :, 0], (xs - K[0, 2]) / K[0, 0] * 1.5, atol=3e-6)


def test_range_gate_and_invalid_pixels():
    W, H = 64, 48
    d = np.full((H, W), 1.0, F)
    d[:, 32:] = 2.0                       # depth step: the gate |dc - d| < 0.05 must stop the filter from mixing the two sides
    d[10, 10] = -np.inf
    c = np.zeros((H, W, 4), np.uint8)
    o = orc.cache_store_frame(d, c, K_of(W, H), cw=64, ch=48)          # same resolution: the resample is the identity
    fin = o["depth"][np.isfinite(o["depth"])]
    assert np.all((np.abs(fin - 1.0) < 1e-6) | (np.abs(fin - 2.0) < 1e-6))
    assert o["depth"][10, 10] == -np.inf and np.all(np.isinf(o["campos"][10, 10]))
    for (y, x) in ((10, 9), (10, 11), (9, 10), (11, 10)):
        assert np.all(np.isinf(o["normals"][y, x]))                   # a neighbour of an invalid pixel has no normal
        assert abs(o["depth"][y, x] - 1.0) < 1e-6                     # but its own depth survives (invalid taps are skipped)
    # one filtered value by hand: 5x5 window of a noisy patch
    rng = np.random.default_rng(0)
    d2 = (1.0 + 0.01 * rng.standard_normal((H, W))).astype(F)
    o2 = orc.cache_store_frame(d2, c, K_of(W, H), cw=64, ch=48)
    y, x = 20, 20
    s = F(0); sw = F(0)
    for m in range(x - 2, x + 3):
        for n in range(y - 2, y + 3):
            if abs(F(d2[y, x]) - F(d2[n, m])) < F(0.05):
                wgt = F(np.exp(-F((m - x) ** 2 + (n - y) ** 2) / F(2.0)))
                sw = F(sw + wgt); s = F(s + F(wgt * d2[n, m]))
    assert abs(o2["depth"][y, x] - s / sw) <= 2e-7
