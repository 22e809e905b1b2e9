"""GPU parity tests of the dense-cache frame builder (row a20) through the C-ABI against the CPU oracle: every output array of
CUDACache::storeFrame bit-identical (the arithmetic contract of oracle/cache_oracle.c), for the reference's default filters, with
the filters off, on noisy frames with holes, and for an unusual cache size; plus the use the path makes of it (local BA on caches
built by the library instead of by the host generator)."""
import numpy as np
import pytest

from bundlefusion_b200 import synth
from bundlefusion_b200.cache import CUDACache
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
F = np.float32
KEYS = ("depth", "campos", "normals", "normalsU", "intensity", "intensityDerivs")


def K_of(W, H):
    fx = 525.0 * W / 640.0
    K = np.eye(4, dtype=F); K[0, 0] = K[1, 1] = fx; K[0, 2] = (W - 1) / 2.0; K[1, 2] = (H - 1) / 2.0
    return K


def build(dev, frames, W, H, cw, ch, **kw):
    import torch
    cache = CUDACache(W, H, cw, ch, len(frames), K_of(W, H), dev, **kw)
    for d, c, _ in frames:
        cache.storeFrame(torch.from_numpy(d).to(dev), W, H, torch.from_numpy(c).to(dev), W, H)
    torch.cuda.synchronize()
    return cache


def assert_equal_frames(g, o):
    for k in KEYS:
        a, b = g[k], o[k]
        if a.dtype.kind == "f":
            np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=k)       # bit patterns (-inf, -0.0 included)
        else:
            np.testing.assert_array_equal(a, b, err_msg=k)


@pytest.mark.parametrize("W,H,cw,ch,kw", [(640, 480, 80, 60, {}), (320, 240, 80, 60, {"colorDownSigma": 0.0, "depthDownSigmaD": 0.0}),
                                          (640, 480, 160, 120, {}), (200, 150, 37, 23, {"colorDownSigma": 1.2, "depthDownSigmaD": 1.6})])
def test_store_frame_matches_oracle_bit_for_bit(cuda_device, W, H, cw, ch, kw):
    frames = [synth.make_frame(60 * i + 5, W, H) for i in range(2)]              # sensor noise + dropout holes
    frames[1][0][H // 3: H // 3 + 9, W // 4: W // 4 + 30] = -np.inf
    cache = build(cuda_device, frames, W, H, cw, ch, **kw)
    for k, (d, c, _) in enumerate(frames):
        assert_equal_frames(cache.download(k), orc.cache_store_frame(d, c, K_of(W, H), cw, ch, **kw))
    fx, fy, mx, my = cache.intrinsics
    assert abs(fx - 525.0 * W / 640.0 * cw / W) < 1e-3 and abs(mx - (W - 1) / 2.0 * (cw - 1) / (W - 1)) < 1e-3


def test_local_ba_on_library_built_caches(cuda_device):
    """The dense term consumes caches built by bfCacheStoreFrame exactly as it consumes host-built ones: same solve through the
    oracle on the oracle's caches, poses within 1e-4 relative L2."""
    import torch
    from bundlefusion_b200.solver import CUDASolverBundling
    dev = cuda_device
    W, H = 320, 240
    prob = synth.make_dense_ba_problem(6, stride=3, W=W, H=H)
    frames = [synth.make_frame(100 + 3 * k, W, H, noise=False, dropout=0.0) for k in range(6)]
    kw = {"colorDownSigma": 2.5, "depthDownSigmaD": 1.0, "depthDownSigmaR": 0.05}
    cache = build(dev, frames, W, H, 80, 60, **kw)
    ocaches = [orc.cache_store_frame(d, c, K_of(W, H), 80, 60, **kw) for d, c, _ in frames]
    N = 6
    corr = torch.from_numpy(prob["corr"].view(np.uint8).reshape(-1).copy()).to(dev)
    rot = torch.from_numpy(prob["init_rot"].copy()).to(dev); trans = torch.from_numpy(prob["init_trans"].copy()).to(dev)
    valid = torch.ones(N, dtype=torch.int32, device=dev)
    s = CUDASolverBundling(N, 1000 * N, dev)
    s.solve(corr, len(prob["corr"]), valid, N, 2, 100, [1.0, 1.0], [1.0, 2.0], [0.0, 0.0], d_rotationAnglesUnknowns=rot, d_translationUnknowns=trans, cudaCache=cache)
    torch.cuda.synchronize()
    o = orc.solve(prob["corr"], prob["init_rot"], prob["init_trans"], 2, 100, [1.0, 1.0], [1.0, 2.0], [0.0, 0.0], ocaches, cache.intrinsics)
    x_g, x_o = np.c_[rot.cpu().numpy(), trans.cpu().numpy()], np.c_[o["rot"], o["trans"]]
    assert s.getStats()["dense_weighted_pairs"] == o["weighted_pairs"] > 3
    assert np.linalg.norm(x_g - x_o) / np.linalg.norm(x_o) < 1e-4
