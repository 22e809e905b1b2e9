"""CUDA SIFT detection (csrc/sift_detect.cu) through the C-ABI against oracle/sift_detect_oracle.c.  The pyramid, the DoG, the extrema
and the lists are bit-comparable (host-computed taps, fmaf in tap order, exact sqrtf); orientations and descriptors go through CUDA's
atan2f / expf / sinf / cosf and atomically ordered histogram sums, so they are compared with a small tolerance.

FIRST HARDWARE RUN PENDING: written when the round's GPU budget was already spent -- the kernels have so far run only under the CPU
emulation of tests/test_sift_detect_emulated.py (where they reproduce the oracle exactly).  The file name sorts it after every other GPU
test on purpose."""
import ctypes as C

import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from oracle import oracle as orc
from tests._cudart import DevBuf, device_count

pytestmark = pytest.mark.gpu


def texture(seed, H, W):
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    I = np.zeros((H, W))
    for s in (2.5, 5.0, 10.0, 20.0):
        n = gaussian_filter(rng.standard_normal((H, W)), (s, 1.5 * s), mode="wrap"); I += n / n.std()
    return np.clip(0.5 + 0.11 * I, 0, 1).astype(np.float32)


def run_gpu(I, D, **kw):
    if device_count() == 0:
        pytest.skip("no CUDA device")
    L = capi.lib()
    o = dict(depthMin=0.1, depthMax=3.0, minKeyScale=3.0, featureCountThreshold=150, maxKeyPoints=1024); o.update(kw)
    P = capi.BFSiftDetectParams(I.shape[1], I.shape[0], D.shape[1], D.shape[0], o["depthMin"], o["depthMax"], o["minKeyScale"], o["featureCountThreshold"], o["maxKeyPoints"])
    d_I, d_D = DevBuf(I.astype(np.float32)), DevBuf(D.astype(np.float32))
    d_kp, d_des = DevBuf(np.zeros((o["maxKeyPoints"], 4), np.float32)), DevBuf(np.zeros((o["maxKeyPoints"], 128), np.uint8))
    d_n, d_lc = DevBuf(np.zeros(1, np.int32)), DevBuf(np.zeros(12, np.int32))
    capi.check(L.bfSiftDetect(C.byref(P), d_I.ptr, d_D.ptr, d_kp.ptr, d_des.ptr, d_n.ptr, d_lc.ptr), "sift detect")
    n = int(d_n.get()[0])
    return d_kp.get()[:n], d_des.get()[:n], d_lc.get()


@pytest.mark.parametrize("seed,H,W,opts", [
    (1, 480, 640, dict()),
    (2, 480, 640, dict(minKeyScale=0.0, featureCountThreshold=100000, maxKeyPoints=4096)),
    (3, 96, 128, dict(minKeyScale=0.0, featureCountThreshold=100000)),
    (4, 240, 320, dict(minKeyScale=2.0, featureCountThreshold=60, maxKeyPoints=48)),
])
def test_detect_matches_oracle(seed, H, W, opts):
    I = texture(seed, H, W)
    rng = np.random.default_rng(seed)
    D = np.full((H, W), 1.5, np.float32)
    D[rng.random(D.shape) < 0.05] = -np.inf
    D[: H // 8] = 3.5
    ko, do, lo = orc.sift_detect(I, D, **opts)
    kg, dg, lg = run_gpu(I, D, **opts)
    # the same key points bit for bit (positions, scales, depths); a key point appears once per orientation, and a second histogram peak that
    # sits exactly at 0.8 x the first may be decided differently by CUDA's expf / atan2f: allow one such feature in fifty
    from collections import Counter
    co, cg = Counter(map(tuple, ko.tolist())), Counter(map(tuple, kg.tolist()))
    assert set(co) == set(cg), (len(ko), len(kg), sorted(set(co) ^ set(cg))[:5])
    slack = max(1, len(ko) // 50)
    assert sum(((co - cg) + (cg - co)).values()) <= slack and int(np.abs(lg - lo).sum()) <= slack, (lg, lo)
    worst = []
    for i in range(len(kg)):
        grp = np.nonzero((ko == kg[i]).all(1))[0]
        worst.append(min(int(np.abs(do[j].astype(np.int32) - dg[i].astype(np.int32)).max()) for j in grp))
    worst = np.sort(np.array(worst))[:len(worst) - slack]
    assert worst.max() <= 4 and (worst <= 1).mean() > 0.9, (worst.max(), (worst <= 1).mean())


def test_detect_rejects_unsupported_sizes_and_null_pointers():
    if device_count() == 0:
        pytest.skip("no CUDA device")
    L = capi.lib()
    P = capi.BFSiftDetectParams(100, 96, 100, 96, 0.1, 3.0, 3.0, 150, 64)
    d = DevBuf(np.zeros((96, 100), np.float32)); k = DevBuf(np.zeros((64, 4), np.float32)); s = DevBuf(np.zeros((64, 128), np.uint8)); n = DevBuf(np.zeros(1, np.int32))
    assert L.bfSiftDetect(C.byref(P), d.ptr, d.ptr, k.ptr, s.ptr, n.ptr, None) != 0
    P.width = 128
    assert L.bfSiftDetect(C.byref(P), None, d.ptr, k.ptr, s.ptr, n.ptr, None) != 0
