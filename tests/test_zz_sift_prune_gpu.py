"""Correspondence / frame invalidation (csrc/sift_prune.cu) through the C-ABI on the GPU against the oracle, bit for bit.

FIRST HARDWARE RUN PENDING: written when the round's GPU budget was already spent -- verified so far under the CPU emulation only
(tests/test_sift_prune_emulated.py, tests/test_filters_emulated.py).  The file name sorts it after the other GPU tests on purpose."""
import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from oracle import oracle as orc
from tests._cudart import DevBuf, device_count
from tests.test_sift_prune_emulated import ENTRY, make_entries

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n", [(0, 1), (1, 127), (3, 100000)])
def test_invalidate_image_to_image(seed, n):
    if device_count() == 0:
        pytest.skip("no CUDA device")
    L = capi.lib()
    e = make_entries(np.random.default_rng(seed), n, 40)
    valid = e[e["i"] != 0xFFFFFFFF]
    pair = (int(valid["i"][0]), int(valid["j"][0])) if len(valid) else (3, 4)
    d = DevBuf(e.view(np.uint8))
    capi.check(L.bfSiftInvalidateImageToImage(d.ptr, n, pair[0], pair[1]), "invalidate")
    assert d.get().tobytes() == orc.sift_invalidate_image_to_image(e, *pair).tobytes()


@pytest.mark.parametrize("comprehensive", [0, 1])
def test_check_for_invalid_frames(comprehensive):
    if device_count() == 0:
        pytest.skip("no CUDA device")
    L = capi.lib()
    rng = np.random.default_rng(5)
    numVars, n = 500, 20000
    e = make_entries(rng, n, numVars)
    rows = rng.integers(0, 4, numVars).astype(np.int32); rows[0] = 3
    valid = np.ones(numVars, np.int32)
    wv, we = orc.sift_check_invalid_frames(rows, valid, e, comprehensive)
    d_rows, d_valid, d_e = DevBuf(rows), DevBuf(valid), DevBuf(e.view(np.uint8))
    capi.check(L.bfSiftCheckForInvalidFrames(d_rows.ptr, d_valid.ptr, numVars, d_e.ptr, n, comprehensive), "check invalid frames")
    assert np.array_equal(d_valid.get(), wv) and d_e.get().tobytes() == we.tobytes()


@pytest.mark.parametrize("seed,numFrames", [(1, 7), (2, 300), (3, 2000)])
def test_filter_frames(seed, numFrames):
    """bfSiftFilterFrames (csrc/sift_filter.cu) -- same pending-hardware status as the kernels above."""
    if device_count() == 0:
        pytest.skip("no CUDA device")
    L = capi.lib()
    rng = np.random.default_rng(seed)
    for trial in range(4):
        cur = int(rng.integers(0, numFrames)); start = int(rng.integers(0, cur + 1)) if trial % 2 else 0
        nf = (rng.integers(0, 12, numFrames) * (rng.random(numFrames) < 0.3)).astype(np.int32)
        if trial == 2:
            nf[:] = 0
        valid = (rng.random(numFrames) < 0.8).astype(np.int32)
        want_last, want_valid = orc.sift_filter_frames(cur, start, numFrames, nf, valid)
        d_nf, d_valid, d_last = DevBuf(nf), DevBuf(valid), DevBuf(np.full(1, 12345, np.int32))
        capi.check(L.bfSiftFilterFrames(cur, start, numFrames, d_nf.ptr, d_valid.ptr, d_last.ptr), "filter frames")
        assert int(d_last.get()[0]) == want_last and np.array_equal(d_valid.get(), want_valid)
