"""Correspondence / frame invalidation kernels (csrc/sift_prune.cu) under the CPU emulation of tests/cuda_emu, against the oracle and by hand."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from tests.cuda_emu import build_emulated

ENTRY = np.dtype([("i", "<u4"), ("j", "<u4"), ("pi", "<f4", 3), ("pj", "<f4", 3)])


@pytest.fixture(scope="module")
def emu():
    L = build_emulated("sift_prune.cu", 2)
    L.bfSiftInvalidateImageToImage.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint]
    L.bfSiftCheckForInvalidFrames.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_int]
    return L


def make_entries(rng, n, numImages):
    e = np.zeros(n, ENTRY)
    i = rng.integers(0, numImages - 1, n); j = i + 1 + rng.integers(0, 3, n)
    e["i"], e["j"] = i, np.minimum(j, numImages - 1)
    e["pi"] = rng.standard_normal((n, 3)); e["pj"] = rng.standard_normal((n, 3))
    e["i"][rng.random(n) < 0.05] = 0xFFFFFFFF                       # some already invalid
    e["j"][e["i"] == 0xFFFFFFFF] = 0xFFFFFFFF
    return e


@pytest.mark.parametrize("seed,n", [(0, 1), (1, 127), (2, 128), (3, 1000)])
def test_invalidate_image_to_image(emu, seed, n):
    rng = np.random.default_rng(seed)
    e = make_entries(rng, n, 12)
    valid = e[e["i"] != 0xFFFFFFFF]
    pair = (int(valid["i"][0]), int(valid["j"][0])) if len(valid) else (3, 4)
    want = orc.sift_invalidate_image_to_image(e, *pair)
    got = e.copy()
    assert emu.bfSiftInvalidateImageToImage(got.ctypes.data, n, pair[0], pair[1]) == 0
    assert got.tobytes() == want.tobytes()
    hit = (e["i"] == pair[0]) & (e["j"] == pair[1])
    assert hit.any() == bool(len(valid)) and np.all(got["i"][hit] == 0xFFFFFFFF) and np.all(got["j"][hit] == 0xFFFFFFFF)
    assert np.array_equal(got["pi"], e["pi"]) and np.array_equal(got[~hit], e[~hit])          # positions untouched, other pairs untouched
    # the pair is directed: (j, i) matches nothing
    got2 = e.copy(); emu.bfSiftInvalidateImageToImage(got2.ctypes.data, n, pair[1], pair[0])
    assert got2.tobytes() == e.tobytes()


@pytest.mark.parametrize("comprehensive", [0, 1])
@pytest.mark.parametrize("seed,n,numVars", [(0, 300, 12), (1, 40, 200), (2, 0, 5)])
def test_check_for_invalid_frames(emu, seed, n, numVars, comprehensive):
    rng = np.random.default_rng(seed)
    e = make_entries(rng, n, numVars) if n else np.zeros(0, ENTRY)
    rows = rng.integers(0, 4, numVars).astype(np.int32)                # some rows empty
    rows[0] = 3
    valid = np.ones(numVars, np.int32); valid[rng.integers(0, numVars)] = 0
    wv, we = orc.sift_check_invalid_frames(rows, valid, e, comprehensive)
    gv, ge = valid.copy(), e.copy()
    assert emu.bfSiftCheckForInvalidFrames(rows.ctypes.data, gv.ctypes.data, numVars, ge.ctypes.data if n else None, n, comprehensive) == 0
    assert np.array_equal(gv, wv) and ge.tobytes() == we.tobytes()
    assert np.all(gv[rows == 0] == 0) and np.array_equal(gv[rows != 0], valid[rows != 0])
    if not comprehensive:
        assert ge.tobytes() == e.tobytes()
    elif n:
        dead = np.isin(e["i"], np.nonzero(rows == 0)[0]) | np.isin(e["j"], np.nonzero(rows == 0)[0])
        was_valid = e["i"] != 0xFFFFFFFF
        assert np.all(ge["i"][dead & was_valid] == 0xFFFFFFFF) and np.array_equal(ge[~dead], e[~dead])


def test_null_arguments(emu):
    assert emu.bfSiftInvalidateImageToImage(None, 0, 1, 2) == 0
    assert emu.bfSiftInvalidateImageToImage(None, 5, 1, 2) != 0
    assert emu.bfSiftCheckForInvalidFrames(None, None, 4, None, 0, 0) != 0
