"""Pins the oracle's Kabsch code against the REFERENCE's own code run on the CPU.  kabsch(), filterKeyPointMatches() and the 3x3 SVD
behind them (FL/SiftGPU/cuda_kabsch.h, cuda_svd3.h) and MYEIGEN::eigenSystem (cuda_SVD.h) are `__host__`-callable as written; compiled by
g++ from the sources under /root/reference (oracle/build_ref.py -> oracle/_ref/libref_kabsch_host.so) they produced the committed golden
file tests/golden/kabsch_reference_host.npz (scripts/make_golden_kabsch_host.py).  The oracle must reproduce it BIT FOR BIT when its
rsqrt is switched to the reference's host flavour (the SSE estimate; the device flavour is CUDA's rsqrtf and the oracle's default is
1 / sqrtf) -- transforms, singular values, and the complete output of the greedy match filter: count, indices, order, distances."""
import ctypes as C
import os

import numpy as np
import pytest

from bundlefusion_b200 import synth
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "kabsch_reference_host.npz")
CANARY_INPUTS = [0.5, 1.0, 2.0, 3.0, 7.25, 1e-3, 123.456, 9.87e5]


def kabsch_cases():
    rng = np.random.default_rng(2024)
    for t in range(400):
        n = int(rng.integers(3, 26))
        src = rng.standard_normal((n, 3)).astype(np.float32)
        T = synth.se3_exp(rng.standard_normal(3) * 0.3, rng.standard_normal(3) * 0.5)
        tgt = (src @ T[:3, :3].T + T[:3, 3] + rng.standard_normal((n, 3)) * 0.003).astype(np.float32)
        if t % 7 == 0:
            tgt[:, 0] *= -1                                   # a reflection: the reference flips the third column
        if t % 11 == 0:
            src[:, 2] = 0; tgt = (src @ T[:3, :3].T + T[:3, 3]).astype(np.float32)     # planar
        yield np.ascontiguousarray(src), np.ascontiguousarray(tgt)


def filter_cases():
    rng = np.random.default_rng(77)
    for seed in range(24):
        pb = synth.make_filter_problem(n_pairs=5, n_inliers=int(rng.integers(6, 55)), n_outliers=int(rng.integers(0, 25)),
                                       noise=float(rng.choice([0.0005, 0.002, 0.004])), seed=seed)
        keys = np.ascontiguousarray(pb["keys"], np.float32); Ki = np.ascontiguousarray(pb["Kinv"], np.float32)
        for p in range(pb["P"] - 1):
            yield keys, np.ascontiguousarray(pb["idxs"][p], np.uint32), np.ascontiguousarray(pb["dists"][p], np.float32), int(min(pb["num"][p], 128)), Ki


def eigen_cases():
    rng = np.random.default_rng(5)
    for t in range(400):
        A = rng.standard_normal((int(rng.integers(3, 26)), 3)).astype(np.float32) * rng.uniform(0.01, 2, 3).astype(np.float32)
        M = ((A - A.mean(0)).T @ (A - A.mean(0)) / len(A)).astype(np.float32)
        yield np.ascontiguousarray(((M + M.T) / 2).astype(np.float32))


@pytest.fixture
def host_rsqrt():
    L = orc.lib()
    vp = C.c_void_p
    L.orc_set_rsqrt_host_estimate.argtypes = [C.c_int]
    L.orc_rsqrt_host_estimate.argtypes = [C.c_float]; L.orc_rsqrt_host_estimate.restype = C.c_float
    L.orc_kabsch.argtypes = [vp, vp, C.c_uint, vp, vp]
    L.orc_filter_pair.argtypes = [vp, vp, vp, C.c_uint, vp, vp, C.c_uint, C.c_float]; L.orc_filter_pair.restype = C.c_uint
    L.orc_eigen_system3.argtypes = [vp, vp, vp]
    g = np.load(GOLDEN)
    mine = np.array([L.orc_rsqrt_host_estimate(float(x)) for x in CANARY_INPUTS], np.float32)
    if not np.array_equal(mine, g["canary"]):
        pytest.skip("this CPU's _mm_rsqrt_ss estimate differs from the one the golden file was made on (vendor-specific table)")
    L.orc_set_rsqrt_host_estimate(1)
    yield L, g
    L.orc_set_rsqrt_host_estimate(0)


def test_kabsch_bit_identical_to_reference_host_code(host_rsqrt):
    L, g = host_rsqrt
    for k, (src, tgt) in enumerate(kabsch_cases()):
        T = np.zeros(16, np.float32); e = np.zeros(3, np.float32)
        L.orc_kabsch(src.ctypes.data, tgt.ctypes.data, len(src), T.ctypes.data, e.ctypes.data)
        assert np.array_equal(T.view(np.uint32), g["kabsch_T"][k].view(np.uint32)), k
        assert np.array_equal(e.view(np.uint32), g["kabsch_evs"][k].view(np.uint32)), k


def test_match_filter_bit_identical_to_reference_host_code(host_rsqrt):
    L, g = host_rsqrt
    kept = 0
    for k, (keys, idx, dist, n, Ki) in enumerate(filter_cases()):
        i, d, T = idx.copy(), dist.copy(), np.zeros(16, np.float32)
        c = L.orc_filter_pair(keys.ctypes.data, i.ctypes.data, d.ctypes.data, n, T.ctypes.data, Ki.ctypes.data, 5, np.float32(0.0004))
        assert c == g["filter_count"][k], k
        assert np.array_equal(i[:c], g["filter_idx"][k][:c]) and np.array_equal(d[:c].view(np.uint32), g["filter_dist"][k][:c].view(np.uint32)), k
        if c > 0:
            assert np.array_equal(T.view(np.uint32), g["filter_T"][k].view(np.uint32)), k
            kept += 1
    assert kept > 60                                           # most pairs of the synthetic problems do pass the filter


def test_eigen_system_bit_identical_to_reference_host_code():
    """MYEIGEN::eigenSystem (Jacobi, rows of the rotation matrix, magnitude sort) needs no rsqrt: identical on any CPU."""
    L = orc.lib(); L.orc_eigen_system3.argtypes = [C.c_void_p] * 3
    g = np.load(GOLDEN)
    for k, M in enumerate(eigen_cases()):
        e = np.zeros(3, np.float32); v = np.zeros(9, np.float32)
        assert L.orc_eigen_system3(M.ctypes.data, e.ctypes.data, v.ctypes.data) == g["eig_ok"][k]
        assert np.array_equal(e.view(np.uint32), g["eig_vals"][k].view(np.uint32)) and np.array_equal(v.view(np.uint32), g["eig_vecs"][k].view(np.uint32)), k


def test_default_rsqrt_stays_close_to_the_host_flavour(host_rsqrt):
    """The oracle's default (1 / sqrtf, standing for the device's rsqrtf) against the 12-bit host estimate: same decisions on well-posed
    fits, transforms within the estimate's accuracy."""
    L, g = host_rsqrt
    L.orc_set_rsqrt_host_estimate(0)
    worst = 0.0
    for k, (src, tgt) in enumerate(kabsch_cases()):
        if len(src) < 6 or k % 7 == 0 or k % 11 == 0:
            continue
        T = np.zeros(16, np.float32); e = np.zeros(3, np.float32)
        L.orc_kabsch(src.ctypes.data, tgt.ctypes.data, len(src), T.ctypes.data, e.ctypes.data)
        worst = max(worst, float(np.abs(T - g["kabsch_T"][k]).max()))
    assert worst < 5e-3, worst
