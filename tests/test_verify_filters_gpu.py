"""CUDA surface-area and dense-verification match filters (csrc/sift_verify.cu) through the C-ABI against oracle/filter_oracle.c, bit for
bit: every operation on the path is an individually rounded IEEE one (+, -, *, /, sqrtf, roundf), so areas, err / corr and the
decisions must be identical.  Device memory comes from the CUDA runtime directly (tests/_cudart.py); no torch needed."""
import ctypes as C

import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from bundlefusion_b200 import synth
from oracle import oracle as orc
from tests._cudart import DevBuf, device_count
from tests.test_verify_filters_oracle import VERIFY

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    if device_count() == 0:
        pytest.skip("no CUDA device")
    return capi.lib()


def _f16(m):
    return np.ascontiguousarray(m, np.float32).reshape(16).ctypes.data_as(C.POINTER(C.c_float))


def run_area(L, cur, start, P, keys, num, fidx, Kinv, thresh):
    d_keys, d_num, d_idx = DevBuf(keys.astype(np.float32)), DevBuf(num.astype(np.int32)), DevBuf(fidx.astype(np.uint32))
    d_areas = DevBuf(np.full((P, 2), -1.0, np.float32))
    capi.check(L.bfSiftFilterMatchesBySurfaceArea(cur, start, P, d_keys.ptr, d_num.ptr, d_idx.ptr, _f16(Kinv), thresh, d_areas.ptr), "surface area")
    return d_num.get(), d_areas.get()


def same(a, b):
    """float32 arrays equal bit for bit; NaNs must sit at the same places (0/0 is 0xFFC00000 on x86 and 0x7FFFFFFF on the GPU)"""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_surface_area_bit_exact(gpu, seed):
    pb = synth.make_area_problem(seed)
    for start, thresh in ((0, 0.032), (2, 0.032), (0, 1e9), (0, 0.0)):
        nf_o, ar_o = orc.sift_filter_surface_area(pb["cur"], start, pb["P"], pb["keys"], pb["num"], pb["fidx"], pb["Kinv"], thresh)
        nf_g, ar_g = run_area(gpu, pb["cur"], start, pb["P"], pb["keys"], pb["num"], pb["fidx"], pb["Kinv"], thresh)
        assert np.array_equal(nf_g, nf_o), (start, thresh, nf_g, nf_o)
        assert same(ar_g, ar_o), (start, thresh, ar_g, ar_o)


def test_surface_area_degenerate_and_empty(gpu):
    keys = np.array([[2, 1, 1, 1], [-2, 1, 1, 1], [2, -1, 1, 1], [-2, -1, 1, 1]] * 2, np.float32)
    num = np.array([4, 0], np.int32); fidx = np.full((2, 25, 2), 0xFFFFFFFF, np.uint32)
    fidx[0, :4, 0] = np.arange(4); fidx[0, :4, 1] = 4 + np.arange(4)
    I = np.eye(4, dtype=np.float32)
    nf_o, ar_o = orc.sift_filter_surface_area(1, 0, 2, keys, num, fidx, I, 0.032)
    nf_g, ar_g = run_area(gpu, 1, 0, 2, keys, num, fidx, I, 0.032)
    assert np.array_equal(nf_g, nf_o) and same(ar_g, ar_o) and nf_g[0] == 0
    # numFrames == startFrame: nothing launched, nothing touched
    nf_g, ar_g = run_area(gpu, 1, 2, 2, keys, num, fidx, I, 0.032)
    assert np.array_equal(nf_g, num) and np.all(ar_g == -1.0)
    assert gpu.bfSiftFilterMatchesBySurfaceArea(1, 0, 2, None, None, None, _f16(I), 0.032, None) != 0


def run_verify(L, cur, start, P, W, H, K, num, T, caches, opt):
    keep = []
    recs = (capi.BFCUDACachedFrame * len(caches))()
    for r, f in zip(recs, caches):
        d, c, n = DevBuf(f["depth"].astype(np.float32)), DevBuf(f["campos"].astype(np.float32)), DevBuf(f["normals"].astype(np.float32))
        keep += [d, c, n]
        r.d_depthDownsampled, r.d_cameraposDownsampled, r.d_normalsDownsampled = d.ptr, c.ptr, n.ptr
    d_recs = DevBuf(np.frombuffer(bytes(recs), np.uint8))
    d_num, d_T, d_stats = DevBuf(num.astype(np.int32)), DevBuf(T.astype(np.float32)), DevBuf(np.full((P, 2), -1.0, np.float32))
    capi.check(L.bfSiftFilterMatchesByDenseVerify(cur, start, P, W, H, _f16(K), d_num.ptr, d_T.ptr, d_recs.ptr, opt["distThresh"], opt["normalThresh"],
                                                  opt["colorThresh"], opt["errThresh"], opt["corrThresh"], opt["dMin"], opt["dMax"], d_stats.ptr), "dense verify")
    return d_num.get(), d_stats.get()


def test_dense_verify_bit_exact(gpu):
    pb = synth.make_dense_verify_problem()
    P, cur = pb["P"], pb["cur"]
    num = np.full(P, 7, np.int32); num[3] = 0
    for start in (0, 2):
        nf_o, st_o = orc.sift_filter_dense_verify(cur, start, P, pb["W"], pb["H"], pb["K"], num, pb["T"], pb["caches"], **VERIFY)
        nf_g, st_g = run_verify(gpu, cur, start, P, pb["W"], pb["H"], pb["K"], num, pb["T"], pb["caches"], VERIFY)
        assert np.array_equal(nf_g, nf_o), (nf_g, nf_o, st_g, st_o)
        assert same(st_g, st_o), (st_g, st_o)
    assert nf_g[2] == 7 and nf_g[4] == 7 and nf_g[3] == 0


def test_dense_verify_invalid_frames_and_bad_transforms(gpu):
    pb = synth.make_dense_verify_problem(n_prev=3)
    P, cur = pb["P"], pb["cur"]
    num = np.array([5, 5, 5, 5], np.int32)
    caches = list(pb["caches"])
    caches[2] = {k: np.full_like(v, -np.inf) for k, v in caches[2].items() if k in ("depth", "campos", "normals")}
    T = pb["T"].copy(); T[0] = np.diag([1, 1, -1, 1]).astype(np.float32) @ T[0]
    nf_o, st_o = orc.sift_filter_dense_verify(cur, 0, P, pb["W"], pb["H"], pb["K"], num, T, caches, **VERIFY)
    nf_g, st_g = run_verify(gpu, cur, 0, P, pb["W"], pb["H"], pb["K"], num, T, caches, VERIFY)
    assert np.array_equal(nf_g, nf_o) and same(st_g, st_o)
    assert list(nf_g) == [0, 0, 0, 5] and np.isnan(st_g[2, 0])


def run_verify_trajectory(L, N, valid, traj, W, H, K, caches, opt):
    keep = []
    recs = (capi.BFCUDACachedFrame * len(caches))()
    for r, f in zip(recs, caches):
        d, c, n = DevBuf(f["depth"].astype(np.float32)), DevBuf(f["campos"].astype(np.float32)), DevBuf(f["normals"].astype(np.float32))
        keep += [d, c, n]
        r.d_depthDownsampled, r.d_cameraposDownsampled, r.d_normalsDownsampled = d.ptr, c.ptr, n.ptr
    d_recs = DevBuf(np.frombuffer(bytes(recs), np.uint8))
    d_valid, d_T, d_ok = DevBuf(np.ascontiguousarray(valid, np.int32)), DevBuf(np.ascontiguousarray(traj, np.float32)), DevBuf(np.full(1, 7, np.int32))
    d_stats = DevBuf(np.full((max(1, N * (N - 1) // 2), 2), -1.0, np.float32))
    capi.check(L.bfSiftVerifyTrajectory(N, d_valid.ptr, d_T.ptr, W, H, _f16(K), d_recs.ptr, opt["distThresh"], opt["normalThresh"], opt["colorThresh"],
                                        opt["errThresh"], opt["corrThresh"], opt["dMin"], opt["dMax"], d_ok.ptr, d_stats.ptr), "verify trajectory")
    return int(d_ok.get()[0]), d_stats.get()


def trajectory_verify_case(n_prev=4, break_pair=False, invalid=None):
    """cached frames of a synthetic sweep with their true camera poses as the trajectory (so every pair agrees); break_pair moves one pose"""
    pb = synth.make_dense_verify_problem(n_prev=n_prev)
    N = pb["P"]
    # trajectory[p] = frame p -> the current frame (index cur = N - 1), from the true poses: a consistent trajectory, trajectory[cur] = identity
    traj = np.stack([np.linalg.inv(pb["gt"][pb["cur"]]) @ pb["gt"][p] for p in range(N)]).astype(np.float32)
    traj[pb["cur"]] = np.eye(4, dtype=np.float32)
    if break_pair:
        traj[1] = traj[1].copy(); traj[1][:3, 3] += np.array([0.25, 0.0, 0.1], np.float32)
    valid = np.ones(N, np.int32)
    if invalid is not None:
        valid[invalid] = 0
    return pb, N, valid, traj


@pytest.mark.parametrize("break_pair,invalid", [(False, None), (True, None), (True, 1)])
def test_verify_trajectory_bit_exact(gpu, break_pair, invalid):
    pb, N, valid, traj = trajectory_verify_case(4, break_pair, invalid)
    opt = dict(VERIFY, errThresh=0.05, corrThresh=0.001, dMin=0.1, dMax=3.0)          # s_verifyOptErrThresh / CorrThresh, FL/Bundler.cpp:266
    ok_o, st_o = orc.sift_verify_trajectory(N, valid, traj, pb["W"], pb["H"], pb["K"], pb["caches"], **opt)
    ok_g, st_g = run_verify_trajectory(gpu, N, valid, traj, pb["W"], pb["H"], pb["K"], pb["caches"], opt)
    assert ok_g == ok_o and same(st_g, st_o), (ok_g, ok_o, st_g, st_o)
    assert ok_o == (0 if (break_pair and invalid is None) else 1)
    # the reference's block decode reaches only row-major pair indices below N (N - 1) / 2: with N = 5, (0,1..4), (1,2..4); pair (2,3) is never looked at
    visited = {(b // N, b % N) for b in range(N * (N - 1) // 2) if b // N < b % N}
    assert (2, 3) not in visited and (1, 4) in visited
    assert all((st_o[b, 0] != -1.0) == (((b // N, b % N) in visited) and valid[b // N] and valid[b % N]) for b in range(N * (N - 1) // 2))


def test_verify_trajectory_too_few_images(gpu):
    pb, N, valid, traj = trajectory_verify_case(4)
    ok_g, _ = run_verify_trajectory(gpu, 1, valid, traj, pb["W"], pb["H"], pb["K"], pb["caches"], dict(VERIFY))
    assert ok_g == 0
