"""GPU parity tests of the Kabsch match filter (row a19) through the C-ABI against the CPU oracle: the same filtered matches in the same
order, the same distances, transforms within 1e-6 (both sides run the same fp32 operations; only acosf / cosf of the condition-number
test come from different libraries), for planted inlier / outlier problems, rejected pairs, skipped / empty pairs, and after the GPU's
own distance sort of shuffled raw matches."""
import ctypes as C

import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from bundlefusion_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def gpu_filter(dev, pb, num, dists, idxs, keys, start=0, sort_first=False, min_num=5, max_res2=0.0004):
    import torch
    L = capi.lib(); L.bfSetStream(None)
    P = pb["P"]
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    t_k, t_n, t_d, t_i = tt(keys), tt(num), tt(dists), tt(idxs.view(np.int32))
    nf = torch.full((P,), -7, dtype=torch.int32, device=dev); fd = torch.zeros(P, 25, device=dev); fi = torch.zeros(P, 25, 2, dtype=torch.int32, device=dev)
    T = torch.zeros(P, 16, device=dev); Ti = torch.zeros(P, 16, device=dev)
    Ki = (C.c_float * 16)(*pb["Kinv"].reshape(-1).tolist())
    torch.cuda.synchronize()
    if sort_first:
        capi.check(L.bfSiftSortKeyPointMatches(pb["cur"], start, P, t_n.data_ptr(), t_d.data_ptr(), t_i.data_ptr()), "sort")
    capi.check(L.bfSiftFilterKeyPointMatches(pb["cur"], start, P, t_k.data_ptr(), t_n.data_ptr(), t_d.data_ptr(), t_i.data_ptr(), nf.data_ptr(), fd.data_ptr(),
                                             fi.data_ptr(), T.data_ptr(), Ti.data_ptr(), Ki, min_num, max_res2), "filter")
    torch.cuda.synchronize()
    return nf.cpu().numpy(), fd.cpu().numpy(), fi.cpu().numpy().view(np.uint32), T.cpu().numpy().reshape(P, 4, 4), Ti.cpu().numpy().reshape(P, 4, 4)


def assert_same(g, o, pairs):
    for p in pairs:
        assert g[0][p] == o[0][p], f"pair {p}: count {g[0][p]} vs {o[0][p]}"
        np.testing.assert_array_equal(g[2][p], o[2][p])
        np.testing.assert_array_equal(g[1][p], o[1][p])
        np.testing.assert_allclose(g[3][p], o[3][p], rtol=0, atol=1e-6)
        np.testing.assert_allclose(g[4][p], o[4][p], rtol=0, atol=1e-5)


@pytest.mark.parametrize("seed,n_in,n_out,noise", [(1, 40, 12, 0.002), (2, 20, 30, 0.004), (3, 60, 0, 0.0005), (4, 6, 3, 0.002)])
def test_filter_matches_oracle(cuda_device, seed, n_in, n_out, noise):
    pb = synth.make_filter_problem(n_pairs=7, n_inliers=n_in, n_outliers=n_out, noise=noise, seed=seed)
    g = gpu_filter(cuda_device, pb, pb["num"], pb["dists"], pb["idxs"], pb["keys"])
    o = orc.sift_filter_matches(pb["cur"], 0, pb["P"], pb["keys"], pb["num"], pb["dists"], pb["idxs"], pb["Kinv"])
    assert_same(g, o, range(pb["P"] - 1))
    assert g[0][pb["cur"]] == -7                                       # the current frame's slot is not touched
    assert (o[0][:pb["P"] - 1] >= 5).sum() >= 5


def test_rejections_empty_and_start_offset(cuda_device):
    pb = synth.make_filter_problem(n_pairs=4, n_inliers=40, n_outliers=0, noise=0.001, seed=2)
    num = pb["num"].copy(); num[0] = 4; num[3] = 0
    idxs = pb["idxs"].copy(); keys = pb["keys"].copy()
    n = pb["n"]
    idxs[1, :n, 1] = pb["cur"] * n + np.random.default_rng(3).permutation(n)
    keys[2 * n:3 * n, 1] = 240.0 + 0.01 * np.arange(n); keys[2 * n:3 * n, 3] = 1.5
    g = gpu_filter(cuda_device, pb, num, pb["dists"], idxs, keys, start=1)
    o = orc.sift_filter_matches(pb["cur"], 1, pb["P"], keys, num, pb["dists"], idxs, pb["Kinv"])
    assert_same(g, o, [1, 2])
    assert g[0][0] == -7 and g[0][3] == 0 and g[0][1] == 0 and g[0][2] == 0


def test_sort_then_filter_chain(cuda_device):
    """Raw matches in arbitrary (append) order: the library's sort + filter equals the oracle's sort + filter."""
    pb = synth.make_filter_problem(n_pairs=5, n_inliers=35, n_outliers=15, noise=0.002, seed=9)
    rng = np.random.default_rng(4)
    d, ix = pb["dists"].copy(), pb["idxs"].copy()
    for p in range(pb["P"] - 1):
        perm = rng.permutation(pb["n"])
        d[p, :pb["n"]] = d[p, perm]; ix[p, :pb["n"]] = ix[p, perm]
    g = gpu_filter(cuda_device, pb, pb["num"], d, ix, pb["keys"], sort_first=True)
    sd, si = orc.sift_sort_matches(pb["cur"], 0, pb["P"], pb["num"], d, ix)
    o = orc.sift_filter_matches(pb["cur"], 0, pb["P"], pb["keys"], pb["num"], sd, si, pb["Kinv"])
    assert_same(g, o, range(pb["P"] - 1))


def test_filter_to_residuals_to_bundle_adjustment(cuda_device):
    """The chain the path exists for: raw matches -> sort -> Kabsch filter -> AddCurrToResiduals -> sparse BA.  The EntryJ list equals the
    oracle's bit for bit, and the solver recovers the poses the correspondences were generated from."""
    import torch
    from bundlefusion_b200.solver import CUDASolverBundling
    dev = cuda_device
    pb = synth.make_filter_problem(n_pairs=6, n_inliers=45, n_outliers=10, noise=0.001, seed=11)
    P, cur = pb["P"], pb["cur"]
    g = gpu_filter(dev, pb, pb["num"], pb["dists"], pb["idxs"], pb["keys"], sort_first=True)
    L = capi.lib(); L.bfSetStream(None)
    nf, fi = torch.from_numpy(g[0].copy()).to(dev), torch.from_numpy(g[2].view(np.int32).copy()).to(dev)
    nf[cur] = 0
    keys = torch.from_numpy(pb["keys"]).to(dev)
    cap = 25 * P
    ent = torch.zeros(cap * 32, dtype=torch.uint8, device=dev); eidx = torch.zeros(cap, 2, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    Ki = (C.c_float * 16)(*pb["Kinv"].reshape(-1).tolist())
    torch.cuda.synchronize()
    capi.check(L.bfSiftAddCurrToResiduals(cur, 0, P, ent.data_ptr(), eidx.data_ptr(), cnt.data_ptr(), nf.data_ptr(), fi.data_ptr(), keys.data_ptr(), Ki), "add")
    torch.cuda.synchronize()
    n = int(cnt.item())
    nf_h = g[0].copy(); nf_h[cur] = 0
    o_ent, o_idx = orc.sift_add_residuals(cur, 0, P, nf_h, g[2], pb["keys"], pb["Kinv"])
    assert n == len(o_ent) > 60
    np.testing.assert_array_equal(ent.cpu().numpy()[:32 * n], o_ent.view(np.uint8).reshape(-1))
    np.testing.assert_array_equal(eidx.cpu().numpy()[:n].view(np.uint32), o_idx)
    # bundle adjustment on those correspondences: image k's pose = T_k (maps frame k into the current frame's coordinates), image 0 fixed
    gt = pb["T_gt"]
    rot0 = np.zeros((P, 3), np.float32); tr0 = np.zeros((P, 3), np.float32)
    rng = np.random.default_rng(1)
    for k in range(P):
        Tk = gt[k] if k == 0 else synth.se3_exp(rng.standard_normal(3) * 0.01, rng.standard_normal(3) * 0.02) @ gt[k]
        r, t = synth.se3_log(Tk); rot0[k], tr0[k] = r, t
    rot, trans = torch.from_numpy(rot0.copy()).to(dev), torch.from_numpy(tr0.copy()).to(dev)
    s = CUDASolverBundling(P, 1000 * P, dev)
    s.solve(ent, n, torch.ones(P, dtype=torch.int32, device=dev), P, 4, 100, [1.0] * 4, d_rotationAnglesUnknowns=rot, d_translationUnknowns=trans)
    torch.cuda.synchronize()
    for k in range(P):
        Tk = synth.se3_exp(rot.cpu().numpy()[k], trans.cpu().numpy()[k])
        np.testing.assert_allclose(Tk, gt[k], atol=1e-2)
