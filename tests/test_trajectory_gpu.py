"""GPU parity tests of the trajectory glue stubs (row a22) against the CPU oracle: products bit-identical; the one branch that inverts
a pose within 2e-5 (the oracle inverts with a different, equally valid cofactor arrangement)."""
import ctypes as C

import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from oracle import oracle as orc
from tests.test_trajectory_oracle import poses

pytestmark = pytest.mark.gpu
F = np.float32


def test_stubs_match_oracle(cuda_device):
    import torch
    dev = cuda_device
    L = capi.lib(); L.bfSetStream(None)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # updateTrajectoryCU / initNextGlobalTransformCU
    G, per = 7, 11
    glob, loc = poses(G + 1, 1), poses(G * per, 2)
    inval = np.ones(G * (per - 1), np.int32); inval[[3, 17, 40]] = 0
    d_g, d_l, d_i = tt(glob), tt(loc), tt(inval)
    d_c = torch.zeros(len(inval) * 16, device=dev)
    torch.cuda.synchronize()
    L.updateTrajectoryCU(d_g.data_ptr(), G, d_c.data_ptr(), len(inval), d_l.data_ptr(), per, G, d_i.data_ptr())
    L.initNextGlobalTransformCU(d_g.data_ptr(), 3, 2, d_l.data_ptr(), 9, per)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_c.cpu().numpy().reshape(-1, 4, 4).view(np.uint32), orc.update_trajectory(glob, loc, per, inval).view(np.uint32))
    np.testing.assert_array_equal(d_g.cpu().numpy().view(np.uint32), orc.init_next_global(glob, 3, 2, loc, 9, per).view(np.uint32))
    # computeSiftTransformCU, all three branches
    n_all, cur, cur_all = 40, 7, 27
    sift, comp, finv = poses(n_all, 3), poses(n_all, 4), poses(cur, 5)
    nf = np.zeros(cur, np.int32); nf[[2, 4]] = 30
    prev = cur_all - (cur - 4)
    for last_valid, exact in ((0, True), (prev + 5, True), (prev - 3, False)):
        d_s, d_out = tt(sift), torch.zeros(16, device=dev)
        d_f, d_n, d_cm = tt(finv), tt(nf), tt(comp)
        torch.cuda.synchronize()
        L.computeSiftTransformCU(d_f.data_ptr(), d_n.data_ptr(), d_cm.data_ptr(), last_valid, d_s.data_ptr(), cur_all, cur, d_out.data_ptr())
        torch.cuda.synchronize()
        traj, out = orc.compute_sift_transform(finv, nf, comp, last_valid, sift, cur_all, cur)
        np.testing.assert_array_equal(d_s.cpu().numpy().view(np.uint32), traj.view(np.uint32))
        if exact:
            np.testing.assert_array_equal(d_out.cpu().numpy().reshape(4, 4).view(np.uint32), out.view(np.uint32))
        else:
            np.testing.assert_allclose(d_out.cpu().numpy().reshape(4, 4), out, atol=2e-5)
    # curFrameIndex == 0: no launch, nothing written
    d_out = torch.zeros(16, device=dev)
    L.computeSiftTransformCU(d_f.data_ptr(), d_n.data_ptr(), d_cm.data_ptr(), 0, d_s.data_ptr(), 5, 0, d_out.data_ptr())
    torch.cuda.synchronize()
    assert float(d_out.abs().sum()) == 0.0


def test_select_reintegration_matches_oracle(cuda_device):
    """bfTrajectorySelectReintegration against the oracle on 5 000 frames: same list (well separated distances), dist within 1e-5
    relative + 1e-9 (sin / cos / asin of the two SE(3) logs differ in the last bits between CUDA and libm)."""
    import torch
    from bundlefusion_b200 import synth
    dev = cuda_device
    L = capi.lib()
    n, topN = 5000, 30
    rng = np.random.default_rng(3)
    integ = np.stack([synth.se3_exp(rng.standard_normal(3) * 0.4, rng.standard_normal(3) * 2) for _ in range(n)]).astype(F)
    opt = integ.copy()
    movers = rng.permutation(n)[:200]
    for r, k in enumerate(movers):
        opt[k] = (synth.se3_exp(rng.standard_normal(3) * 0.002 * (r + 1) / 20, rng.standard_normal(3) * 0.004 * (r + 1) / 20) @ integ[k].astype(np.float64)).astype(F)
    state = np.ones(n, np.int32); state[movers[-5:]] = 0
    opt[movers[-8], 0, 0] = -np.inf
    d_o, d_i, d_s = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (opt, integ, state))
    d_dist = torch.zeros(n, device=dev); d_list = torch.full((topN,), -1, dtype=torch.int32, device=dev); d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    L.bfSetStream(None); torch.cuda.synchronize()
    capi.check(L.bfTrajectorySelectReintegration(d_o.data_ptr(), d_i.data_ptr(), d_s.data_ptr(), n, topN, 0.0004, 2.0, d_dist.data_ptr(), d_list.data_ptr(), d_cnt.data_ptr()),
               "bfTrajectorySelectReintegration")
    torch.cuda.synchronize()
    od, ol = orc.select_reintegration(opt, integ, state, topN, 0.0004)
    cnt = int(d_cnt.item())
    assert cnt == len(ol) == topN
    gd = d_dist.cpu().numpy()
    np.testing.assert_allclose(gd, od, rtol=1e-4, atol=1e-9)
    # the selected frames are the same; the order may differ only between frames whose distances agree to 1e-4 relative
    gl = d_list.cpu().numpy()[:cnt]
    assert set(gl.tolist()) == set(ol.tolist())
    assert np.all(np.diff(gd[gl]) <= 0)
    assert gd[movers[-8]] == -1 and np.all(gd[movers[-5:]] == -1)
