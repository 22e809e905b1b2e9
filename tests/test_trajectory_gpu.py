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
