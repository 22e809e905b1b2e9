"""Parity of the bundle-adjustment solver against the REFERENCE'S OWN CUDA solver (oracle/_ref/libref_solver*.so:
FL/Solver/SolverBundling.cu + FL/SBA.cu built for sm_100a with the compatibility patch of oracle/build_ref.py), on identical
correspondences, cache frames and initial poses, on the GPU.  This is what pins the CPU oracle (oracle/solver_oracle.c) too: the
same cases are run through it and compared with the reference's output.

Tolerance (BASELINE.json north_star): solved poses within 1e-4 relative L2.  The reference accumulates with float atomics, so its
own result moves in the last bits from run to run; the PCG early-out |p.Ap| < 5e-7 is absolute (see tests/test_solver_gpu.py)."""
import ctypes as C

import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from bundlefusion_b200 import synth
from bundlefusion_b200.solver import CUDASolverBundling, DeviceCache
from oracle import oracle as orc
from oracle import ref_solver

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12)


def _dev_inputs(dev, prob, corr):
    import torch
    c = np.ascontiguousarray(corr)
    corr_t = torch.from_numpy(c.view(np.uint8).reshape(-1).copy()).to(dev) if len(c) else torch.zeros(32, dtype=torch.uint8, device=dev)
    rot = torch.from_numpy(prob["init_rot"].copy()).to(dev)
    trans = torch.from_numpy(prob["init_trans"].copy()).to(dev)
    valid = torch.ones(len(prob["init_rot"]), dtype=torch.int32, device=dev)
    return corr_t, rot, trans, valid


def run_ours(dev, prob, corr, n_gn, n_pcg, wS, wD=None, wC=None, cache=None):
    import torch
    N = len(prob["init_rot"])
    corr_t, rot, trans, valid = _dev_inputs(dev, prob, corr)
    s = CUDASolverBundling(N, max(len(corr), 1000 * N), dev)
    s.solve(corr_t, len(corr), valid, N, n_gn, n_pcg, wS, wD, wC, d_rotationAnglesUnknowns=rot, d_translationUnknowns=trans, cudaCache=cache)
    torch.cuda.synchronize()
    return np.c_[rot.cpu().numpy(), trans.cpu().numpy()], s.getStats()


def run_ref(dev, prob, corr, n_gn, n_pcg, wS, wD=None, wC=None, cache=None, fast=True):
    N = len(prob["init_rot"])
    corr_t, rot, trans, valid = _dev_inputs(dev, prob, corr)
    s = ref_solver.ReferenceSolverBundling(N, max(len(corr), 1000 * N), dev, fast_math=fast)
    conv = s.solve(corr_t, len(corr), valid, N, n_gn, n_pcg, wS, wD, wC, d_rot=rot, d_trans=trans, cudaCache=cache,
                   record_convergence=len(corr) > 0)    # EvalResidual launches a 0-block grid when there are no correspondences
    return np.c_[rot.cpu().numpy(), trans.cpu().numpy()], conv, s


@pytest.fixture(autouse=True)
def _need_ref():
    if not ref_solver.available(True) or not ref_solver.available(False):
        pytest.skip("oracle/_ref/libref_solver*.so not built (needs /root/reference at build time)")


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("n_images,degree,n_gn,n_pcg", [(2, 1, 4, 50), (11, 10, 2, 100), (60, 8, 3, 150), (200, 12, 4, 150)])
def test_sparse_solve_matches_reference_cuda(cuda_device, n_images, degree, n_gn, n_pcg, fast):
    cpp = 256 if n_images == 2 else 25
    prob = synth.make_ba_problem(n_images, degree=degree, corr_per_pair=cpp, noise=0.002, seed=5)
    w = [1.0] * n_gn
    x_ref, conv, _ = run_ref(cuda_device, prob, prob["corr"], n_gn, n_pcg, w, fast=fast)
    x_our, st = run_ours(cuda_device, prob, prob["corr"], n_gn, n_pcg, w)
    o = orc.solve_sparse(prob["corr"], prob["init_rot"], prob["init_trans"], n_gn, n_pcg)
    x_orc = np.c_[o["rot"], o["trans"]]
    assert st["error"] == 0
    assert np.isfinite(x_ref).all()
    assert rel_l2(x_our, x_ref) < TOL, "CUDA path vs reference CUDA"
    assert rel_l2(x_orc, x_ref) < TOL, "CPU oracle vs reference CUDA (pins the oracle)"
    # the reference's own energy record (EvalResidual after each GN iteration) agrees with the oracle's energy of our result
    e_ref_final = float(conv[conv >= 0][-1])
    e_our = orc.energy(prob["corr"], x_our[:, :3].astype(np.float32), x_our[:, 3:].astype(np.float32))
    assert e_our <= e_ref_final * 1.02 + 1e-7
    np.testing.assert_array_equal(x_ref[0], x_our[0])                       # variable 0 is fixed in both


def test_dense_only_matches_reference_cuda(cuda_device):
    prob = synth.make_dense_ba_problem(5, stride=3, perturb_rot=0.004, perturb_trans=0.008, W=320, H=240)
    cache = DeviceCache(prob["caches"], prob["intrinsics"], cuda_device)
    wS, wD, wC = [0.0] * 3, [1.0, 2.0, 3.0], [0.0] * 3
    empty = prob["corr"][:0]
    x_ref, _, s = run_ref(cuda_device, prob, empty, 3, 60, wS, wD, wC, cache=cache, fast=False)
    x_our, st = run_ours(cuda_device, prob, empty, 3, 60, wS, wD, wC, cache=cache)
    o = orc.solve(empty, prob["init_rot"], prob["init_trans"], 3, 60, wS, wD, wC, prob["caches"], prob["intrinsics"])
    n_overlap_ref = int(s._bufs["d_numDenseOverlappingImages"].cpu().numpy()[0])
    assert st["dense_overlap_pairs"] == n_overlap_ref == o["overlap_pairs"]
    assert rel_l2(x_our, x_ref) < TOL
    assert rel_l2(np.c_[o["rot"], o["trans"]], x_ref) < TOL


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("wC", [[0.0, 0.0], [0.1, 0.1]])
def test_local_chunk_sparse_plus_dense_matches_reference_cuda(cuda_device, wC, fast):
    """The reference's local BA configuration (FL/SBA.cpp:28-31): 11 frames, 2 GN x 100 PCG, sparse 1, dense depth 1 -> 2."""
    prob = synth.make_dense_ba_problem(11, stride=3, W=320, H=240)
    cache = DeviceCache(prob["caches"], prob["intrinsics"], cuda_device)
    wS, wD = [1.0, 1.0], [1.0, 2.0]
    x_ref, _, s = run_ref(cuda_device, prob, prob["corr"], 2, 100, wS, wD, wC, cache=cache, fast=fast)
    x_our, st = run_ours(cuda_device, prob, prob["corr"], 2, 100, wS, wD, wC, cache=cache)
    o = orc.solve(prob["corr"], prob["init_rot"], prob["init_trans"], 2, 100, wS, wD, wC, prob["caches"], prob["intrinsics"])
    assert st["dense_overlap_pairs"] == int(s._bufs["d_numDenseOverlappingImages"].cpu().numpy()[0])
    assert rel_l2(x_our, x_ref) < TOL
    assert rel_l2(np.c_[o["rot"], o["trans"]], x_ref) < TOL


def test_dense_system_matches_reference_cuda(cuda_device):
    """The assembled dense normal equations themselves: the reference's d_denseJtJ (6N x 6N) / d_denseJtr after BuildDenseSystem
    against the oracle's, entry for entry (float-atomic summation order differs: relative Frobenius 1e-4)."""
    prob = synth.make_dense_ba_problem(6, stride=3, W=320, H=240)
    cache = DeviceCache(prob["caches"], prob["intrinsics"], cuda_device)
    N = 6
    # one GN iteration with zero PCG iterations leaves the poses untouched and the system of iteration 0 in the buffers
    x_ref, _, s = run_ref(cuda_device, prob, prob["corr"][:0], 1, 0, [0.0], [1.0], [0.1], cache=cache, fast=False)
    np.testing.assert_array_equal(x_ref, np.c_[prob["init_rot"], prob["init_trans"]])
    JtJ_ref = s._bufs["d_denseJtJ"].cpu().numpy().reshape(6 * N, 6 * N)
    Jtr_ref = s._bufs["d_denseJtr"].cpu().numpy()
    JtJ, Jtr, _ = orc.build_dense(prob["init_rot"], prob["init_trans"], prob["caches"], prob["intrinsics"], 1.0, 0.1)
    assert np.linalg.norm(JtJ_ref) > 0
    assert rel_l2(JtJ, JtJ_ref) < TOL
    assert rel_l2(Jtr, Jtr_ref) < TOL


def test_max_residual_and_pose_stubs_match_reference_cuda(cuda_device):
    import torch
    dev = cuda_device
    prob = synth.make_ba_problem(8, degree=7, corr_per_pair=25, noise=0.0, perturb_rot=0.0, perturb_trans=0.0)
    prob["corr"]["pj"][333] += np.array([0.0, 0.4, 0.0], np.float32)
    x_ref, _, s = run_ref(dev, prob, prob["corr"], 1, 1, [1.0], fast=False)
    v_ref, i_ref = s.max_residual()
    v_orc, i_orc = orc.max_residual(prob["corr"], x_ref[:, :3].astype(np.float32).copy(), x_ref[:, 3:].astype(np.float32).copy())
    assert i_ref == i_orc and abs(v_ref - v_orc) < 1e-5
    # pose <-> matrix stubs, ours against the reference's, on poses with small, moderate and near-pi rotations
    rng = np.random.default_rng(2)
    N = 64
    rot = (rng.standard_normal((N, 3)) * np.r_[np.full(16, 1e-4), np.full(32, 0.5), np.full(16, 1.6)][:, None]).astype(np.float32)
    trans = rng.standard_normal((N, 3)).astype(np.float32)
    outs = []
    L_our = capi.lib()
    L_our.bfSetStream(None)
    for L in (s.L, L_our):
        r = torch.from_numpy(rot).to(dev); t = torch.from_numpy(trans).to(dev)
        T = torch.zeros(N * 16, device=dev); Ti = torch.zeros(N * 16, device=dev); T2 = torch.zeros(N * 16, device=dev)
        r2 = torch.zeros_like(r); t2 = torch.zeros_like(t); valid = torch.ones(N, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        P = C.c_void_p
        L.convertLiePosesToMatricesCU(P(r.data_ptr()), P(t.data_ptr()), C.c_uint(N), P(T.data_ptr()), P(Ti.data_ptr()))
        L.convertMatricesToPosesCU(P(T.data_ptr()), C.c_uint(N), P(r2.data_ptr()), P(t2.data_ptr()), P(valid.data_ptr()))
        L.convertPosesToMatricesCU(P(r2.data_ptr()), P(t2.data_ptr()), C.c_uint(N), P(T2.data_ptr()), P(valid.data_ptr()))
        torch.cuda.synchronize()
        outs.append([x.cpu().numpy() for x in (T, Ti, r2, t2, T2)])
    for a, b, tol in zip(outs[0], outs[1], (2e-6, 1e-5, 2e-4, 2e-4, 2e-4)):
        np.testing.assert_allclose(b, a, atol=tol)
