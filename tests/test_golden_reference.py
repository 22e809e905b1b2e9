"""The CPU oracle against committed outputs of the REFERENCE'S OWN CUDA (tests/golden/*.npz, produced on a B200 by
scripts/make_golden_from_reference.py from oracle/_ref's IEEE build of CUDASceneRepHashSDF.cu / SolverBundling.cu): the same seeded inputs
are regenerated here and replayed through oracle/*.c.  This is the pin of the oracle that runs without a GPU.
TSDF: allocated block set, heap count and every voxel word (CRC per block) bit-identical to the reference's kernels.
Solver: poses within 1e-4 relative L2 (BASELINE north_star), dense overlap count equal."""
import os
import zlib

import numpy as np
import pytest

from bundlefusion_b200 import synth
from bundlefusion_b200.scene_rep import camera_params, default_hash_params
from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F = np.float32


def _load(name):
    p = os.path.join(GOLD, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated yet (scripts/make_golden_from_reference.py on a GPU box)")
    return np.load(p)


def test_tsdf_oracle_equals_reference_kernels():
    g = _load("tsdf_reference_ieee.npz")
    c = eval(bytes(g["case"]).decode())
    cam = camera_params(c["W"], c["H"])
    hp = default_hash_params(num_buckets=c["num_buckets"], num_sdf_blocks=c["num_sdf_blocks"])
    o = orc.OracleSceneRepHashSDF(hp)
    frames = [synth.make_frame(i, c["W"], c["H"]) for i in c["frames"]]
    for d, col, T in frames:
        o.integrate(T, d, col, cam)
    k = c["reint"]
    d, col, T = frames[k]
    T2 = T.copy(); T2[:3, 3] += np.array(c["shift"], F)
    o.deIntegrate(T, d, col, cam); o.integrate(T2, d, col, cam)
    o.garbageCollect()
    b, v = orc.canonical_blocks(o.download())
    np.testing.assert_array_equal(b, g["blocks"])
    crcs = np.array([zlib.crc32(np.ascontiguousarray(x).tobytes()) for x in v], np.uint32)
    np.testing.assert_array_equal(crcs, g["crcs"])
    np.testing.assert_array_equal(v[:4], g["first_voxels"])
    assert o.getHeapFreeCount() == int(g["heap_free"])


def test_solver_oracle_matches_reference_kernels():
    g = _load("solver_reference_ieee.npz")
    s = eval(bytes(g["sparse_case"]).decode())
    prob = synth.make_ba_problem(s["n_images"], degree=s["degree"], corr_per_pair=s["corr_per_pair"], noise=s["noise"], seed=s["seed"])
    o = orc.solve_sparse(prob["corr"], prob["init_rot"], prob["init_trans"], s["n_gn"], s["n_pcg"])
    x_o, x_r = np.c_[o["rot"], o["trans"]], np.c_[g["sparse_rot"], g["sparse_trans"]]
    assert np.linalg.norm(x_o - x_r) / np.linalg.norm(x_r) < 1e-4
    e = orc.energy(prob["corr"], o["rot"], o["trans"])
    assert abs(e - float(g["sparse_energy"][-1])) <= 0.02 * float(g["sparse_energy"][-1]) + 1e-7
    dc = eval(bytes(g["dense_case"]).decode())
    dp = synth.make_dense_ba_problem(dc["n_images"], stride=dc["stride"], W=dc["W"], H=dc["H"])
    o = orc.solve(dp["corr"], dp["init_rot"], dp["init_trans"], dc["n_gn"], dc["n_pcg"], [1.0] * dc["n_gn"], [1.0, 2.0], [0.0, 0.0], dp["caches"], dp["intrinsics"])
    x_o, x_r = np.c_[o["rot"], o["trans"]], np.c_[g["dense_rot"], g["dense_trans"]]
    assert np.linalg.norm(x_o - x_r) / np.linalg.norm(x_r) < 1e-4
    assert o["overlap_pairs"] == int(g["dense_overlap"])
