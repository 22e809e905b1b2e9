"""Pins oracle/marchingcubes_oracle.c against the REFERENCE's own iso-surface kernel (FL/DepthSensing/CUDAMarchingCubesSDF.cu: extractIsoSurfaceKernel with
MarchingCubesSDFUtil.h's extractIsoSurfaceAtPosition / vertexInterp and the tables of Tables.h), executed on the CPU: oracle/build_ref.py
(build_marchingcubes_emulated) compiles the reference sources where they lie against the CUDA emulation, scripts/make_golden_marchingcubes_emulated.py runs them
on the seeded scene below and commits the triangles as tests/golden/marchingcubes_reference_emulated.npz; this test replays the scene through the oracle.
The reference appends triangles with one atomicAdd each -- the ORDER of its soup is not defined -- so the statement is on the multiset of triangles: every
triangle with its three vertices in the reference's order, positions and colours bit for bit.  Both tables are compared entry by entry."""
import os

import numpy as np

from bundlefusion_b200 import synth
from bundlefusion_b200.marching_cubes import marching_cubes_params
from bundlefusion_b200.scene_rep import camera_params, default_hash_params
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "marchingcubes_reference_emulated.npz")
SW, SH = 80, 60              # frames the scene is fused from
BUCKETS = 499                # a small table: the reference's kernel runs one 512-thread CTA per hash slot (the emulation runs them one after the other)


def golden_scene(small=False):
    """small: a sixth of the blocks (the emulated product kernel spends most of its time in 512-thread barriers per block)"""
    cam = camera_params(SW, SH)
    hp = default_hash_params(num_buckets=BUCKETS, num_sdf_blocks=1500)
    sc = orc.OracleSceneRepHashSDF(hp)
    frames = []
    for i in range(2):
        d, c, T = synth.make_frame(40 + 2 * i, SW, SH)
        keep = np.zeros_like(d, bool); keep[20:40, 25:55] = True                # the middle of the frame: a few hundred blocks
        if small:
            keep[:, :] = False; keep[26:34, 34:46] = True
        frames.append((np.where(keep, d, -np.inf).astype(np.float32), c, T))
    for d, c, T in frames:
        sc.integrate(T, d, c, cam)
    assert sc.dropped == 0
    return sc, cam, frames


def golden_params(hp, box=None, cap=40000):
    p = marching_cubes_params(BUCKETS, voxel_size=float(hp.m_virtualVoxelSize), max_num_triangles=cap)
    if box is not None:
        p.m_boxEnabled = 1
        for k in range(3):
            p.m_minCorner[k], p.m_maxCorner[k] = box[0][k], box[1][k]
    return p


def scene_box(sc):
    """an axis-aligned box through the middle of the occupied blocks"""
    occ = sc.hash[sc.hash[:, 3] != -2][:, :3].astype(np.float32) * 8 * float(sc.hp.m_virtualVoxelSize)
    lo, hi = occ.min(0), occ.max(0)
    mid = 0.5 * (lo + hi)
    return (lo[0] - 1.0, lo[1] - 1.0, lo[2] - 1.0), (float(mid[0]), hi[1] + 1.0, hi[2] + 1.0)


def canon(tri):
    """the soup as a sorted list of 72-byte rows (order of the soup removed, order inside a triangle kept)"""
    rows = np.ascontiguousarray(tri, np.float32).reshape(len(tri), 18).view(np.uint32)
    return rows[np.lexsort(rows.T[::-1])]


def test_tables_equal_the_references():
    g = np.load(GOLDEN)
    edge, tri = orc.marchingcubes_tables()
    assert np.array_equal(edge, g["edge_table"]) and np.array_equal(tri, g["tri_table"])


def test_oracle_reproduces_the_reference_kernel_bit_for_bit():
    g = np.load(GOLDEN)
    sc, cam, frames = golden_scene()
    assert int((sc.hash[:, 3] != -2).sum()) == int(g["num_blocks"]) > 300
    tri, found = orc.marchingcubes_extract(sc, golden_params(sc.hp))
    assert found == len(tri) == len(g["triangles"]) > 10000
    assert np.array_equal(canon(tri), canon(g["triangles"]))
    # colours are the cell's voxel colour / 255, positions lie within half a voxel of a voxel centre of an occupied block
    assert tri[..., 3:].min() >= 0.0 and tri[..., 3:].max() <= 1.0
    # the box: only cells whose centre is inside
    box = scene_box(sc)
    tb, fb = orc.marchingcubes_extract(sc, golden_params(sc.hp, box))
    assert 0 < len(tb) < len(tri) and np.array_equal(canon(tb), canon(g["triangles_box"]))
    # a full buffer: the count stops at the capacity (which triangles survive depends on the order; the reference's is undefined)
    tc, fc = orc.marchingcubes_extract(sc, golden_params(sc.hp, cap=1000))
    assert len(tc) == 1000 == int(g["num_capped"]) and fc == found
