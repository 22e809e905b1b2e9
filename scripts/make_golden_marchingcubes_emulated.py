"""Generates tests/golden/marchingcubes_reference_emulated.npz: the triangles of the REFERENCE's own iso-surface kernel (FL/DepthSensing/CUDAMarchingCubesSDF.cu:
extractIsoSurfaceKernel) executed on the CPU (oracle/_ref/libref_marchingcubes_emulated.so, built by oracle/build_ref.py build_marchingcubes_emulated against the
CUDA emulation) on the seeded scene of tests/test_marchingcubes_reference_emulated.py, and the reference's edge / triangle tables (FL/DepthSensing/Tables.h).

    python oracle/build_ref.py && python scripts/make_golden_marchingcubes_emulated.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_marchingcubes_reference_emulated import GOLDEN, golden_params, golden_scene, scene_box              # noqa: E402


def main():
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_marchingcubes_emulated.so"))
    sizes = (C.c_int * 5)()
    R.ref_marchingcubes_sizes(sizes)
    from bundlefusion_b200._capi import BFHashParams, BFMarchingCubesParams
    assert list(sizes) == [32, C.sizeof(BFHashParams), C.sizeof(BFMarchingCubesParams), 12, 72], list(sizes)      # the reference's struct layouts = this repo's
    vp = C.c_void_p
    R.ref_marchingcubes_extract.argtypes = [vp] * 6
    R.ref_marchingcubes_tables.argtypes = [vp, vp]
    sc, cam, frames = golden_scene()
    out = {"num_blocks": np.int32((sc.hash[:, 3] != -2).sum())}
    edge = np.zeros(256, np.int32); tri = np.zeros((256, 16), np.int32)
    R.ref_marchingcubes_tables(edge.ctypes.data, tri.ctypes.data)
    out["edge_table"], out["tri_table"] = edge, tri

    def run(p):
        buf = np.zeros((int(p.m_maxNumTriangles), 3, 6), np.float32)
        n = np.zeros(1, np.uint32)
        R.ref_marchingcubes_extract(sc.hash.ctypes.data, sc.voxels.ctypes.data, C.addressof(sc.hp), C.addressof(p), buf.ctypes.data, n.ctypes.data)
        return buf[:int(n[0])].copy()
    out["triangles"] = run(golden_params(sc.hp))
    out["triangles_box"] = run(golden_params(sc.hp, scene_box(sc)))
    out["num_capped"] = np.int32(len(run(golden_params(sc.hp, cap=1000))))
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes;", len(out["triangles"]), "triangles,", len(out["triangles_box"]), "in the box,", int(out["num_blocks"]), "blocks")


if __name__ == "__main__":
    main()
