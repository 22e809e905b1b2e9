"""Generates tests/golden/raycast_reference_emulated.npz: outputs of the REFERENCE's own ray-cast kernels (FL/DepthSensing/CUDARayCastSDF.cu: renderKernel,
rayIntervalSplatKernel) executed on the CPU (oracle/_ref/libref_raycast_emulated.so, built by oracle/build_ref.py build_raycast_emulated against the CUDA
emulation) on the seeded scene of tests/test_raycast_reference_emulated.py.  The interval images handed to renderKernel come from the oracle's rasterisation
of the reference kernel's quads (the reference leaves that step to Direct3D 11) and are stored in the file.

    python oracle/build_ref.py && python scripts/make_golden_raycast_emulated.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc                                                                                  # noqa: E402
from tests.test_raycast_reference_emulated import GOLDEN, VH, VW, golden_scene, view_params, view_pose              # noqa: E402


def main():
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_raycast_emulated.so"))
    sizes = (C.c_int * 5)()
    R.ref_raycast_sizes(sizes)
    from bundlefusion_b200._capi import BFDepthCameraParams, BFHashParams, BFRayCastParams
    assert list(sizes) == [32, C.sizeof(BFHashParams), C.sizeof(BFRayCastParams), 12, C.sizeof(BFDepthCameraParams)], list(sizes)      # the reference's struct layouts = this repo's
    vp = C.c_void_p
    R.ref_raycast_render.argtypes = [vp] * 10
    R.ref_raycast_quads.argtypes = [vp] * 5
    sc, cam, frames = golden_scene()
    T = view_pose(frames)
    out = {"num_occupied": np.int32(sc.num_occupied)}
    p = view_params(cam, False)
    orc.raycast_set_pose(p, T)
    out["ray_min"], out["ray_max"] = orc.raycast_splat(sc, cam, p, 1), orc.raycast_splat(sc, cam, p, 0)
    n = int(sc.num_occupied)
    for splat_min, key in ((1, "quads_min"), (0, "quads_max")):
        p.m_splatMinimum = splat_min; p.m_numOccupiedSDFBlocks = n; p.m_maxNumVertices = 6 * n
        vb = np.full((6 * n, 4), 7.0, np.float32)
        R.ref_raycast_quads(sc.compactified.ctypes.data, C.addressof(sc.hp), C.addressof(cam), C.addressof(p), vb.ctypes.data)
        out[key] = vb
    for grad in (0, 1):
        p = view_params(cam, bool(grad))
        orc.raycast_set_pose(p, T)
        d = np.zeros((VH, VW), np.float32); d4 = np.zeros((VH, VW, 4), np.float32); nr = np.zeros((VH, VW, 4), np.float32); co = np.zeros((VH, VW, 4), np.float32)
        R.ref_raycast_render(sc.hash.ctypes.data, sc.voxels.ctypes.data, C.addressof(sc.hp), C.addressof(p), out["ray_min"].ctypes.data, out["ray_max"].ctypes.data,
                             d.ctypes.data, d4.ctypes.data, nr.ctypes.data, co.ctypes.data)
        out[f"depth_g{grad}"], out[f"depth4_g{grad}"], out[f"colors_g{grad}"] = d, d4, co
        if grad:
            out["normals_g1"] = nr
        print(f"useGradients={grad}: {int(np.isfinite(d).sum())} of {d.size} pixels hit")
    os.makedirs(os.path.dirname(GOLDEN), exist_ok=True)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
