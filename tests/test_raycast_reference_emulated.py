"""Pins oracle/raycast_oracle.c against the REFERENCE's own ray-cast kernels (FL/DepthSensing/CUDARayCastSDF.cu: renderKernel with RayCastSDFUtil.h's
trilinear sampling / bisection / gradient, rayIntervalSplatKernel), executed on the CPU: oracle/build_ref.py (build_raycast_emulated) compiles the reference
sources where they lie against the CUDA emulation, scripts/make_golden_raycast_emulated.py runs them on the seeded scene below and commits the outputs as
tests/golden/raycast_reference_emulated.npz; this test replays the same scene through the oracle.  Depth, camera-space positions, colours, gradient normals
and the per-block quads: bit for bit.  (The Direct3D 11 rasterisation of the quads into the two interval images is not the reference's code and is not
pinned -- the golden file carries the interval images the reference kernel was given.)"""
import ctypes as C
import os

import numpy as np

from bundlefusion_b200 import synth
from bundlefusion_b200.raycast import ray_cast_params
from bundlefusion_b200.scene_rep import camera_params, default_hash_params
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raycast_reference_emulated.npz")
SW, SH = 160, 120            # frames the scene is fused from
VW, VH = 96, 72              # ray-cast view (other intrinsics than the sensor's: CUDARayCastSDF::parametersFromGlobalAppState adapts them, h:24-34)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def golden_scene():
    cam = camera_params(SW, SH)
    hp = default_hash_params(num_buckets=20011, num_sdf_blocks=30000)
    sc = orc.OracleSceneRepHashSDF(hp)
    frames = [synth.make_frame(40 + i, SW, SH) for i in range(3)]
    for d, c, T in frames:
        sc.integrate(T, d, c, cam)
    return sc, cam, frames


def view_params(cam, grad):
    sx, sy = VW / SW, VH / SH
    return ray_cast_params(VW, VH, cam.fx * sx, cam.fy * sy, cam.mx * (VW - 1) / (SW - 1), cam.my * (VH - 1) / (SH - 1), use_gradients=grad)


def view_pose(frames):
    T = np.array(frames[1][2], np.float32)
    T[:3, 3] += np.array([0.03, -0.02, 0.05], np.float32)                   # not a pose the model was fused from
    return T


def test_oracle_reproduces_the_reference_kernels_bit_for_bit():
    g = np.load(GOLDEN)
    sc, cam, frames = golden_scene()
    T = view_pose(frames)
    for grad in (0, 1):
        p = view_params(cam, bool(grad))
        orc.raycast_set_pose(p, T)
        rmin, rmax = g["ray_min"], g["ray_max"]
        o = orc.raycast_render(sc, p, rmin, rmax)
        assert np.isfinite(g[f"depth_g{grad}"]).mean() > 0.8
        for k in ("depth", "depth4", "colors"):
            assert np.array_equal(bits(o[k]), bits(g[f"{k}_g{grad}"])), (k, grad)
        if grad:
            assert np.array_equal(bits(o["normals"]), bits(g["normals_g1"]))
    # the quads of rayIntervalSplatKernel (six vertices per compactified entry; entries the reference skips stay at the fill value)
    p = view_params(cam, False)
    orc.raycast_set_pose(p, T)
    L = orc.lib()
    L.orc_raycast_block_quad.argtypes = [C.c_void_p] * 5
    n = int(sc.num_occupied)
    assert n == int(g["num_occupied"]) and n > 500
    for splat_min, key in ((1, "quads_min"), (0, "quads_max")):
        p.m_splatMinimum = splat_min
        vb = g[key].reshape(n, 6, 4)
        drawn = 0
        for e in range(n):
            q = np.zeros(6, np.float32)
            ent = np.ascontiguousarray(sc.compactified[e])
            if not L.orc_raycast_block_quad(C.addressof(sc.hp), C.addressof(cam), C.addressof(p), ent.ctypes.data, q.ctypes.data):
                assert np.all(vb[e] == 7.0), e
                continue
            drawn += 1
            v = vb[e]
            assert np.array_equal(bits(v[:, 2]), bits(np.full(6, q[4], np.float32))) and np.array_equal(bits(v[:, 3]), bits(np.full(6, q[5], np.float32))), e
            assert np.array_equal(bits(v[:, 0]), bits(np.array([q[2], q[0], q[2], q[0], q[2], q[0]], np.float32))), e
            assert np.array_equal(bits(v[:, 1]), bits(np.array([q[1], q[1], q[3], q[1], q[3], q[3]], np.float32))), e
        assert drawn > 0.9 * n
    # and the interval images the reference kernel was given are the oracle's rasterisation of those quads
    assert np.array_equal(bits(orc.raycast_splat(sc, cam, p, 1)), bits(g["ray_min"])) and np.array_equal(bits(orc.raycast_splat(sc, cam, p, 0)), bits(g["ray_max"]))
