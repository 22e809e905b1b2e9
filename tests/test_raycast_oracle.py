"""Ray cast of the hashed TSDF: known-answer tests of oracle/raycast_oracle.c, and the CUDA source (csrc/raycast.cu) executed on
the CPU through the emulation of tests/cuda_emu against the oracle, bit for bit.  The reference ships no test or golden vector for this path; how the oracle is
pinned against the reference's own kernel is in tests/test_raycast_reference_emulated.py."""
import ctypes as C
import os

import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from bundlefusion_b200 import synth
from bundlefusion_b200.raycast import ray_cast_params
from bundlefusion_b200.scene_rep import camera_params, default_hash_params
from oracle import oracle as orc

W, H = 160, 120


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def scene(n_frames=3, first=5, voxel=0.010, buckets=20011, blocks=30000):
    cam = camera_params(W, H)
    hp = default_hash_params(num_buckets=buckets, num_sdf_blocks=blocks, voxel_size=voxel)
    sc = orc.OracleSceneRepHashSDF(hp)
    frames = [synth.make_frame(first + i, W, H) for i in range(n_frames)]
    for d, c, T in frames:
        sc.integrate(T, d, c, cam)
    return sc, cam, frames


def test_raycast_recovers_the_integrated_surface():
    sc, cam, frames = scene()
    p = ray_cast_params(W, H, cam.fx, cam.fy, cam.mx, cam.my)
    o = orc.raycast_frame(sc, cam, p, frames[1][2])
    d, src = o["depth"], frames[1][0]
    hit = np.isfinite(d)
    assert hit.mean() > 0.97
    both = hit & np.isfinite(src)
    err = np.abs(d[both] - src[both])
    assert np.median(err) < 0.004 and np.percentile(err, 95) < 0.012           # sensor noise 1.2 mm z^2 + one-voxel interpolation
    # positions are the depth back-projected through the ray-cast intrinsics, colours are bytes / 255
    x, y = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    assert np.allclose(o["depth4"][..., 2][hit], d[hit]) and np.allclose(o["depth4"][..., 0][hit], (d * (x - cam.mx) / cam.fx)[hit], atol=1e-6)
    col = o["colors"][hit]
    assert col.min() >= 0.0 and col.max() <= 1.0 and np.allclose(col[:, :3] * 255.0, np.round(col[:, :3] * 255.0), atol=1e-4) and np.all(col[:, 3] == 1.0)
    assert np.all(np.isneginf(o["depth4"][~hit])) and np.all(np.isneginf(o["colors"][~hit]))
    # normals of the rendered positions: unit length, facing the camera (z component negative in the reference's convention: n = -(a x b) / |a x b|)
    n = o["normals"]; nv = np.isfinite(n[..., 0])
    assert nv.sum() > 0.9 * hit.sum() and np.allclose(np.linalg.norm(n[nv][:, :3], axis=1), 1.0, atol=1e-5)
    assert abs(np.mean(n[nv][:, 2])) > 0.9


def test_interval_images_bracket_the_surface_and_gate_the_march():
    sc, cam, frames = scene()
    p = ray_cast_params(W, H, cam.fx, cam.fy, cam.mx, cam.my)
    o = orc.raycast_frame(sc, cam, p, frames[2][2])
    d = o["depth"]; hit = np.isfinite(d)
    rmin, rmax = o["ray_min"], o["ray_max"]
    assert np.all(rmin[hit] <= d[hit] + 1e-3) and np.all(rmax[hit] >= d[hit] - 1e-3) and np.all(rmin[hit] <= rmax[hit])
    assert np.all(rmin[np.isfinite(rmin)] >= p.m_minDepth - 1e-6) and np.all(rmax[np.isfinite(rmax)] <= p.m_maxDepth + 1e-6)
    # a pixel without an interval is not marched; "0" counts as no interval too (CUDARayCastSDF.cu:43-44)
    rmin2 = rmin.copy(); rmin2[:, : W // 2] = -np.inf
    rmax2 = rmax.copy(); rmax2[: H // 2, W // 2:] = 0.0
    o2 = orc.raycast_render(sc, p, rmin2, rmax2)
    assert np.all(np.isneginf(o2["depth"][:, : W // 2])) and np.all(np.isneginf(o2["depth"][: H // 2, W // 2:]))
    assert np.array_equal(bits(o2["depth"][H // 2:, W // 2:]), bits(d[H // 2:, W // 2:]))
    # the whole depth range as the interval finds the same surface within the march's own resolution, but not the same bits: the start of the ray
    # fixes the sample positions -- which is why the interval images are part of the parity surface
    full = orc.raycast_render(sc, p, np.full((H, W), p.m_minDepth, np.float32), np.full((H, W), p.m_maxDepth, np.float32))
    b = hit & np.isfinite(full["depth"])
    assert b.mean() > 0.95 and np.median(np.abs(full["depth"][b] - d[b])) < 2e-3


def test_gradient_normals_and_empty_scene():
    sc, cam, frames = scene()
    p = ray_cast_params(W, H, cam.fx, cam.fy, cam.mx, cam.my, use_gradients=True)
    o = orc.raycast_frame(sc, cam, p, frames[0][2])
    hit = np.isfinite(o["depth"])
    n = o["normals"][hit]
    assert np.all(np.isfinite(n)) and np.all(n[:, 3] == 1.0)
    ln = np.linalg.norm(n[:, :3], axis=1)
    assert np.all((np.abs(ln - 1.0) < 1e-4) | (ln == 0.0))
    # gradient normals agree with the finite-difference ones where both exist
    p2 = ray_cast_params(W, H, cam.fx, cam.fy, cam.mx, cam.my)
    o2 = orc.raycast_frame(sc, cam, p2, frames[0][2])
    both = hit & np.isfinite(o2["normals"][..., 0]) & (np.linalg.norm(o["normals"][..., :3], axis=2) > 0.5)
    cosang = np.abs(np.sum(o["normals"][both][:, :3] * o2["normals"][both][:, :3], axis=1))
    assert np.median(cosang) > 0.97
    empty = orc.OracleSceneRepHashSDF(default_hash_params(num_buckets=1009, num_sdf_blocks=100))
    oe = orc.raycast_frame(empty, cam, p2, np.eye(4, dtype=np.float32))
    assert np.all(np.isneginf(oe["depth"])) and np.all(np.isneginf(oe["ray_min"])) and np.all(np.isneginf(oe["ray_max"]))


# ---- the CUDA source on the CPU ---------------------------------------------------------------------------------------------------------
_EXTRA = r'''
#include "%s"
#include "%s"
#include "%s"
namespace bf { static BFHashParams g_emuHp; static BFDepthCameraParams g_emuCp;
const BFHashParams* bound_hash_params() { return &g_emuHp; } const BFDepthCameraParams* bound_camera_params() { return &g_emuCp; } }
extern "C" void updateConstantHashParams(const BFHashParams* p) { bf::g_emuHp = *p; }
extern "C" void updateConstantDepthCameraParams(const BFDepthCameraParams* p) { bf::g_emuCp = *p; }
extern "C" void bfMat4Inverse(const float* m, float* o) { bf::mat4_inverse_ref(m, o); }
'''


@pytest.fixture(scope="module")
def emu():
    from tests.cuda_emu import build_emulated
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    extra = _EXTRA % (os.path.join(root, "include", "bf_tsdf.h"), os.path.join(root, "include", "bf_raycast.h"), os.path.join(root, "bundlefusion_b200", "csrc", "mat4.cuh"))
    L = build_emulated("raycast.cu", 6, extra_pre=extra)
    vp = C.c_void_p
    L.bfRayCastRenderPose.argtypes = [vp] * 6
    L.bfRayCastSplat.argtypes = [vp] * 5
    L.bfRayCastRender.argtypes = [vp] * 4
    L.rayIntervalSplatCUDA.argtypes = [vp] * 3; L.rayIntervalSplatCUDA.restype = None
    L.resetRayIntervalSplatCUDA.argtypes = [vp] * 2; L.resetRayIntervalSplatCUDA.restype = None
    L.renderCS.argtypes = [vp] * 3; L.renderCS.restype = None
    L.updateConstantRayCastParams.argtypes = [vp]; L.updateConstantHashParams.argtypes = [vp]; L.updateConstantDepthCameraParams.argtypes = [vp]
    return L


def run_emulated(L, sc, cam, p, T, use_stubs=False):
    Hh, Ww = p.m_height, p.m_width
    b = {"depth": np.zeros((Hh, Ww), np.float32), "depth4": np.zeros((Hh, Ww, 4), np.float32), "normals": np.zeros((Hh, Ww, 4), np.float32),
         "colors": np.zeros((Hh, Ww, 4), np.float32), "ray_min": np.zeros((Hh, Ww), np.float32), "ray_max": np.zeros((Hh, Ww), np.float32)}
    d = capi.BFRayCastData(b["depth"].ctypes.data, b["depth4"].ctypes.data, b["normals"].ctypes.data, b["colors"].ctypes.data, None, b["ray_min"].ctypes.data, b["ray_max"].ctypes.data)
    sc.compactified_counter[0] = sc.num_occupied          # the device-side count of the last compactify
    Tf = np.ascontiguousarray(T, np.float32).reshape(16)
    if not use_stubs:
        assert L.bfRayCastRenderPose(C.addressof(sc.hd), C.addressof(sc.hp), C.addressof(cam), C.addressof(d), C.addressof(p), Tf.ctypes.data) == 0
    else:                                                  # the reference's call sequence: constants, splat through the extension, renderCS
        orc.raycast_set_pose(p, T)
        L.updateConstantHashParams(C.addressof(sc.hp)); L.updateConstantDepthCameraParams(C.addressof(cam)); L.updateConstantRayCastParams(C.addressof(p))
        assert L.bfRayCastSplat(C.addressof(sc.hd), C.addressof(sc.hp), C.addressof(cam), C.addressof(d), C.addressof(p)) == 0
        L.renderCS(C.addressof(sc.hd), C.addressof(d), C.addressof(p))
    return b


@pytest.mark.parametrize("grad", [False, True])
def test_emulated_raycast_matches_oracle_bit_for_bit(emu, grad):
    sc, cam, frames = scene(n_frames=2, first=11)
    p = ray_cast_params(W, H, cam.fx, cam.fy, cam.mx, cam.my, use_gradients=grad)
    T = frames[1][2]
    want = orc.raycast_frame(sc, cam, p, T)
    got = run_emulated(emu, sc, cam, p, T)
    for k in ("ray_min", "ray_max", "depth", "depth4", "colors", "normals"):
        assert np.array_equal(bits(got[k]), bits(want[k])), k
    assert np.isfinite(got["depth"]).mean() > 0.9


def test_emulated_stubs_and_vertex_buffer(emu):
    sc, cam, frames = scene(n_frames=2, first=3)
    p = ray_cast_params(W, H, cam.fx, cam.fy, cam.mx, cam.my)
    T = frames[0][2]
    want = orc.raycast_frame(sc, cam, p, T)
    got = run_emulated(emu, sc, cam, p, T, use_stubs=True)
    for k in ("ray_min", "ray_max", "depth", "depth4", "colors"):
        assert np.array_equal(bits(got[k]), bits(want[k])), k
    # rayIntervalSplatCUDA: six vertices per entry, the reference's triangle order (CUDARayCastSDF.cu:160-169)
    n = int(sc.num_occupied)
    vb = np.full((n * 6, 4), 7.0, np.float32)
    d = capi.BFRayCastData(None, None, None, None, vb.ctypes.data, None, None)
    p.m_numOccupiedSDFBlocks = n; p.m_maxNumVertices = n * 6
    emu.resetRayIntervalSplatCUDA(C.addressof(d), C.addressof(p))
    assert np.all(np.isneginf(vb))
    for splat_min in (1, 0):
        p.m_splatMinimum = splat_min
        emu.rayIntervalSplatCUDA(C.addressof(sc.hd), C.addressof(d), C.addressof(p))
        L = orc.lib()
        L.orc_raycast_block_quad.argtypes = [C.c_void_p] * 5
        for e in (0, n // 2, n - 1):
            q = np.zeros(6, np.float32)
            ent = np.ascontiguousarray(sc.compactified[e])
            if not L.orc_raycast_block_quad(C.addressof(sc.hp), C.addressof(cam), C.addressof(p), ent.ctypes.data, q.ctypes.data):
                continue
            v = vb[6 * e:6 * e + 6]
            assert np.array_equal(bits(v[:, 2]), bits(np.full(6, q[4], np.float32))) and np.array_equal(bits(v[:, 3]), bits(np.full(6, q[5], np.float32)))
            assert np.array_equal(bits(v[:, 0]), bits(np.array([q[2], q[0], q[2], q[0], q[2], q[0]], np.float32)))
            assert np.array_equal(bits(v[:, 1]), bits(np.array([q[1], q[1], q[3], q[1], q[3], q[3]], np.float32)))
