"""Ray cast of the hashed TSDF on the GPU (csrc/raycast.cu behind include/bf_raycast.h) against oracle/raycast_oracle.c: interval images, depth, camera-space
positions, colours and normals bit for bit.  The model is fused by the library's bit-exact TSDF kernels (arithmetic="exact"), the oracle's by its own;
hash slots and list order differ between the two, voxel values do not -- and nothing else reaches the images."""
import ctypes as C

import numpy as np
import pytest

from bundlefusion_b200 import synth
from bundlefusion_b200.raycast import CUDARayCastSDF, ray_cast_params
from bundlefusion_b200.scene_rep import CUDASceneRepHashSDF, camera_params, default_hash_params
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def build(dev, W, H, n_frames, first, voxel=0.010, buckets=100003, blocks=90000):
    import torch
    cam = camera_params(W, H)
    hp = default_hash_params(num_buckets=buckets, num_sdf_blocks=blocks, voxel_size=voxel)
    gpu = CUDASceneRepHashSDF(hp, dev, arithmetic="exact")
    cpu = orc.OracleSceneRepHashSDF(hp)
    frames = [synth.make_frame(first + i, W, H) for i in range(n_frames)]
    for d, c, T in frames:
        gpu.integrate(T, torch.from_numpy(d).to(dev), torch.from_numpy(c).to(dev), cam)
        cpu.integrate(T, d, c, cam)
    return gpu, cpu, cam, frames


@pytest.mark.parametrize("W,H,grad", [(160, 120, False), (160, 120, True), (320, 240, False)])
def test_raycast_matches_oracle_bit_for_bit(cuda_device, W, H, grad):
    gpu, cpu, cam, frames = build(cuda_device, W, H, 3, 5)
    p = ray_cast_params(W, H, cam.fx, cam.fy, cam.mx, cam.my, use_gradients=grad)
    rc = CUDARayCastSDF(p, cuda_device)
    for T in (frames[2][2], frames[1][2]):                 # the last integrated pose, and an earlier one (part of the model outside the list's frustum)
        rc.render(gpu.getHashData(), gpu.getHashParams(), cam, T)
        got = rc.download()
        want = orc.raycast_frame(cpu, cam, p, T)
        for k in ("ray_min", "ray_max", "depth", "depth4", "colors", "normals"):
            assert np.array_equal(bits(got[k]), bits(want[k])), (k, int(np.count_nonzero(bits(got[k]) != bits(want[k]))))
        assert np.isfinite(got["depth"]).mean() > 0.9
    gpu.close()


def test_raycast_of_the_fast_arithmetic_model_and_of_nothing(cuda_device):
    """the default (tolerance-arithmetic) TSDF gives a model within 1e-5 m of the exact one: its ray cast stays within the march's resolution of the oracle's"""
    import torch
    W, H = 160, 120
    cam = camera_params(W, H)
    hp = default_hash_params(num_buckets=100003, num_sdf_blocks=90000)
    fast = CUDASceneRepHashSDF(hp, cuda_device, arithmetic="fast")
    cpu = orc.OracleSceneRepHashSDF(hp)
    frames = [synth.make_frame(20 + i, W, H) for i in range(3)]
    p = ray_cast_params(W, H, cam.fx, cam.fy, cam.mx, cam.my)
    rc = CUDARayCastSDF(p, cuda_device)
    rc.render(fast.getHashData(), fast.getHashParams(), cam, frames[0][2])          # empty model: nothing to see, nothing crashes
    assert np.all(np.isneginf(rc.download()["depth"]))
    for d, c, T in frames:
        fast.integrate(T, torch.from_numpy(d).to(cuda_device), torch.from_numpy(c).to(cuda_device), cam)
        cpu.integrate(T, d, c, cam)
    rc.render(fast.getHashData(), fast.getHashParams(), cam, frames[2][2])
    got, want = rc.download(), orc.raycast_frame(cpu, cam, p, frames[2][2])
    assert np.array_equal(bits(got["ray_min"]), bits(want["ray_min"])) and np.array_equal(bits(got["ray_max"]), bits(want["ray_max"]))      # the block set is identical
    both = np.isfinite(got["depth"]) & np.isfinite(want["depth"])
    assert both.mean() > 0.9 and np.mean(np.isfinite(got["depth"]) != np.isfinite(want["depth"])) < 2e-3
    assert np.percentile(np.abs(got["depth"][both] - want["depth"][both]), 99.9) < 1e-4
    fast.close()
