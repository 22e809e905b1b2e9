"""Timing of the ray cast (bfRayCastRenderPose: interval splat + ray march + normals) on the GPU box: 640x480 view of a model fused from a synthetic stream,
CUDA events on the library's stream; the oracle on one host core beside it.  Prints one JSON line per case."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from bundlefusion_b200 import _capi as capi, synth_gpu  # noqa: E402
from bundlefusion_b200.raycast import CUDARayCastSDF, ray_cast_params  # noqa: E402
from bundlefusion_b200.scene_rep import CUDASceneRepHashSDF, camera_params, default_hash_params  # noqa: E402

dev = torch.device("cuda:0")
L = capi.lib()
for (W, H, voxel, n_frames) in ((640, 480, 0.010, 24), (640, 480, 0.005, 12), (320, 240, 0.010, 24)):
    cam = camera_params(W, H)
    hp = default_hash_params(num_buckets=2000003, num_sdf_blocks=1500000, voxel_size=voxel)
    sc = CUDASceneRepHashSDF(hp, dev)
    d, c, poses = synth_gpu.make_frames(list(range(0, 2 * n_frames, 2)), W, H, device=str(dev), texture="rich")
    for i in range(n_frames):
        sc.integrate(poses[i], d[i], c[i], cam)
    p = ray_cast_params(W, H, cam.fx, cam.fy, cam.mx, cam.my)
    rc = CUDARayCastSDF(p, dev)
    T = poses[n_frames - 1]
    for _ in range(5):
        rc.render(sc.getHashData(), sc.getHashParams(), cam, T)
    torch.cuda.synchronize()
    l0 = L.bfGetLaunchCount()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    a.record()
    for _ in range(reps):
        rc.render(sc.getHashData(), sc.getHashParams(), cam, T)
    b.record(); torch.cuda.synchronize()
    out = rc.download()
    hit = np.isfinite(out["depth"])
    print(json.dumps({"view": [W, H], "voxel_m": voxel, "frames_fused": n_frames, "occupied_blocks_in_list": sc.getNumOccupiedBlocks(), "us_per_render": round(a.elapsed_time(b) * 1e3 / reps, 1),
                      "launches_per_render": (L.bfGetLaunchCount() - l0) // reps, "pixels_hit_frac": round(float(hit.mean()), 4),
                      "mean_interval_m": round(float(np.mean((out["ray_max"] - out["ray_min"])[np.isfinite(out["ray_min"]) & np.isfinite(out["ray_max"])])), 3)}))
    sc.close()
