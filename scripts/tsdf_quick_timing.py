"""Quick device timing of the TSDF frame pipeline (development aid; bench.py is the contract)."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundlefusion_b200 import synth
from bundlefusion_b200.scene_rep import CUDASceneRepHashSDF, camera_params, default_hash_params

dev = torch.device("cuda:0")
W, H = 640, 480
cam = camera_params(W, H)
for (nb, ns, vs) in [(800000, 200000, 0.010), (4000000, 1000000, 0.005)]:
    hp = default_hash_params(num_buckets=nb, num_sdf_blocks=ns, voxel_size=vs)
    sc = CUDASceneRepHashSDF(hp, dev)
    frames = [synth.make_frame(10 * i, W, H) for i in range(16)]
    devf = [(torch.from_numpy(f[0]).to(dev), torch.from_numpy(f[1]).to(dev), f[2]) for f in frames]
    for d, c, T in devf:
        sc.integrate(T, d, c, cam)
    torch.cuda.synchronize()
    st = sc.getLastFrameStats()
    # steady state: de-integrate + re-integrate the same frames (hash is warm)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    reps = 5
    ev[0].record()
    for _ in range(reps):
        for d, c, T in devf:
            sc.deIntegrate(T, d, c, cam)
            sc.integrate(T, d, c, cam)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / (reps * len(devf) * 2)
    st2 = sc.getLastFrameStats()
    print(json.dumps({"buckets": nb, "blocks": ns, "voxel": vs, "ms_per_pass": ms, "stats": st2, "heap_free": sc.getHeapFreeCount(),
                      "alg_MB": (24 * st2["U"] + 20 * st2["E"] + 2 * W * H * 4) / 1e6,
                      "GBps_alg": (24 * st2["U"] + 20 * st2["E"] + 2 * W * H * 4) / 1e9 / (ms / 1e3)}))
    sc.close()
    del sc
