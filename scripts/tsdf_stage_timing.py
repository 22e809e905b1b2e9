"""Warm (steady-state, L2 as it is in the real loop) CUDA-event timing of the TSDF stages, via the C-ABI stubs.
Development aid; bench.py is the contract.  usage: tsdf_stage_timing.py [voxel_size] [num_blocks] [num_buckets]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundlefusion_b200 import synth, _capi as capi
from bundlefusion_b200.scene_rep import CUDASceneRepHashSDF, camera_params, default_hash_params, set_pose

vs = float(sys.argv[1]) if len(sys.argv) > 1 else 0.010
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 800000
dev = torch.device("cuda:0")
W, H = 640, 480
cam = camera_params(W, H)
hp = default_hash_params(num_buckets=nb, num_sdf_blocks=ns, voxel_size=vs)
sc = CUDASceneRepHashSDF(hp, dev)
frames = [synth.make_frame(10 * i, W, H) for i in range(16)]
devf = [(torch.from_numpy(f[0]).to(dev), torch.from_numpy(f[1]).to(dev), f[2]) for f in frames]
for d, c, T in devf:
    sc.integrate(T, d, c, cam)
torch.cuda.synchronize()
L, hd, p = sc.lib, sc.m_hashData, sc.m_hashParams

def timed(fn, reps=40, warm=5):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us

out = {"voxel": vs}
k = [0]
def cur():
    k[0] = (k[0] + 1) % len(devf); return devf[k[0]]
def latch(d, c, T):
    set_pose(p, T)
    dd = capi.BFDepthCameraData(); dd.d_depthData = d.data_ptr(); dd.d_colorData = c.data_ptr()
    L.updateConstantHashParams(C.byref(p)); L.updateConstantDepthCameraParams(C.byref(cam)); L.bindInputDepthColorTextures(C.byref(dd), W, H)
    return dd
def f_alloc():
    d, c, T = cur(); dd = latch(d, c, T)
    L.allocCUDA(C.byref(hd), C.byref(p), C.byref(dd), C.byref(cam), None)
out["alloc_us"] = timed(f_alloc)
d, c, T = devf[3]; dd = latch(d, c, T)
n = L.compactifyHashAllInOneCUDA(C.byref(hd), C.byref(p)); p.m_numOccupiedBlocks = n
out["E"] = n
def f_int(): L.integrateDepthMapCUDA(C.byref(hd), C.byref(p), C.byref(dd), C.byref(cam))
def f_deint(): L.deIntegrateDepthMapCUDA(C.byref(hd), C.byref(p), C.byref(dd), C.byref(cam))
out["integrate_us"] = timed(f_int)
out["deintegrate_us"] = timed(f_deint, reps=40, warm=0)      # undo 40 of the 45 integrations
st = sc.getLastFrameStats(); out["U"] = st["U"] // 40
# whole passes replayed by the C-side op loop (no Python between launches): 16 frames per call
dl, cl = [f[0] for f in devf], [f[1] for f in devf]
ops_i = [(capi.BF_TSDF_OP_INTEGRATE, i, devf[i][2]) for i in range(16)]
ops_d = [(capi.BF_TSDF_OP_DEINTEGRATE, i, devf[i][2]) for i in range(16)]
ops_g = [(capi.BF_TSDF_OP_GARBAGE_COLLECT, 0, None)] * 16
out["pass_integrate_us"] = timed(lambda: sc.runOps(ops_i, dl, cl, cam), reps=4, warm=0) / 16
out["pass_deintegrate_us"] = timed(lambda: sc.runOps(ops_d, dl, cl, cam), reps=4, warm=0) / 16
out["gc_us"] = timed(lambda: sc.runOps(ops_g, dl, cl, cam), reps=4, warm=1) / 16
out["alg_MB"] = (24 * out["U"] + 20 * out["E"] + 2 * W * H * 4) / 1e6
out["integrate_GBps_alg"] = out["alg_MB"] / out["integrate_us"] * 1e3 / 1e3
print(json.dumps(out))
