"""Times the surface-area and dense-verification match filters (csrc/sift_verify.cu) through the C-ABI on the GPU, torch-free
(CUDA runtime via ctypes), and the oracle on one host core beside them.  Prints one JSON line per configuration.

    python scripts/verify_filters_timing.py > gpurun_out/r1_verify_filters_timing.jsonl
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundlefusion_b200 import _capi as capi          # noqa: E402
from bundlefusion_b200 import synth                   # noqa: E402
from oracle import oracle as orc                      # noqa: E402
from tests._cudart import DevBuf, runtime             # noqa: E402
from tests.test_verify_filters_oracle import VERIFY   # noqa: E402


def f16(m):
    return np.ascontiguousarray(m, np.float32).reshape(16).ctypes.data_as(C.POINTER(C.c_float))


def gpu_time_us(fn, reps=50, warm=5):
    rt = runtime()
    rt.cudaEventCreate.argtypes = [C.POINTER(C.c_void_p)]
    rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
    rt.cudaEventSynchronize.argtypes = [C.c_void_p]
    rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    e0, e1 = C.c_void_p(), C.c_void_p()
    rt.cudaEventCreate(C.byref(e0)); rt.cudaEventCreate(C.byref(e1))
    for _ in range(warm):
        fn()
    rt.cudaDeviceSynchronize()
    rt.cudaEventRecord(e0, None)
    for _ in range(reps):
        fn()
    rt.cudaEventRecord(e1, None)
    rt.cudaEventSynchronize(e1)
    ms = C.c_float(0)
    rt.cudaEventElapsedTime(C.byref(ms), e0, e1)
    return 1000.0 * ms.value / reps


def main():
    L = capi.lib()
    rng = np.random.default_rng(0)
    W, H = 640, 480
    fx = 525.0; K = np.array([[fx, 0, 319.5, 0], [0, fx, 239.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    Kinv = np.linalg.inv(K).astype(np.float32)
    dv = synth.make_dense_verify_problem(n_prev=4)
    keep = []
    for P in (16, 128, 512):
        cur = P - 1
        n = 25
        keys = np.c_[rng.uniform(40, W - 40, 2 * P * n), rng.uniform(40, H - 40, 2 * P * n), np.ones(2 * P * n), rng.uniform(0.8, 3.0, 2 * P * n)].astype(np.float32)
        num = np.full(P, n, np.int32)
        fidx = np.zeros((P, 25, 2), np.uint32)
        for p in range(P):
            fidx[p, :, 0] = 2 * p * n + np.arange(n); fidx[p, :, 1] = (2 * p + 1) * n + np.arange(n)
        d_keys, d_idx, d_num0 = DevBuf(keys), DevBuf(fidx), DevBuf(num)
        d_num = DevBuf(num)

        def area():
            capi.check(L.bfSiftFilterMatchesBySurfaceArea(cur, 0, P, d_keys.ptr, d_num.ptr, d_idx.ptr, f16(Kinv), 0.0, None), "area")   # thresh 0: nothing is zeroed, every launch does full work
        us_area = gpu_time_us(area)
        t0 = time.perf_counter(); orc.sift_filter_surface_area(cur, 0, P, keys, num, fidx, Kinv, 0.0); cpu_area = 1e6 * (time.perf_counter() - t0)

        # dense verify: P pairs over 3 distinct cached frames (the pair's transform is the exact one: every pixel does full work)
        nfr = dv["P"]
        recs = (capi.BFCUDACachedFrame * P)()
        caches, T = [], np.zeros((P, 4, 4), np.float32)
        bufs = []
        for f in dv["caches"]:
            bufs.append((DevBuf(f["depth"]), DevBuf(f["campos"]), DevBuf(f["normals"])))
        for p in range(P):
            k = dv["cur"] if p == cur else 2 * (p % 2)                      # frames 0 and 2 carry their exact transform
            recs[p].d_depthDownsampled, recs[p].d_cameraposDownsampled, recs[p].d_normalsDownsampled = bufs[k][0].ptr, bufs[k][1].ptr, bufs[k][2].ptr
            caches.append(dv["caches"][k]); T[p] = dv["T"][k]
        d_recs, d_T = DevBuf(np.frombuffer(bytes(recs), np.uint8)), DevBuf(T)
        o = dict(VERIFY); o["errThresh"] = 1e9; o["corrThresh"] = -1.0     # nothing is zeroed between repetitions (NaN err still is: none here)

        def verify():
            capi.check(L.bfSiftFilterMatchesByDenseVerify(cur, 0, P, dv["W"], dv["H"], f16(dv["K"]), d_num.ptr, d_T.ptr, d_recs.ptr, o["distThresh"], o["normalThresh"],
                                                          o["colorThresh"], o["errThresh"], o["corrThresh"], o["dMin"], o["dMax"], None), "verify")
        us_verify = gpu_time_us(verify)
        still = int((d_num.get() != 0).sum())
        Pc = min(P, 64)
        t0 = time.perf_counter(); orc.sift_filter_dense_verify(Pc - 1, 0, Pc, dv["W"], dv["H"], dv["K"], num[:Pc], T[:Pc], caches[:Pc - 1] + [dv["caches"][dv["cur"]]], **o)
        cpu_verify = 1e6 * (time.perf_counter() - t0) * (P - 1) / max(Pc - 1, 1)
        # algorithmic bytes of the dense check per pair: both frames' depth + campos + normals read once (W H (4 + 16 + 16) B x 2), gathers hit L2
        bytes_pair = 2 * dv["W"] * dv["H"] * 36
        print(json.dumps({"pairs": P - 1, "surface_area_us": round(us_area, 2), "surface_area_us_per_pair": round(us_area / (P - 1), 4),
                          "dense_verify_us": round(us_verify, 2), "dense_verify_us_per_pair": round(us_verify / (P - 1), 4),
                          "dense_verify_GBps_algorithmic": round(bytes_pair * (P - 1) / us_verify / 1e3, 1), "pairs_still_valid": still,
                          "oracle_1core_surface_area_us": round(cpu_area, 1), "oracle_1core_dense_verify_us": round(cpu_verify, 1)}), flush=True)
        keep += [d_keys, d_idx, d_num0, d_num, d_recs, d_T, bufs]


if __name__ == "__main__":
    main()
