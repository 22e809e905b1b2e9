"""Times bfSiftDetect (csrc/sift_detect.cu) through the C-ABI on the GPU, torch-free (CUDA events via ctypes on the library's stream), and
the oracle on one host core beside it; checks the result against the oracle on the way.  One JSON line per configuration.

    python scripts/sift_detect_timing.py > gpurun_out/sift_detect_timing.jsonl
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundlefusion_b200 import _capi as capi                      # noqa: E402
from oracle import oracle as orc                                  # noqa: E402
from scripts.verify_filters_timing import gpu_time_us             # noqa: E402
from tests._cudart import DevBuf                                  # noqa: E402
from tests.test_zz_sift_detect_gpu import texture                 # noqa: E402


def main():
    L = capi.lib()
    for (H, W, opts) in ((480, 640, dict(minKeyScale=3.0, featureCountThreshold=150, maxKeyPoints=1024)),
                         (480, 640, dict(minKeyScale=0.0, featureCountThreshold=100000, maxKeyPoints=4096)),
                         (960, 1280, dict(minKeyScale=3.0, featureCountThreshold=150, maxKeyPoints=1024))):
        I = texture(7, H, W); D = np.full((H, W), 1.5, np.float32)
        P = capi.BFSiftDetectParams(W, H, W, H, 0.1, 3.0, opts["minKeyScale"], opts["featureCountThreshold"], opts["maxKeyPoints"])
        d_I, d_D = DevBuf(I), DevBuf(D)
        d_kp, d_des = DevBuf(np.zeros((opts["maxKeyPoints"], 4), np.float32)), DevBuf(np.zeros((opts["maxKeyPoints"], 128), np.uint8))
        d_n, d_lc = DevBuf(np.zeros(1, np.int32)), DevBuf(np.zeros(12, np.int32))

        def run():
            capi.check(L.bfSiftDetect(C.byref(P), d_I.ptr, d_D.ptr, d_kp.ptr, d_des.ptr, d_n.ptr, d_lc.ptr), "bfSiftDetect")
        us = gpu_time_us(run, reps=30, warm=3)
        n = int(d_n.get()[0])
        t0 = time.perf_counter(); ko, do, lo = orc.sift_detect(I, D, depthMin=0.1, depthMax=3.0, **opts); cpu_ms = 1e3 * (time.perf_counter() - t0)
        kg = d_kp.get()[:n]
        same_set = set(map(tuple, kg.tolist())) == set(map(tuple, ko.tolist()))
        # algorithmic bytes: every Gaussian / DoG / gradient level written once and read about three times (next level, DoG neighbours, samplers)
        px = sum((W >> o) * (H >> o) for o in range(4))
        print(json.dumps({"size": [W, H], "options": opts, "us_per_frame": round(us, 1), "launches_per_frame": 27, "key_points": n, "oracle_key_points": len(ko),
                          "same_key_point_set": bool(same_set), "level_counts": d_lc.get().tolist(), "oracle_1core_ms": round(cpu_ms, 1),
                          "pyramid_bytes": px * 4 * (6 + 5 + 2 * 3), "workspace_bytes": int(L.bfSiftDetectWorkspaceBytes())}), flush=True)


if __name__ == "__main__":
    main()
