"""CUDA-event timing of the batched SIFT descriptor matcher: one frame (1024 keys) against K earlier frames (development aid)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundlefusion_b200 import synth
from bundlefusion_b200.sift import ImagePairMatch, SiftMatchGPU
dev = torch.device("cuda:0")
cur = torch.from_numpy(synth.make_sift_descriptors(1024, seed=1)).to(dev)
m = SiftMatchGPU(device=dev)
for K in (1, 16, 128, 500):
    prevs = [torch.from_numpy(synth.make_sift_descriptors(1024, seed=10 + (k % 8))).to(dev) for k in range(min(K, 8))]
    ipms = [ImagePairMatch(dev) for _ in range(K)]
    jobs = [(prevs[k % len(prevs)], 1024, cur, 1024, ipms[k], (0, 0)) for k in range(K)]
    # the job table is built once (a C++ host builds it in microseconds; doing it in Python per call would be what gets timed)
    from bundlefusion_b200._capi import BFSiftMatchJob
    arr = (BFSiftMatchJob * K)(*[m._job(*p) for p in jobs])
    m._bind_stream()
    run = lambda: m.lib.bfSiftMatchBatch(arr, K, 0.7, 0.8)
    for _ in range(3): run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50 if K <= 16 else 10
    a.record()
    for _ in range(reps): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    macs = 2 * K * 1024 * 1024 * 128          # both directions are computed
    print(json.dumps({"pairs": K, "ms_per_batch": ms, "us_per_pair": ms * 1e3 / K, "int8_TOPS": 2 * macs / (ms * 1e-3) / 1e12,
                      "descriptor_GBps": K * 2 * 1024 * 128 / (ms * 1e-3) / 1e9}))
