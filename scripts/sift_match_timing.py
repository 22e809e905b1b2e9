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
    for _ in range(3): m.matchBatch(jobs)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if K <= 16 else 5
    a.record()
    for _ in range(reps): m.matchBatch(jobs)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    macs = 2 * K * 1024 * 1024 * 128          # both directions are computed
    print(json.dumps({"pairs": K, "ms_per_batch": ms, "us_per_pair": ms * 1e3 / K, "int8_TOPS": 2 * macs / (ms * 1e-3) / 1e12,
                      "descriptor_GBps": K * 2 * 1024 * 128 / (ms * 1e-3) / 1e9}))
