"""Diagnosis of the dense-only bundle-adjustment parity (round-1 failure: 2.1e-2 vs the reference on one box, < 1e-4 on another).
Runs the reference's CUDA solver several times, this library and the oracle on the chaotic and on the well-posed configuration and prints
every pairwise relative L2, plus the oracle's own sensitivity to 2-ulp input perturbations.  One JSON line per configuration."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                                         # noqa: E402
from bundlefusion_b200 import synth                                   # noqa: E402
from bundlefusion_b200.solver import DeviceCache                      # noqa: E402
from oracle import oracle as orc                                      # noqa: E402
from tests.test_solver_vs_reference_gpu import oracle_sensitivity, rel_l2, run_ours, run_ref   # noqa: E402

dev = torch.device("cuda:0")
for (n, stride, gn, pcg) in ((5, 3, 3, 60), (5, 3, 1, 10), (8, 2, 2, 10), (8, 2, 3, 60)):
    prob = synth.make_dense_ba_problem(n, stride=stride, perturb_rot=0.004, perturb_trans=0.008, W=320, H=240)
    cache = DeviceCache(prob["caches"], prob["intrinsics"], dev)
    wS, wD, wC = [0.0] * gn, [1.0, 2.0, 3.0][:gn], [0.0] * gn
    empty = prob["corr"][:0]
    sens, o = oracle_sensitivity(prob, empty, gn, pcg, wS, wD, wC, n=6)
    x_orc = np.c_[o["rot"], o["trans"]]
    refs = [run_ref(dev, prob, empty, gn, pcg, wS, wD, wC, cache=cache, fast=False)[0] for _ in range(5)]
    refs_fast = [run_ref(dev, prob, empty, gn, pcg, wS, wD, wC, cache=cache, fast=True)[0] for _ in range(2)]
    ours = [run_ours(dev, prob, empty, gn, pcg, wS, wD, wC, cache=cache) for _ in range(2)]
    print(json.dumps({"images": n, "stride": stride, "gn": gn, "pcg": pcg, "oracle_ran": [int(o["gn"]), int(o["pcg"])], "ours_ran": [int(ours[0][1]["gn"]), int(ours[0][1]["pcg"])],
                      "oracle_sensitivity_2ulp": sens, "oracle_decision_margin": orc.decision_margin(o["trace"]),
                      "ref_vs_ref0": [rel_l2(r, refs[0]) for r in refs[1:]], "reffast_vs_ref0": [rel_l2(r, refs[0]) for r in refs_fast],
                      "ours_vs_ref0": rel_l2(ours[0][0], refs[0]), "ours_vs_ours": rel_l2(ours[1][0], ours[0][0]), "ours_vs_oracle": rel_l2(ours[0][0], x_orc),
                      "oracle_vs_ref0": rel_l2(x_orc, refs[0])}), flush=True)
