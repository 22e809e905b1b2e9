"""CUDA-event timing of the solver at local / global sizes (development aid)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundlefusion_b200 import synth
from bundlefusion_b200.solver import CUDASolverBundling
dev = torch.device("cuda:0")
for (N, deg, gn, pcg) in [(11, 10, 2, 100), (32, 10, 3, 150), (64, 12, 3, 150), (128, 15, 3, 150), (500, 15, 3, 150), (2000, 15, 3, 150)]:
    prob = synth.make_ba_problem(N, degree=deg, corr_per_pair=25, noise=0.002, seed=32, stride=10 if N <= 500 else 2)
    corr = torch.from_numpy(prob["corr"].view(np.uint8).reshape(-1).copy()).to(dev)
    r0 = torch.from_numpy(prob["init_rot"]).to(dev); t0 = torch.from_numpy(prob["init_trans"]).to(dev)
    valid = torch.ones(N, dtype=torch.int32, device=dev)
    sol = CUDASolverBundling(N, max(len(prob["corr"]), 1000 * N), dev)
    rot, trans = r0.clone(), t0.clone()
    def run():
        rot.copy_(r0); trans.copy_(t0)
        sol.solve(corr, len(prob["corr"]), valid, N, gn, pcg, [1.0] * gn, d_rotationAnglesUnknowns=rot, d_translationUnknowns=trans)
    for _ in range(3): run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    st = sol.getStats()
    ms = a.elapsed_time(b) / 10
    print(json.dumps({"N": N, "C": len(prob["corr"]), "ms_per_solve": ms, "gn": st["gn"], "pcg": st["pcg"], "us_per_pcg_iter": ms * 1e3 / max(1, st["pcg"]), "pairs": st["pairs"]}))
