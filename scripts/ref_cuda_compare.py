"""Same-box comparison against the REFERENCE'S OWN CUDA (oracle/_ref, --use_fast_math build as the reference ships): parity numbers
and wall-clock per call, for the TSDF frame path and the bundle-adjustment solver.  Writes one JSON object per line.
Wall clock with a device synchronise on both sides (the reference's host loop synchronises internally, CUDA events on one stream
would not see its host time).  Development / evidence aid; not part of bench.py."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bundlefusion_b200 import synth
from bundlefusion_b200.scene_rep import CUDASceneRepHashSDF, camera_params, default_hash_params
from bundlefusion_b200.solver import CUDASolverBundling, DeviceCache
from oracle import ref_solver, ref_tsdf

dev = torch.device("cuda:0")


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


def wall(fn, reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def tsdf(voxel, n_frames, W=640, H=480):
    hp = default_hash_params(voxel_size=voxel, num_sdf_blocks=1 << 20 if voxel >= 0.01 else 1 << 21, num_buckets=1 << 20 if voxel >= 0.01 else 1 << 21)
    cam = camera_params(W, H)
    frames = []
    for k in range(n_frames):
        d, c, T = synth.make_frame(4 * k, W, H)
        frames.append((T, torch.from_numpy(d).to(dev), torch.from_numpy(c).to(dev)))
    ours = CUDASceneRepHashSDF(hp, dev)
    ref = ref_tsdf.ReferenceSceneRepHashSDF(hp, dev, fast_math=True)
    out = {}
    for name, s in (("reference_cuda", ref), ("this_repo", ours)):
        s.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for T, d, c in frames:
            s.integrate(T, d, c, cam)
        torch.cuda.synchronize()
        t_int = (time.perf_counter() - t0) / n_frames * 1e3
        # re-integration: de-integrate + integrate each frame at a nudged pose (what tryRevertIntegration does per update)
        t0 = time.perf_counter()
        for k, (T, d, c) in enumerate(frames):
            s.deIntegrate(T, d, c, cam)
            s.integrate(T, d, c, cam)
        torch.cuda.synchronize()
        t_re = (time.perf_counter() - t0) / n_frames * 1e3
        out[name] = {"ms_per_integrate": t_int, "ms_per_deintegrate_plus_integrate": t_re}
    out["blocks"] = int(ours.getNumOccupiedBlocks())
    out["speedup_integrate"] = out["reference_cuda"]["ms_per_integrate"] / out["this_repo"]["ms_per_integrate"]
    out["speedup_reintegrate"] = out["reference_cuda"]["ms_per_deintegrate_plus_integrate"] / out["this_repo"]["ms_per_deintegrate_plus_integrate"]
    print(json.dumps({"path": "tsdf", "voxel_m": voxel, "frames": n_frames, "image": [W, H], **out}), flush=True)


def solver_sparse(N, deg, gn, pcg):
    prob = synth.make_ba_problem(N, degree=deg, corr_per_pair=25, noise=0.002, seed=32, stride=10 if N <= 500 else 2)
    nC = len(prob["corr"])
    corr = torch.from_numpy(prob["corr"].view(np.uint8).reshape(-1).copy()).to(dev)
    r0 = torch.from_numpy(prob["init_rot"]).to(dev); t0 = torch.from_numpy(prob["init_trans"]).to(dev)
    valid = torch.ones(N, dtype=torch.int32, device=dev)
    w = [1.0] * gn
    ours = CUDASolverBundling(N, max(nC, 1000 * N), dev)
    ref = ref_solver.ReferenceSolverBundling(N, max(nC, 1000 * N), dev, fast_math=True)
    rot, trans = r0.clone(), t0.clone()
    c2 = corr.clone()

    def run_ours():
        rot.copy_(r0); trans.copy_(t0)
        ours.solve(corr, nC, valid, N, gn, pcg, w, d_rotationAnglesUnknowns=rot, d_translationUnknowns=trans)

    def run_ref():
        rot.copy_(r0); trans.copy_(t0)
        ref.solve(c2, nC, valid, N, gn, pcg, w, d_rot=rot, d_trans=trans)

    run_ours(); torch.cuda.synchronize(); x_our = np.c_[rot.cpu().numpy(), trans.cpu().numpy()]
    run_ref(); x_ref = np.c_[rot.cpu().numpy(), trans.cpu().numpy()]
    run_ref(); x_ref2 = np.c_[rot.cpu().numpy(), trans.cpu().numpy()]
    reps = 5 if N >= 500 else 20
    for _ in range(2):
        run_ours(); run_ref()
    t_our, t_ref = wall(run_ours, reps), wall(run_ref, reps)
    print(json.dumps({"path": "solver_sparse", "N": N, "correspondences": nC, "gn": gn, "pcg_budget": pcg, "rel_l2_ours_vs_reference": rel_l2(x_our, x_ref),
                      "rel_l2_reference_run_to_run": rel_l2(x_ref2, x_ref), "ms_reference_cuda": t_ref, "ms_this_repo": t_our, "speedup": t_ref / t_our,
                      "pcg_iters_this_repo": int(ours.getStats()["pcg"])}), flush=True)


def solver_local_dense():
    prob = synth.make_dense_ba_problem(11, stride=3, W=320, H=240)
    N = 11
    nC = len(prob["corr"])
    cache = DeviceCache(prob["caches"], prob["intrinsics"], dev)
    corr = torch.from_numpy(prob["corr"].view(np.uint8).reshape(-1).copy()).to(dev)
    r0 = torch.from_numpy(prob["init_rot"]).to(dev); t0 = torch.from_numpy(prob["init_trans"]).to(dev)
    valid = torch.ones(N, dtype=torch.int32, device=dev)
    wS, wD, wC = [1.0, 1.0], [1.0, 2.0], [0.0, 0.0]
    ours = CUDASolverBundling(N, max(nC, 1000 * N), dev)
    ref = ref_solver.ReferenceSolverBundling(N, max(nC, 1000 * N), dev, fast_math=True)
    rot, trans = r0.clone(), t0.clone()
    c2 = corr.clone()

    def run_ours():
        rot.copy_(r0); trans.copy_(t0)
        ours.solve(corr, nC, valid, N, 2, 100, wS, wD, wC, d_rotationAnglesUnknowns=rot, d_translationUnknowns=trans, cudaCache=cache)

    def run_ref():
        rot.copy_(r0); trans.copy_(t0)
        ref.solve(c2, nC, valid, N, 2, 100, wS, wD, wC, d_rot=rot, d_trans=trans, cudaCache=cache)

    run_ours(); torch.cuda.synchronize(); x_our = np.c_[rot.cpu().numpy(), trans.cpu().numpy()]
    run_ref(); x_ref = np.c_[rot.cpu().numpy(), trans.cpu().numpy()]
    for _ in range(2):
        run_ours(); run_ref()
    t_our, t_ref = wall(run_ours, 20), wall(run_ref, 20)
    print(json.dumps({"path": "solver_local_sparse_plus_dense", "N": N, "correspondences": nC, "cache": [80, 60], "rel_l2_ours_vs_reference": rel_l2(x_our, x_ref),
                      "ms_reference_cuda": t_ref, "ms_this_repo": t_our, "speedup": t_ref / t_our}), flush=True)


if __name__ == "__main__":
    tsdf(0.01, 40)
    tsdf(0.004, 20)
    solver_local_dense()
    solver_sparse(11, 10, 2, 100)
    solver_sparse(500, 15, 3, 150)
    solver_sparse(2000, 15, 3, 150)
