"""Sharded PCG over the GPUs of one box (bfSolverPeer*): run under torchrun, one rank per GPU.
Every rank solves the same sparse problem (a) alone and (b) sharded; checks: sharded poses bit-identical on all ranks, within 1e-4 rel-L2 of the single-GPU
solve; prints the time per PCG iteration of both.   torchrun --nproc-per-node N scripts/solver_peers_check.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bundlefusion_b200 import synth
from bundlefusion_b200.solver import CUDASolverBundling

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
gloo = dist.new_group(backend="gloo")


def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


for N, deg, gn, pcg in ((500, 15, 3, 150), (2000, 15, 3, 150)):
    prob = synth.make_ba_problem(N, degree=deg, corr_per_pair=25, noise=0.002, seed=32, stride=10 if N <= 500 else 2)
    nC = len(prob["corr"])
    corr = torch.from_numpy(prob["corr"].view(np.uint8).reshape(-1).copy()).to(dev)
    r0 = torch.from_numpy(prob["init_rot"]).to(dev); t0 = torch.from_numpy(prob["init_trans"]).to(dev)
    valid = torch.ones(N, dtype=torch.int32, device=dev)
    w = [1.0] * gn
    out = {}
    for mode in ("single", "sharded"):
        sv = CUDASolverBundling(N, max(nC, 1000 * N), dev)
        if mode == "sharded":
            sv.connect_peers(gloo)
        rot, trans = r0.clone(), t0.clone()

        def run():
            rot.copy_(r0); trans.copy_(t0)
            sv.solve(corr, nC, valid, N, gn, pcg, w, d_rotationAnglesUnknowns=rot, d_translationUnknowns=trans)

        run(); torch.cuda.synchronize(); dist.barrier()
        x = np.c_[rot.cpu().numpy(), trans.cpu().numpy()]
        iters = int(sv.getStats()["pcg"])
        reps = 5
        for _ in range(2): run()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): run()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[mode] = {"x": x, "ms": float(t.item()), "iters": iters}
        if mode == "sharded":
            sv.disconnect_peers()
        sv.close()
        dist.barrier()
    xs = torch.from_numpy(out["sharded"]["x"]).to(dev)
    gathered = [torch.zeros_like(xs) for _ in range(world)]
    dist.all_gather(gathered, xs)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    if rank == 0:
        print(json.dumps({"N": N, "correspondences": nC, "world": world, "pcg_iterations": out["single"]["iters"],
                          "sharded_bit_identical_on_all_ranks": bool(same), "rel_l2_sharded_vs_single": rel(out["sharded"]["x"], out["single"]["x"]),
                          "ms_single": out["single"]["ms"], "ms_sharded": out["sharded"]["ms"],
                          "us_per_pcg_iteration_single": 1e3 * out["single"]["ms"] / max(1, out["single"]["iters"]),
                          "us_per_pcg_iteration_sharded": 1e3 * out["sharded"]["ms"] / max(1, out["sharded"]["iters"])}), flush=True)
        assert same and rel(out["sharded"]["x"], out["single"]["x"]) < 1e-4
dist.destroy_process_group()
