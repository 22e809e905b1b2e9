"""Generates tests/golden/*.npz on a GPU box: outputs of the REFERENCE'S OWN CUDA (oracle/_ref, IEEE build, see oracle/build_ref.py) on small
seeded inputs that bundlefusion_b200/synth.py regenerates anywhere.  tests/test_golden_reference.py replays the same inputs through the CPU
oracle in the `-m "not gpu"` suite and compares -- the oracle pinned against the reference without a GPU in the loop.
Run:  gpurun -- python scripts/make_golden_from_reference.py   (writes gpurun_out/golden/, copy into tests/golden/)."""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bundlefusion_b200 import synth
from bundlefusion_b200.scene_rep import camera_params, default_hash_params
from bundlefusion_b200.solver import DeviceCache
from oracle import oracle as orc
from oracle import ref_solver, ref_tsdf

TSDF_CASE = {"W": 160, "H": 120, "frames": [0, 35, 70], "num_buckets": 20011, "num_sdf_blocks": 30000, "reint": 1, "shift": [0.011, -0.006, 0.004]}
BA_SPARSE = {"n_images": 11, "degree": 10, "corr_per_pair": 25, "noise": 0.002, "seed": 5, "n_gn": 2, "n_pcg": 100}
BA_DENSE = {"n_images": 5, "stride": 3, "W": 320, "H": 240, "n_gn": 2, "n_pcg": 60}


def block_crcs(vox):
    return np.array([zlib.crc32(np.ascontiguousarray(v).tobytes()) for v in vox], np.uint32)


def main():
    dev = torch.device("cuda:0")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    # ---- TSDF through the reference's kernels (IEEE build) ----
    c = TSDF_CASE
    cam = camera_params(c["W"], c["H"])
    hp = default_hash_params(num_buckets=c["num_buckets"], num_sdf_blocks=c["num_sdf_blocks"])
    ref = ref_tsdf.ReferenceSceneRepHashSDF(hp, dev, fast_math=False)
    frames = [synth.make_frame(i, c["W"], c["H"]) for i in c["frames"]]
    devf = [(torch.from_numpy(d).to(dev), torch.from_numpy(col).to(dev)) for d, col, _ in frames]
    for (d, col, T), (dd, dc) in zip(frames, devf):
        ref.integrate(T, dd, dc, cam)
    k = c["reint"]
    T = frames[k][2]; T2 = T.copy(); T2[:3, 3] += np.array(c["shift"], np.float32)
    ref.deIntegrate(T, devf[k][0], devf[k][1], cam); ref.integrate(T2, devf[k][0], devf[k][1], cam)
    ref.garbageCollect()
    snap = ref.download()
    b, v = orc.canonical_blocks(snap)
    np.savez_compressed(os.path.join(out, "tsdf_reference_ieee.npz"), pose_inverse=np.bytes_(b"reference host formula (bfMat4Inverse)"), blocks=b, crcs=block_crcs(v), first_voxels=v[:4], heap_free=np.int64(ref.getHeapFreeCount()),
                        case=np.bytes_(repr(c)))
    # ---- solver through the reference's kernels (IEEE build) ----
    s = BA_SPARSE
    prob = synth.make_ba_problem(s["n_images"], degree=s["degree"], corr_per_pair=s["corr_per_pair"], noise=s["noise"], seed=s["seed"])
    N = s["n_images"]
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    corr = tt(prob["corr"].view(np.uint8).reshape(-1).copy()); rot, trans = tt(prob["init_rot"].copy()), tt(prob["init_trans"].copy())
    valid = torch.ones(N, dtype=torch.int32, device=dev)
    rs = ref_solver.ReferenceSolverBundling(N, max(len(prob["corr"]), 1000 * N), dev, fast_math=False)
    conv = rs.solve(corr, len(prob["corr"]), valid, N, s["n_gn"], s["n_pcg"], [1.0] * s["n_gn"], d_rot=rot, d_trans=trans, record_convergence=True)
    sparse = {"rot": rot.cpu().numpy(), "trans": trans.cpu().numpy(), "energy": conv}
    dcase = BA_DENSE
    dp = synth.make_dense_ba_problem(dcase["n_images"], stride=dcase["stride"], W=dcase["W"], H=dcase["H"])
    N = dcase["n_images"]
    cache = DeviceCache(dp["caches"], dp["intrinsics"], dev)
    corr = tt(dp["corr"].view(np.uint8).reshape(-1).copy()); rot, trans = tt(dp["init_rot"].copy()), tt(dp["init_trans"].copy())
    valid = torch.ones(N, dtype=torch.int32, device=dev)
    rs = ref_solver.ReferenceSolverBundling(N, max(len(dp["corr"]), 1000 * N), dev, fast_math=False)
    rs.solve(corr, len(dp["corr"]), valid, N, dcase["n_gn"], dcase["n_pcg"], [1.0] * dcase["n_gn"], [1.0, 2.0], [0.0, 0.0], d_rot=rot, d_trans=trans, cudaCache=cache)
    np.savez_compressed(os.path.join(out, "solver_reference_ieee.npz"), sparse_rot=sparse["rot"], sparse_trans=sparse["trans"], sparse_energy=sparse["energy"],
                        dense_rot=rot.cpu().numpy(), dense_trans=trans.cpu().numpy(),
                        dense_overlap=np.int64(rs._bufs["d_numDenseOverlappingImages"].cpu().numpy()[0]),
                        sparse_case=np.bytes_(repr(BA_SPARSE)), dense_case=np.bytes_(repr(BA_DENSE)))
    print("golden written to", out, os.listdir(out))


if __name__ == "__main__":
    main()
