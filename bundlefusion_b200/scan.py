"""A recorded sequence in, a mesh and a trajectory out: the application-level path of the reference (FriedLiver's main loop over a ``SensorDataReader``: every frame through
``OnlineBundler::processInput`` / ``process`` and ``reintegrate`` / ``integrate``, then ``CUDAMarchingCubesHashSDF::extractIsoSurface`` + ``saveMesh`` and the optimised
trajectory; FL/FriedLiver.cpp, FL/DepthSensing/DepthSensing.cpp:966-1129, 1180-1230) over this library's pieces: ``sens.SensorDataReader`` -> ``FrameLoop.step`` (look-ahead on
the next decoded frame) -> ``CUDAMarchingCubesHashSDF``.  Host glue only; every stage is the C-ABI's.

    python -m bundlefusion_b200.scan scan.sens out.ply [--frames N] [--device cuda:0]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json

import numpy as np

from . import _capi as capi
from .frame_loop import FrameLoop, default_params
from .marching_cubes import CUDAMarchingCubesHashSDF, marching_cubes_params
from .sens import SensorDataReader


class _LoopScene:
    """what CUDAMarchingCubesHashSDF.extractIsoSurface reads of a scene: the loop's hash data and parameters"""

    def __init__(self, loop: FrameLoop):
        self.m_hashData = loop.lib.bfFrameLoopGetHashData(loop._h).contents
        self.m_hashParams = loop.lib.bfFrameLoopGetHashParams(loop._h).contents


def reconstruct(sens_path: str, ply_path: str | None = None, device="cuda:0", max_frames: int | None = None, hash_buckets: int | None = None,
                sdf_blocks: int | None = None, max_num_triangles: int = 6_000_000, tweak=None) -> dict:
    """runs the sequence; returns {"frames", "valid", "keyframes", "trajectory" [n, 4, 4], "mesh_path", "triangles", "status" (per frame)}"""
    r = SensorDataReader(sens_path)
    hd = r.header
    n = len(r) if max_frames is None else min(len(r), max_frames)
    P = default_params(hd.depthWidth, hd.depthHeight)
    P.colorWidth, P.colorHeight = hd.colorWidth, hd.colorHeight          # the ingest resamples colour to the integration size, as CUDAImageManager::process does
    for k in range(16):
        P.depthIntrinsics[k] = hd.depthIntrinsic[k]; P.colorIntrinsics[k] = hd.colorIntrinsic[k]
    P.maxNumFrames = max(32, n + int(P.submapSize))                                  # frame store / trajectories: the sequence plus one chunk of slack
    P.maxNumImages = max(8, (int(P.maxNumFrames) + int(P.submapSize) - 1) // int(P.submapSize) + 1)      # one keyframe per chunk
    if hash_buckets is not None:
        P.hash.m_hashNumBuckets = hash_buckets
    if sdf_blocks is not None:
        P.hash.m_numSDFBlocks = sdf_blocks
    if tweak is not None:
        tweak(P)
    loop = FrameLoop(P, device)
    status = []
    cur = r.frame(0, pinned=True) if n else None
    for i in range(n):
        nxt = r.frame(i + 1, pinned=True) if i + 1 < n else None                      # decode the next frame while the device works on this one
        st = loop.step(cur[0], cur[1], *((nxt[0], nxt[1]) if nxt is not None else (None, None)))
        status.append(st.as_dict())
        cur = nxt
    # after the last frame the reference keeps turning its loop without input (FL/OnlineBundler.cpp:170-197) until the last, partial chunk is solved and the
    # re-integration list is empty: one chunk's worth of turns plus the frames one solve may move
    for _ in range(int(P.submapSize) + 2):
        loop.step_past_end()
    loop.join()
    out = {"frames": n, "valid": int(sum(s["validTransform"] for s in status)), "keyframes": loop.counters()["keyframes"], "trajectory": loop.trajectory(n), "status": status,
           "mesh_path": None, "triangles": 0}
    if ply_path is not None:
        mp = marching_cubes_params(int(P.hash.m_hashNumBuckets), float(P.hash.m_virtualVoxelSize), max_num_triangles)
        mc = CUDAMarchingCubesHashSDF(mp, device)
        out["triangles"] = mc.extractIsoSurface(_LoopScene(loop)) // 3
        out["mesh_path"] = mc.saveMesh(ply_path, overwrite=True)
        mc.close()
    loop.close(); r.close()
    return out


def ate(estimated, recorded) -> dict:
    """Absolute trajectory error of camera-to-world poses [n, 4, 4] against the poses recorded with the sequence (the `.sens` frames carry them, e.g. ScanNet's): the
    camera centres of the frames valid in both are aligned by the least-squares rigid transform (Horn / Kabsch, float64, no scale), then RMSE / mean / max of the residual
    distances in metres.  Frames whose pose is not finite on either side (lost tracking: -inf) are left out."""
    E, R = np.asarray(estimated, np.float64), np.asarray(recorded, np.float64)
    n = min(len(E), len(R))
    ok = np.array([np.isfinite(E[i]).all() and np.isfinite(R[i]).all() for i in range(n)], bool)
    if ok.sum() < 3:
        return {"frames": int(ok.sum()), "rmse": float("nan"), "mean": float("nan"), "max": float("nan")}
    a, b = E[:n][ok][:, :3, 3], R[:n][ok][:, :3, 3]
    ca, cb = a.mean(0), b.mean(0)
    U, _, Vt = np.linalg.svd((a - ca).T @ (b - cb))
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
    Rot = Vt.T @ D @ U.T
    d = np.linalg.norm((Rot @ (a - ca).T).T + cb - b, axis=1)
    return {"frames": int(ok.sum()), "rmse": float(np.sqrt((d * d).mean())), "mean": float(d.mean()), "max": float(d.max())}


def recorded_poses(sens_path: str, max_frames: int | None = None) -> np.ndarray:
    """the camera-to-world poses stored with the frames ([n, 4, 4]; all -inf where the recorder had none)"""
    r = SensorDataReader(sens_path)
    n = len(r) if max_frames is None else min(len(r), max_frames)
    out = np.stack([r.frame_pose(i) for i in range(n)]) if n else np.zeros((0, 4, 4), np.float32)
    r.close()
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("sens"); ap.add_argument("ply")
    ap.add_argument("--frames", type=int, default=None); ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--hash-buckets", type=int, default=None); ap.add_argument("--sdf-blocks", type=int, default=None)
    ap.add_argument("--trajectory", default=None, help="write the camera-to-world poses here (one 4x4 per frame, text)")
    a = ap.parse_args(argv)
    o = reconstruct(a.sens, a.ply, a.device, a.frames, a.hash_buckets, a.sdf_blocks)
    if a.trajectory:
        np.savetxt(a.trajectory, o["trajectory"].reshape(-1, 16))
    summary = {k: o[k] for k in ("frames", "valid", "keyframes", "triangles", "mesh_path")}
    gt = recorded_poses(a.sens, a.frames)
    if np.isfinite(gt).all(axis=(1, 2)).sum() >= 3:
        summary["ate"] = ate(o["trajectory"], gt[:len(o["trajectory"])])
    print(json.dumps(summary))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
