"""The application-level path in one call (bundlefusion_b200/scan.py: `.sens` reader -> frame loop with look-ahead -> iso-surface extraction -> saveMesh): a synthetic
recording is reconstructed and the mesh lies on the recorded surfaces.  Every stage has its own parity test; this one checks the hand-overs between them."""
import numpy as np
import pytest

from bundlefusion_b200 import scan, sens, synth

pytestmark = pytest.mark.gpu
W, H = 320, 240


def test_sens_file_to_trajectory_and_mesh(cuda_device, tmp_path):
    from bundlefusion_b200.frame_loop import default_params
    n = 14
    K = np.array(list(default_params(W, H).depthIntrinsics), np.float32).reshape(4, 4)
    path = str(tmp_path / "seq.sens")
    w = sens.SensorDataWriter(path, W, H, K, depth_shift=1000.0, zlib_depth=True)
    truth, depths = [], []
    for i in range(n):
        d, c, T = synth.make_frame(2 * i, W, H, texture="rich")
        w.append(np.where(np.isfinite(d), np.clip(np.round(d * 1000.0), 1, 65535), 0).astype(np.uint16), c[..., :3], T)
        truth.append(T); depths.append(d)
    w.finish()
    ply = str(tmp_path / "out" / "scan.ply")
    o = scan.reconstruct(path, ply, cuda_device, hash_buckets=100003, sdf_blocks=90000, max_num_triangles=2_000_000)
    assert o["frames"] == n and o["valid"] >= n - 1 and o["mesh_path"] == ply and o["triangles"] > 20000
    # the trajectory: relative to the first frame (the loop starts at the identity), within a few centimetres of the recorded camera path (the loop test's bound)
    traj = o["trajectory"]
    assert n - 3 <= len(traj) <= n
    T0inv = np.linalg.inv(truth[0].astype(np.float64))
    err = [np.abs((T0inv @ truth[i].astype(np.float64))[:3, 3] - traj[i][:3, 3].astype(np.float64)).max() for i in range(len(traj)) if np.isfinite(traj[i]).all()]
    assert len(err) >= n - 4 and max(err) < 0.05
    # the mesh: a PLY whose vertices are observed surface points -- most lie within a few centimetres of the depth some recorded frame saw along its ray (poses are good to a few centimetres)
    head, body = open(ply, "rb").read().split(b"end_header\n", 1)
    nv = int([l for l in head.decode().splitlines() if l.startswith("element vertex")][0].split()[-1])
    v = np.frombuffer(body[:16 * nv], np.dtype([("p", "<f4", 3), ("c", "u1", 4)]))["p"].astype(np.float64)
    assert nv > 10000
    fx, fy, mx, my = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    sample = v[:: max(1, nv // 4000)]
    best = np.full(len(sample), np.inf)
    for i in (0, n // 2, n - 1):                                     # the model lives in the first camera's frame
        cam = (np.linalg.inv(T0inv @ truth[i].astype(np.float64)) @ np.c_[sample, np.ones(len(sample))].T).T[:, :3]
        z = cam[:, 2]
        u = np.round(cam[:, 0] / np.maximum(z, 1e-6) * fx + mx).astype(int); vv = np.round(cam[:, 1] / np.maximum(z, 1e-6) * fy + my).astype(int)
        ok = (z > 0.1) & (u >= 0) & (u < W) & (vv >= 0) & (vv < H)
        dd = np.full(len(sample), np.inf); dd[ok] = depths[i][vv[ok], u[ok]]
        best = np.minimum(best, np.where(np.isfinite(dd), np.abs(dd - z), np.inf))
    seen = np.isfinite(best)
    assert seen.mean() > 0.9 and np.percentile(best[seen], 90) < 0.06
