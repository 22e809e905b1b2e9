"""A recorded sequence through the frame loop: synthetic frames are written to a `.sens` file (millimetre depth, zlib; RGB colour), read back by the library's
reader as pinned host frames -- what FL/SensorDataReader.cpp hands the reference's loop -- and stepped through bfFrameLoopStepAhead.  The loop must track it."""
import numpy as np
import pytest

from bundlefusion_b200 import sens, synth
from bundlefusion_b200.frame_loop import FrameLoop, default_params

pytestmark = pytest.mark.gpu
W, H = 320, 240


def test_sens_file_drives_the_frame_loop(cuda_device, tmp_path):
    n = 14
    path = str(tmp_path / "seq.sens")
    p = default_params(W, H)
    K = np.array(list(p.depthIntrinsics), np.float32).reshape(4, 4)
    w = sens.SensorDataWriter(path, W, H, K, depth_shift=1000.0, zlib_depth=True)
    truth = []
    for i in range(n):
        d, c, T = synth.make_frame(2 * i, W, H, texture="rich")
        w.append(np.where(np.isfinite(d), np.clip(np.round(d * 1000.0), 1, 65535), 0).astype(np.uint16), c[..., :3], T)
        truth.append(T)
    w.finish()
    r = sens.SensorDataReader(path)
    assert len(r) == n and (r.header.depthWidth, r.header.depthHeight) == (W, H)
    fr = [r.frame(i, pinned=True) for i in range(n)]
    assert all(np.array_equal(f[2], T.astype(np.float32)) for f, T in zip(fr, truth))                 # the recorded poses come back
    p.maxNumImages = 8; p.maxNumFrames = 32
    p.hash.m_hashNumBuckets = 100003; p.hash.m_numSDFBlocks = 90000
    loop = FrameLoop(p, cuda_device)
    stats = []
    for i in range(n):
        nxt = (fr[i + 1][0], fr[i + 1][1]) if i + 1 < n else (None, None)
        stats.append(loop.step(fr[i][0], fr[i][1], *nxt).as_dict())
    assert sum(s["validTransform"] for s in stats) >= n - 1 and all(s["numKeyPoints"] > 40 for s in stats)
    assert sum(1 for s in stats if s["localSolved"] >= 0) == 1 and loop.counters()["integrations"] >= n - 1
    loop.close(); r.close()
