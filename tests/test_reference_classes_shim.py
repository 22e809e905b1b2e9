"""include/bf_reference_classes.hpp: the reference's SIFT class surface (SiftGPU, SiftMatchGPU, SIFTImageManager -- FL/SiftGPU/SiftGPU.h:73-114,
SiftMatch.h:15-47, SIFTImageManager.h:62-330) as a header-only C++ shim over the C-ABI.

CPU: tests/shim/shim_bundler.cpp -- Bundler::detectFeatures / matchAndFilter / fuseToGlobal written against those classes -- compiles with g++
and links against libbundlefusion_b200.so with no unresolved symbol.
GPU: the program runs on synthetic frames; every output (key points, descriptors, matched frame, correspondence list, fused keyframe) equals what
the same sequence gives when driven through the C-ABI from Python (a second, independent piece of bookkeeping: strided key layout, no classes).
"""
import ctypes as C
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "shim", "shim_bundler.cpp")
LIBDIR = os.path.join(ROOT, "bundlefusion_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_shim(tmpdir) -> str:
    exe = os.path.join(str(tmpdir), "shim_bundler")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"), SRC, "-o", exe,
           "-L" + LIBDIR, "-lbundlefusion_b200", "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + os.path.join(CUDA, "lib64"),
           "-Wl,--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_shim_compiles_and_links(tmp_path):
    assert os.path.exists(os.path.join(LIBDIR, "libbundlefusion_b200.so")), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    exe = build_shim(tmp_path)
    # every class member resolves to library symbols: the dynamic symbol table of the program names them
    nm = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("bfSiftDetect", "bfSiftMatchBatch", "bfSiftSortKeyPointMatches", "bfSiftFilterKeyPointMatches", "bfSiftFilterMatchesBySurfaceArea",
                "bfSiftFilterMatchesByDenseVerify", "bfSiftVerifyTrajectory", "bfSiftAddCurrToResiduals", "bfSiftInvalidateImageToImage",
                "bfSiftCheckForInvalidFrames", "bfSiftFilterFrames", "bfSiftFuseToGlobal"):
        assert sym in nm, sym
    # without a device the program must fail loudly, not fall back
    r = subprocess.run([exe, "/nonexistent", "/dev/null"], capture_output=True, text=True)
    assert r.returncode == 2


@pytest.mark.gpu
def test_shim_run_equals_c_abi_sequence(tmp_path):
    import torch
    from bundlefusion_b200 import _capi as capi
    from bundlefusion_b200 import synth
    from bundlefusion_b200.cache import intrinsics_inverse

    W, H, n, maxKeys = 320, 240, 4, 512
    K4 = np.eye(4, dtype=np.float32)
    fx = 525.0 * W / 640.0
    K4[0, 0] = K4[1, 1] = fx; K4[0, 2] = (W - 1) / 2.0; K4[1, 2] = (H - 1) / 2.0
    Kinv = intrinsics_inverse(K4).astype(np.float32)
    inten, depth = [], []
    for i in range(n):
        # one camera position, fresh sensor noise per frame: the chunk's poses are the identity, which is what the program hands fuseToGlobal
        d, c, _ = synth.make_frame(100 + i, W, H, pose=synth.lissajous_pose(100), texture="rich")
        c = c.astype(np.float32)
        inten.append(((0.299 * c[..., 0] + 0.587 * c[..., 1] + 0.114 * c[..., 2]) / 255.0).astype(np.float32))
        depth.append(d.astype(np.float32))
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as fp:
        fp.write(struct.pack("4I", W, H, n, maxKeys)); fp.write(K4.tobytes()); fp.write(Kinv.tobytes())
        for i in range(n):
            fp.write(inten[i].tobytes()); fp.write(depth[i].tobytes())
    exe = build_shim(tmp_path)
    r = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr

    # ---- the same sequence through the C-ABI, strided layout ----
    lib = capi.lib()
    dev = torch.device("cuda:0")
    lib.bfSetStream(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    z = lambda *s, dt=torch.int32: torch.zeros(*s, dtype=dt, device=dev)
    keys = z(n, maxKeys, 4, dt=torch.float32); des = z(n, maxKeys, 128, dt=torch.uint8)
    numKeysDev = z(1); counts = []
    numRaw = z(n + 1); rawD = z((n + 1) * 128, dt=torch.float32); rawI = z((n + 1) * 128, 2)
    numF = z(n + 1); fD = z((n + 1) * 25, dt=torch.float32); fI = z((n + 1) * 25, 2)
    T = z(n + 1, 16, dt=torch.float32); Ti = z(n + 1, 16, dt=torch.float32)
    valid = z(n + 1); valid[0] = 1
    last = z(1); nRes = z(1)
    glob = z(25 * (n + 1) * n // 2, 8); globI = z(25 * (n + 1) * n // 2, 2)
    KinvC = (C.c_float * 16)(*Kinv.reshape(-1)); KC = (C.c_float * 16)(*K4.reshape(-1))
    p = capi.BFSiftDetectParams(W, H, W, H, 0.1, 4.0, 3.0, 150, maxKeys)
    per_frame = []
    for i in range(n):
        di = torch.from_numpy(inten[i]).to(dev); dd = torch.from_numpy(depth[i]).to(dev)
        capi.check(lib.bfSiftDetect(C.byref(p), di.data_ptr(), dd.data_ptr(), keys[i].data_ptr(), des[i].data_ptr(), numKeysDev.data_ptr(), None), "detect")
        k = min(int(numKeysDev.item()), maxKeys); counts.append(k)
        lastMatched = -1
        if i >= 1 and k > 0:
            jobs = (capi.BFSiftMatchJob * i)()
            for prev in range(i):
                j = jobs[prev]
                j.d_des1 = des[prev].data_ptr(); j.num1 = counts[prev]; j.d_des2 = des[i].data_ptr(); j.num2 = k
                j.out.d_numMatches = numRaw[prev:].data_ptr(); j.out.d_distances = rawD[prev * 128:].data_ptr(); j.out.d_keyPointIndices = rawI[prev * 128:].data_ptr()
                j.keyPointOffset[0] = prev * maxKeys; j.keyPointOffset[1] = i * maxKeys
            capi.check(lib.bfSiftMatchBatch(jobs, i, 0.7, 0.8), "match")
            nf = i + 1
            capi.check(lib.bfSiftSortKeyPointMatches(i, 0, nf, numRaw.data_ptr(), rawD.data_ptr(), rawI.data_ptr()), "sort")
            capi.check(lib.bfSiftFilterKeyPointMatches(i, 0, nf, keys.data_ptr(), numRaw.data_ptr(), rawD.data_ptr(), rawI.data_ptr(), numF.data_ptr(), fD.data_ptr(), fI.data_ptr(),
                                                       T.data_ptr(), Ti.data_ptr(), KinvC, 5, 0.0004), "filter")
            capi.check(lib.bfSiftFilterMatchesBySurfaceArea(i, 0, nf, keys.data_ptr(), numF.data_ptr(), fI.data_ptr(), KinvC, 0.032, None), "area")
            capi.check(lib.bfSiftFilterFrames(i, 0, nf, numF.data_ptr(), valid.data_ptr(), last.data_ptr()), "filterFrames")
            lastMatched = int(last.item())
            if lastMatched >= 0:
                capi.check(lib.bfSiftAddCurrToResiduals(i, 0, nf, glob.data_ptr(), globI.data_ptr(), nRes.data_ptr(), numF.data_ptr(), fI.data_ptr(), keys.data_ptr(), KinvC), "residuals")
        per_frame.append((k, lastMatched, int(nRes.item())))
    nCorr = int(nRes.item())
    poses = torch.eye(4, device=dev).reshape(1, 16).repeat(n, 1).contiguous()
    cnt = torch.tensor(counts, dtype=torch.int32, device=dev)
    outK = z(maxKeys, 4, dt=torch.float32); outD = z(maxKeys, 128, dt=torch.uint8); outN = z(1)
    capi.check(lib.bfSiftFuseToGlobal(glob.data_ptr(), globI.data_ptr(), nRes.data_ptr(), poses.data_ptr(), n, keys.data_ptr(), des.data_ptr(), cnt.data_ptr(), maxKeys, KC,
                                       max(nCorr, 1), outK.data_ptr(), outD.data_ptr(), outN.data_ptr(), maxKeys, None), "fuse")
    torch.cuda.synchronize()

    # ---- compare ----
    raw = open(fout, "rb").read()
    off = 0
    for i in range(n):
        k, lm, nc = struct.unpack_from("iiI", raw, off); off += 12
        assert (k, lm, nc) == per_frame[i], (i, (k, lm, nc), per_frame[i])
    assert max(c for c in counts) > 50 and per_frame[-1][2] > 0, "the synthetic frames must exercise the chain"
    for i in range(n):
        k = counts[i]
        a = np.frombuffer(raw, np.float32, k * 4, off).reshape(k, 4); off += k * 16
        b = np.frombuffer(raw, np.uint8, k * 128, off).reshape(k, 128); off += k * 128
        assert np.array_equal(a.view(np.uint32), keys[i, :k].cpu().numpy().view(np.uint32))
        assert np.array_equal(b, des[i, :k].cpu().numpy())
    a = np.frombuffer(raw, np.uint32, nCorr * 8, off).reshape(nCorr, 8); off += nCorr * 32
    assert np.array_equal(a, glob[:nCorr].cpu().numpy().view(np.uint32))
    (fused,) = struct.unpack_from("i", raw, off); off += 4
    assert fused == int(outN.item()) and fused > 0
    a = np.frombuffer(raw, np.uint32, fused * 4, off).reshape(fused, 4); off += fused * 16
    b = np.frombuffer(raw, np.uint8, fused * 128, off).reshape(fused, 128); off += fused * 128
    assert np.array_equal(a, outK[:fused].cpu().numpy().view(np.uint32))
    assert np.array_equal(b, outD[:fused].cpu().numpy())
    assert off == len(raw)
