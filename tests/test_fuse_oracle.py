"""SIFTImageManager::fuseToGlobal / computeTracks (FL/SiftGPU/SIFTImageManager.cpp:366-476): the oracle (oracle/fuse_oracle.c) against an
independent Python restatement with real recursion, and the CUDA kernel (csrc/sift_fuse.cu) run on the CPU through the emulation of
tests/cuda_emu against the oracle, bit for bit."""
import ctypes as C
import sys

import numpy as np
import pytest

from bundlefusion_b200 import synth
from oracle import oracle as orc
from tests.cuda_emu import build_emulated

F = np.float32


def python_fuse(pb, max_keys=1024):
    """line-by-line Python version of computeTracks + fuseToGlobal (float32 arithmetic through numpy scalars)"""
    sys.setrecursionlimit(20000)
    corr, kidx, T, keys, descs, num, K, S = (pb[k] for k in ("corr", "keyIdx", "transforms", "keys", "descs", "numKeys", "K", "keyStride"))
    def xf(M, p):
        return np.array([F(F(F(M[r, 0] * p[0]) + F(M[r, 1] * p[1])) + F(M[r, 2] * p[2])) + M[r, 3] for r in range(3)], F)
    per_key = {}
    for c in range(len(corr)):
        e = corr[c]
        if e["i"] == 0xFFFFFFFF:
            continue
        kx, ky = int(kidx[c, 0]), int(kidx[c, 1])
        a, b = xf(T[e["i"]], e["pi"]), xf(T[e["j"]], e["pj"])
        d = a - b
        err = np.sqrt(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))
        ok = err < F(0.03)
        per_key.setdefault(kx, []).append((int(e["j"]), ky, e["pj"].copy() if ok else None))
        per_key.setdefault(ky, []).append((int(e["i"]), kx, e["pi"].copy() if ok else None))
    marker = set()
    def find(track, cur):
        for (img, key, pos) in per_key.get(cur, []):
            if key not in marker:
                track.append((img, key, pos)); marker.add(key)
                find(track, key)
    out_k, out_d = [], []
    for i in range(len(num)):
        for k in range(int(num[i])):
            track = []
            find(track, i * S + k)
            if not track:
                continue
            pos = np.zeros(3, F); n = 0
            for (img, key, p) in track:
                if p is not None:
                    pos = (pos + xf(T[img], p)).astype(F); n += 1
            if n == 0:
                continue
            pos = (pos / F(n)).astype(F)
            q = xf(K, pos)
            if len(out_k) < max_keys:
                out_k.append((F(q[0] / q[2]), F(q[1] / q[2]), keys[track[0][1], 2], q[2])); out_d.append(descs[track[0][1]])
    return np.array(out_k, F).reshape(-1, 4), np.array(out_d, np.uint8).reshape(-1, 128)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fuse_oracle_matches_python_restatement(seed):
    pb = synth.make_fuse_problem(seed=seed)
    ko, do = orc.sift_fuse_to_global(pb["corr"], pb["keyIdx"], pb["transforms"], pb["keys"], pb["descs"], pb["numKeys"], pb["keyStride"], pb["K"])
    kp, dp = python_fuse(pb)
    assert len(ko) == len(kp) > 50
    assert np.array_equal(ko.view(np.uint32), kp.view(np.uint32)) and np.array_equal(do, dp)
    # every fused key lies where its 3-D point projects into the first frame (outlier correspondences never contribute a position)
    assert (ko[:, 3] > 0.9).all() and (ko[:, 3] < 3.2).all()


def test_fuse_oracle_edge_cases():
    pb = synth.make_fuse_problem(seed=3, n_images=3, n_points=20)
    # no correspondences at all -> no keys; all invalid -> no keys; capacity smaller than the number of tracks -> the first tracks
    k, d = orc.sift_fuse_to_global(pb["corr"][:0], pb["keyIdx"][:0], pb["transforms"], pb["keys"], pb["descs"], pb["numKeys"], pb["keyStride"], pb["K"])
    assert len(k) == 0
    c2 = pb["corr"].copy(); c2["i"][:] = 0xFFFFFFFF
    assert len(orc.sift_fuse_to_global(c2, pb["keyIdx"], pb["transforms"], pb["keys"], pb["descs"], pb["numKeys"], pb["keyStride"], pb["K"])[0]) == 0
    full, _ = orc.sift_fuse_to_global(pb["corr"], pb["keyIdx"], pb["transforms"], pb["keys"], pb["descs"], pb["numKeys"], pb["keyStride"], pb["K"])
    cut, _ = orc.sift_fuse_to_global(pb["corr"], pb["keyIdx"], pb["transforms"], pb["keys"], pb["descs"], pb["numKeys"], pb["keyStride"], pb["K"], maxKeys=5)
    assert len(full) > 5 and np.array_equal(cut, full[:5])


@pytest.fixture(scope="module")
def fuse_emu():
    L = build_emulated("sift_fuse.cu", 1)
    vp, u = C.c_void_p, C.c_uint
    L.bfSiftFuseToGlobal.argtypes = [vp, vp, vp, vp, u, vp, vp, vp, u, C.POINTER(C.c_float), u, vp, vp, vp, u, vp]
    return L


def run_fuse(L, pb, max_keys=1024, to_dev=lambda a: a, from_dev=lambda a: a):
    corr = np.ascontiguousarray(pb["corr"]); n = np.array([len(corr)], np.int32)
    ok, od, on, st = np.zeros((max_keys, 4), F), np.zeros((max_keys, 128), np.uint8), np.full(1, -1, np.int32), np.full(1, -1, np.int32)
    Kp = np.ascontiguousarray(pb["K"], F).reshape(16).ctypes.data_as(C.POINTER(C.c_float))
    bufs = [to_dev(x) for x in (corr if len(corr) else np.zeros(1, corr.dtype), np.ascontiguousarray(pb["keyIdx"], np.uint32) if len(corr) else np.zeros((1, 2), np.uint32), n,
                                np.ascontiguousarray(pb["transforms"], F), np.ascontiguousarray(pb["keys"], F), np.ascontiguousarray(pb["descs"], np.uint8),
                                np.ascontiguousarray(pb["numKeys"], np.int32), ok, od, on, st)]
    ptr = lambda b: b.ptr if hasattr(b, "ptr") else b.ctypes.data
    rc = L.bfSiftFuseToGlobal(ptr(bufs[0]), ptr(bufs[1]), ptr(bufs[2]), ptr(bufs[3]), len(pb["transforms"]), ptr(bufs[4]), ptr(bufs[5]), ptr(bufs[6]), pb["keyStride"], Kp,
                              max(1, len(corr)), ptr(bufs[7]), ptr(bufs[8]), ptr(bufs[9]), max_keys, ptr(bufs[10]))
    assert rc == 0
    ok, od, on, st = (from_dev(b) for b in bufs[7:])
    return ok[: int(on[0])], od[: int(on[0])], int(st[0])


@pytest.mark.parametrize("seed", [0, 4])
def test_fuse_kernel_emulated_matches_oracle(fuse_emu, seed):
    pb = synth.make_fuse_problem(seed=seed, n_images=5, n_points=90, key_stride=128)
    ko, do = orc.sift_fuse_to_global(pb["corr"], pb["keyIdx"], pb["transforms"], pb["keys"], pb["descs"], pb["numKeys"], pb["keyStride"], pb["K"])
    kg, dg, st = run_fuse(fuse_emu, pb)
    assert st == 0 and len(kg) == len(ko) > 30
    assert np.array_equal(kg.view(np.uint32), ko.view(np.uint32)) and np.array_equal(dg, do)
    kg, dg, st = run_fuse(fuse_emu, pb, max_keys=7)
    assert np.array_equal(kg.view(np.uint32), ko[:7].view(np.uint32)) and np.array_equal(dg, do[:7])
