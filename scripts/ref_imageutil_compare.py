"""Runs the REFERENCE's own image kernels (oracle/_ref/libref_imageutil.so: FL/CUDAImageUtil.cu called in the order of CUDACache::storeFrame
and CUDAImageManager::process, plus the three trajectory stubs of FL/OnlineBundler.cu) beside this repo's CUDA path (rows a20, a21, a22) on
the same synthetic inputs and prints a comparison report (JSON lines).  Torch-free; asserts nothing -- the first step of pinning those
rows' oracles against the reference (read the report, then turn what holds into tests).

    python scripts/ref_imageutil_compare.py > gpurun_out/ref_imageutil_compare.jsonl
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bundlefusion_b200 import _capi as capi          # noqa: E402
from bundlefusion_b200 import synth                   # noqa: E402
from tests._cudart import DevBuf                      # noqa: E402


class RefFrame(C.Structure):                          # CUDACachedFrame BY VALUE (six device pointers)
    _fields_ = [(n, C.c_void_p) for n in ("depth", "campos", "intensity", "derivs", "normalsU", "normals")]


def cmp(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        fa, fb = np.isfinite(a), np.isfinite(b)
        same_mask = bool(np.array_equal(fa, fb) and np.array_equal(a[~fa], b[~fb]))
        both = fa & fb
        return {"equal_bits": bool(np.array_equal(a.view(np.uint32), b.view(np.uint32))), "same_invalid_mask": same_mask,
                "max_abs_diff": float(np.abs(a[both] - b[both]).max()) if both.any() else 0.0, "n_diff": int((a[both] != b[both]).sum())}
    return {"equal": bool(np.array_equal(a, b)), "n_diff": int((a != b).sum()), "max_abs_diff": int(np.abs(a.astype(np.int64) - b.astype(np.int64)).max()) if a.size else 0}


def main():
    L = capi.lib()
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_imageutil.so"))
    vp, u, f, i = C.c_void_p, C.c_uint, C.c_float, C.c_int
    R.refCacheStoreFrame.argtypes = [vp, u, u, vp, u, u, u, u, C.POINTER(f), f, f, f, RefFrame, vp, vp, vp, vp]
    R.refIngestFrame.argtypes = [vp, vp, u, u, vp, u, u, u, u, i, i, f, f, f, f, vp, vp]
    out = lambda **kw: print(json.dumps(kw), flush=True)

    for frame_idx, (W, H) in ((100, (640, 480)), (250, (640, 480)), (40, (320, 240))):
        depth, color, _ = synth.make_frame(frame_idx, W, H)
        fx = 525.0 * W / 640.0
        K = np.array([[fx, 0, (W - 1) / 2.0, 0], [0, fx, (H - 1) / 2.0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        Kinv = np.linalg.inv(K).astype(np.float32)
        cw, ch = 80, 60
        # ---- a20: cache frame ----
        res = {}
        for who in ("ref", "ours"):
            d_depth, d_color = DevBuf(depth.astype(np.float32)), DevBuf(color.astype(np.uint8))
            bufs = {"depth": DevBuf(np.zeros((ch, cw), np.float32)), "campos": DevBuf(np.zeros((ch, cw, 4), np.float32)), "intensity": DevBuf(np.zeros((ch, cw), np.float32)),
                    "derivs": DevBuf(np.zeros((ch, cw, 2), np.float32)), "normalsU": DevBuf(np.zeros((ch, cw, 4), np.uint8)), "normals": DevBuf(np.zeros((ch, cw, 4), np.float32))}
            if who == "ref":
                fr = RefFrame(*[bufs[n].ptr for n in ("depth", "campos", "intensity", "derivs", "normalsU", "normals")])
                h1, h2, h3, h4 = DevBuf(np.zeros((H, W), np.float32)), DevBuf(np.zeros((H, W, 4), np.float32)), DevBuf(np.zeros((H, W, 4), np.float32)), DevBuf(np.zeros((ch, cw), np.float32))
                rc = R.refCacheStoreFrame(d_depth.ptr, W, H, d_color.ptr, W, H, cw, ch, Kinv.reshape(16).ctypes.data_as(C.POINTER(f)), 2.5, 1.0, 0.05, fr, h1.ptr, h2.ptr, h3.ptr, h4.ptr)
            else:
                P = capi.BFCacheParams(W, H, W, H, cw, ch, (C.c_float * 16)(*Kinv.reshape(16).tolist()), 2.5, 1.0, 0.05)
                fr = capi.BFCUDACachedFrame(bufs["depth"].ptr, bufs["campos"].ptr, bufs["intensity"].ptr, bufs["derivs"].ptr, bufs["normalsU"].ptr, bufs["normals"].ptr)
                rc = L.bfCacheStoreFrame(C.byref(P), d_depth.ptr, d_color.ptr, C.byref(fr))
            res[who] = {k: v.get() for k, v in bufs.items()}; res[who]["rc"] = rc
        out(test="cache_store_frame", frame=frame_idx, size=[W, H], rc=[res["ref"]["rc"], res["ours"]["rc"]],
            **{k: cmp(res["ref"][k], res["ours"][k]) for k in ("depth", "campos", "normals", "normalsU", "intensity", "derivs")})
        # ---- a21: ingest ----
        for (wi, hi, erode, sigD) in ((W, H, 1, 2.0), (W // 2, H // 2, 1, 2.0), (W, H, 0, 0.0)):
            res = {}
            for who in ("ref", "ours"):
                d_raw, d_color = DevBuf(depth.astype(np.float32)), DevBuf(color.astype(np.uint8))
                d_od, d_oc = DevBuf(np.zeros((hi, wi), np.float32)), DevBuf(np.zeros((hi, wi, 4), np.uint8))
                if who == "ref":
                    d_f = DevBuf(np.zeros((H, W), np.float32))
                    rc = R.refIngestFrame(d_raw.ptr, d_f.ptr, W, H, d_color.ptr, W, H, wi, hi, erode, 3, 0.05, 0.3, sigD, 0.05, d_od.ptr, d_oc.ptr)
                else:
                    P = capi.BFIngestParams(W, H, W, H, wi, hi, 2 if erode else 0, 3, 0.05, 0.3, sigD, 0.05)
                    rc = L.bfIngestFrame(C.byref(P), d_raw.ptr, d_color.ptr, d_od.ptr, d_oc.ptr)
                res[who] = dict(depth=d_od.get(), color=d_oc.get(), rc=rc)
            out(test="ingest_frame", frame=frame_idx, size=[W, H], integration=[wi, hi], erode=erode, sigmaD=sigD, rc=[res["ref"]["rc"], res["ours"]["rc"]],
                depth=cmp(res["ref"]["depth"], res["ours"]["depth"]), color=cmp(res["ref"]["color"], res["ours"]["color"]))

    # ---- a22: trajectory stubs (same names in both libraries; each handle resolves its own) ----
    rng = np.random.default_rng(3)
    sub, nGlob = 10, 6
    nAll = sub * nGlob
    glob = np.stack([synth.se3_exp(rng.standard_normal(3) * 0.2, rng.standard_normal(3)) for _ in range(nGlob)]).astype(np.float32)
    loc = np.stack([synth.se3_exp(rng.standard_normal(3) * 0.05, rng.standard_normal(3) * 0.1) for _ in range(nGlob * (sub + 1))]).astype(np.float32)
    glob[3] = -np.inf; loc[17] = -np.inf
    for who, lib in (("ref", R), ("ours", L)):
        lib.updateTrajectoryCU.argtypes = [vp, u, vp, u, vp, u, u, vp]; lib.updateTrajectoryCU.restype = None
    res = {}
    for who, lib in (("ref", R), ("ours", L)):
        d_g, d_l, d_c, d_inv = DevBuf(glob), DevBuf(loc), DevBuf(np.zeros((nAll, 4, 4), np.float32)), DevBuf(np.zeros(nAll, np.int32))
        lib.updateTrajectoryCU(d_g.ptr, nGlob, d_c.ptr, nAll, d_l.ptr, sub + 1, nGlob, d_inv.ptr)
        res[who] = dict(complete=d_c.get(), inv=d_inv.get())
    out(test="updateTrajectoryCU", complete=cmp(res["ref"]["complete"], res["ours"]["complete"]), invalidate_list=cmp(res["ref"]["inv"], res["ours"]["inv"]))


if __name__ == "__main__":
    main()
