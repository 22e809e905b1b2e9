"""Runs the REFERENCE's own match-manager kernels (oracle/_ref/libref_siftmgr.so, built by oracle/build_ref.py from
FL/SiftGPU/SIFTImageManager.cu) beside this repo's CUDA path and its oracle on the same synthetic inputs and prints a comparison report
(JSON lines).  Torch-free.  It asserts nothing: it is the first step of pinning rows a19's oracles against the reference -- read the
report, then turn what holds into tests (as scripts/ref_cuda_compare.py was for the TSDF rows).

    python scripts/ref_siftmgr_compare.py [--fast] > gpurun_out/ref_siftmgr_compare.jsonl
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
from oracle import oracle as orc                      # noqa: E402
from tests._cudart import DevBuf                      # noqa: E402
from tests.test_verify_filters_oracle import VERIFY   # noqa: E402


def f16(m):
    return np.ascontiguousarray(m, np.float32).reshape(16).ctypes.data_as(C.POINTER(C.c_float))


def main():
    fast = "--fast" in sys.argv
    L = capi.lib()
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_siftmgr_fast.so" if fast else "libref_siftmgr.so"))
    vp, u, f = C.c_void_p, C.c_uint, C.c_float
    R.refSortKeyPointMatches.argtypes = [u, u, u, vp, vp, vp]
    R.refFilterKeyPointMatches.argtypes = [u, u, u] + [vp] * 9 + [C.POINTER(f), u, f]
    R.refFilterMatchesBySurfaceArea.argtypes = [u, u, u, vp, vp, vp, C.POINTER(f), f]
    R.refFilterMatchesByDenseVerify.argtypes = [u] * 5 + [C.POINTER(f), vp, vp, vp, vp] + [f] * 7
    R.refAddCurrToResiduals.argtypes = [u, u, u] + [vp] * 6 + [u, C.POINTER(f)]
    out = lambda **kw: print(json.dumps(kw), flush=True)

    # ---- sort + Kabsch filter + residual assembly ----
    for seed, n_in, n_out, noise in ((1, 40, 12, 0.002), (2, 20, 30, 0.004), (3, 60, 0, 0.0005), (4, 6, 3, 0.002)):
        pb = synth.make_filter_problem(n_pairs=7, n_inliers=n_in, n_outliers=n_out, noise=noise, seed=seed)
        P, cur = pb["P"], pb["cur"]
        rng = np.random.default_rng(seed)
        perm = np.stack([rng.permutation(128) for _ in range(P)])                       # unsorted input for the sort
        d_uns = np.take_along_axis(pb["dists"], perm, 1); i_uns = np.take_along_axis(pb["idxs"], perm[..., None], 1)
        n = pb["num"]
        for p in range(P):                                                              # keep the valid entries in the first num[p] slots
            o = np.argsort(perm[p] >= n[p], kind="stable"); d_uns[p] = d_uns[p][o]; i_uns[p] = i_uns[p][o]
        res = {}
        for who in ("ref", "ours"):
            d_keys, d_num = DevBuf(pb["keys"]), DevBuf(n)
            d_d, d_i = DevBuf(d_uns), DevBuf(i_uns)
            if who == "ref":
                rc = R.refSortKeyPointMatches(cur, 0, P, d_num.ptr, d_d.ptr, d_i.ptr)
            else:
                rc = L.bfSiftSortKeyPointMatches(cur, 0, P, d_num.ptr, d_d.ptr, d_i.ptr)
            sd, si = d_d.get(), d_i.get()
            d_nf, d_fd, d_fi = DevBuf(np.full(P, -7, np.int32)), DevBuf(np.zeros((P, 25), np.float32)), DevBuf(np.zeros((P, 25, 2), np.uint32))
            d_T, d_Ti = DevBuf(np.zeros((P, 16), np.float32)), DevBuf(np.zeros((P, 16), np.float32))
            if who == "ref":
                rc |= R.refFilterKeyPointMatches(cur, 0, P, d_keys.ptr, d_num.ptr, d_d.ptr, d_i.ptr, d_nf.ptr, d_fd.ptr, d_fi.ptr, d_T.ptr, d_Ti.ptr, f16(pb["Kinv"]), 5, 0.0004)
            else:
                rc |= L.bfSiftFilterKeyPointMatches(cur, 0, P, d_keys.ptr, d_num.ptr, d_d.ptr, d_i.ptr, d_nf.ptr, d_fd.ptr, d_fi.ptr, d_T.ptr, d_Ti.ptr, f16(pb["Kinv"]), 5, 0.0004)
            nf = d_nf.get()
            d_ent, d_eidx, d_cnt = DevBuf(np.zeros(32 * 25 * P, np.uint8)), DevBuf(np.zeros((25 * P, 2), np.uint32)), DevBuf(np.zeros(1, np.int32))
            if who == "ref":
                rc |= R.refAddCurrToResiduals(cur, 0, P, d_ent.ptr, d_eidx.ptr, d_cnt.ptr, d_nf.ptr, d_fi.ptr, d_keys.ptr, 1024, f16(pb["Kinv"]))
            else:
                rc |= L.bfSiftAddCurrToResiduals(cur, 0, P, d_ent.ptr, d_eidx.ptr, d_cnt.ptr, d_nf.ptr, d_fi.ptr, d_keys.ptr, f16(pb["Kinv"]))
            cnt = int(d_cnt.get()[0])
            ent = d_ent.get()[:32 * cnt].reshape(cnt, 32)
            res[who] = dict(rc=rc, sd=sd, si=si, nf=nf, fd=d_fd.get(), fi=d_fi.get(), T=d_T.get().reshape(P, 4, 4), Ti=d_Ti.get().reshape(P, 4, 4), cnt=cnt,
                            ent=set(bytes(e) for e in ent))
        a, b = res["ref"], res["ours"]
        pairs = [p for p in range(P) if p != cur]
        out(test="sort+filter+residuals", seed=seed, rc=[a["rc"], b["rc"]],
            sorted_dist_equal=bool(all(np.array_equal(a["sd"][p, :n[p]], b["sd"][p, :n[p]]) for p in pairs)),
            sorted_idx_equal=bool(all(np.array_equal(a["si"][p, :n[p]], b["si"][p, :n[p]]) for p in pairs)),
            num_filtered_ref=a["nf"].tolist(), num_filtered_ours=b["nf"].tolist(),
            filtered_idx_equal=[bool(a["nf"][p] == b["nf"][p] and np.array_equal(a["fi"][p, :max(a["nf"][p], 0)], b["fi"][p, :max(b["nf"][p], 0)])) for p in pairs],
            filtered_idx_same_set=[bool(set(map(tuple, a["fi"][p, :max(a["nf"][p], 0)])) == set(map(tuple, b["fi"][p, :max(b["nf"][p], 0)]))) for p in pairs],
            T_maxdiff=[float(np.abs(a["T"][p] - b["T"][p]).max()) for p in pairs if a["nf"][p] > 0 and b["nf"][p] > 0],
            Tinv_maxdiff=[float(np.abs(a["Ti"][p] - b["Ti"][p]).max()) for p in pairs if a["nf"][p] > 0 and b["nf"][p] > 0],
            residual_count=[a["cnt"], b["cnt"]], residual_entries_same_set=bool(a["ent"] == b["ent"]))

    # ---- surface area ----
    for seed in range(4):
        pb = synth.make_area_problem(seed)
        for thresh in (0.032, 0.002, 0.5):
            nf = {}
            for who in ("ref", "ours"):
                d_keys, d_num, d_idx = DevBuf(pb["keys"]), DevBuf(pb["num"]), DevBuf(pb["fidx"])
                if who == "ref":
                    rc = R.refFilterMatchesBySurfaceArea(pb["cur"], 0, pb["P"], d_keys.ptr, d_num.ptr, d_idx.ptr, f16(pb["Kinv"]), thresh)
                else:
                    rc = L.bfSiftFilterMatchesBySurfaceArea(pb["cur"], 0, pb["P"], d_keys.ptr, d_num.ptr, d_idx.ptr, f16(pb["Kinv"]), thresh, None)
                nf[who] = d_num.get().tolist()
            _, areas = orc.sift_filter_surface_area(pb["cur"], 0, pb["P"], pb["keys"], pb["num"], pb["fidx"], pb["Kinv"], thresh)
            out(test="surface_area", seed=seed, thresh=thresh, ref=nf["ref"], ours=nf["ours"], equal=nf["ref"] == nf["ours"], oracle_areas=np.round(areas, 5).tolist())

    # ---- dense verify ----
    dv = synth.make_dense_verify_problem()
    P, cur = dv["P"], dv["cur"]
    num = np.full(P, 7, np.int32)
    Tinv = np.stack([np.linalg.inv(t.astype(np.float64)).astype(np.float32) for t in dv["T"]])
    for label, opt in (("defaults", VERIFY), ("tight", dict(VERIFY, errThresh=0.02, corrThresh=0.6))):
        nf = {}
        for who in ("ref", "ours"):
            keep = []
            recs = (capi.BFCUDACachedFrame * P)()
            for r, fr in zip(recs, dv["caches"]):
                bufs = [DevBuf(fr[k].astype(np.float32)) for k in ("depth", "campos", "normals")]
                keep += bufs
                r.d_depthDownsampled, r.d_cameraposDownsampled, r.d_normalsDownsampled = bufs[0].ptr, bufs[1].ptr, bufs[2].ptr
            d_recs, d_num, d_T, d_Ti = DevBuf(np.frombuffer(bytes(recs), np.uint8)), DevBuf(num), DevBuf(dv["T"]), DevBuf(Tinv)
            a = (opt["distThresh"], opt["normalThresh"], opt["colorThresh"], opt["errThresh"], opt["corrThresh"], opt["dMin"], opt["dMax"])
            if who == "ref":
                rc = R.refFilterMatchesByDenseVerify(cur, 0, P, dv["W"], dv["H"], f16(dv["K"]), d_num.ptr, d_T.ptr, d_Ti.ptr, d_recs.ptr, *a)
            else:
                rc = L.bfSiftFilterMatchesByDenseVerify(cur, 0, P, dv["W"], dv["H"], f16(dv["K"]), d_num.ptr, d_T.ptr, d_recs.ptr, *a, None)
            nf[who] = d_num.get().tolist()
        _, stats = orc.sift_filter_dense_verify(cur, 0, P, dv["W"], dv["H"], dv["K"], num, dv["T"], dv["caches"], **opt)
        out(test="dense_verify", options=label, ref=nf["ref"], ours=nf["ours"], equal=nf["ref"] == nf["ours"], oracle_err_corr=np.round(stats, 5).tolist())


if __name__ == "__main__":
    main()
