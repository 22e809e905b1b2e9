"""Generates tests/golden/kabsch_reference_host.npz: outputs of the REFERENCE's own host-callable Kabsch code (kabsch, filterKeyPointMatches
of FL/SiftGPU/cuda_kabsch.h with the SVD of cuda_svd3.h, MYEIGEN::eigenSystem of cuda_SVD.h), run on the CPU through
oracle/_ref/libref_kabsch_host.so (built by oracle/build_ref.py from the sources under /root/reference), on seeded inputs that
tests/test_kabsch_reference_host.py regenerates.  Needs /root/reference only through that prebuilt library.

    python oracle/build_ref.py && python scripts/make_golden_kabsch_host.py

The reference's host rsqrt is the SSE estimate _mm_rsqrt_ss, whose low bits depend on the CPU vendor; a few estimates are stored as a
canary so that the test can tell when it runs on a CPU with another table."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_kabsch_reference_host import CANARY_INPUTS, eigen_cases, filter_cases, kabsch_cases    # noqa: E402
from oracle import oracle as orc                                                                        # noqa: E402


def main():
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_kabsch_host.so"))
    vp = C.c_void_p
    R.refHostFilterKeyPointMatches.argtypes = [vp, vp, vp, C.c_uint, vp, vp, C.c_uint, C.c_float]; R.refHostFilterKeyPointMatches.restype = C.c_uint
    R.refHostKabsch.argtypes = [vp, vp, C.c_uint, vp, vp]
    R.refHostEigenSystem.argtypes = [vp, vp, vp]
    out = {}
    kT, kE = [], []
    for src, tgt in kabsch_cases():
        T = np.zeros(16, np.float32); e = np.zeros(3, np.float32)
        R.refHostKabsch(src.ctypes.data, tgt.ctypes.data, len(src), T.ctypes.data, e.ctypes.data)
        kT.append(T); kE.append(e)
    out["kabsch_T"], out["kabsch_evs"] = np.stack(kT), np.stack(kE)
    fc, fi, fd, fT = [], [], [], []
    for keys, idx, dist, n, Ki in filter_cases():
        i, d, T = idx.copy(), dist.copy(), np.zeros(16, np.float32)
        c = R.refHostFilterKeyPointMatches(keys.ctypes.data, i.ctypes.data, d.ctypes.data, n, T.ctypes.data, Ki.ctypes.data, 5, 0.0004)
        fc.append(c); fi.append(i[:25].copy()); fd.append(d[:25].copy()); fT.append(T)
    out["filter_count"], out["filter_idx"], out["filter_dist"], out["filter_T"] = np.array(fc, np.int32), np.stack(fi), np.stack(fd), np.stack(fT)
    ok, ev, vec = [], [], []
    for M in eigen_cases():
        e = np.zeros(3, np.float32); v = np.zeros(9, np.float32)
        ok.append(R.refHostEigenSystem(M.ctypes.data, e.ctypes.data, v.ctypes.data)); ev.append(e); vec.append(v)
    out["eig_ok"], out["eig_vals"], out["eig_vecs"] = np.array(ok, np.int32), np.stack(ev), np.stack(vec)
    L = orc.lib(); L.orc_rsqrt_host_estimate.argtypes = [C.c_float]; L.orc_rsqrt_host_estimate.restype = C.c_float
    out["canary"] = np.array([L.orc_rsqrt_host_estimate(float(x)) for x in CANARY_INPUTS], np.float32)
    path = os.path.join(ROOT, "tests", "golden", "kabsch_reference_host.npz")
    np.savez_compressed(path, **out)
    print("written", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
