"""Generates tests/golden/mat4_inverse_reference.npz: the reference's two host 4x4 inverses -- float4x4::getInverse (FL/SiftGPU/cuda_SimpleMatrixUtil.h, in
oracle/_ref/libref_kabsch_host.so) and mLib's mat4f::getInverse (in oracle/_ref/libref_mesh_host.so), both compiled by g++ from /root/reference by oracle/build_ref.py --
on the matrices of tests/test_mat4_inverse_reference.py.

    python oracle/build_ref.py && python scripts/make_golden_mat4_inverse.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_mat4_inverse_reference import GOLDEN, matrices, ref_inverses              # noqa: E402


def main():
    M = matrices()
    a, b = ref_inverses(M)
    np.savez_compressed(GOLDEN, matrices=M, float4x4=a, mat4f=b)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes;", len(M), "matrices; the two reference classes agree:", bool(np.array_equal(a.view(np.uint32), b.view(np.uint32))))


if __name__ == "__main__":
    main()
