"""Generates tests/golden/mesh_reference_host.npz: the REFERENCE's mesh clean-up and PLY writer (mLib MeshDataf::mergeCloseVertices / removeDuplicateFaces /
applyTransform, MeshIOf::saveToFile, driven by the statements of CUDAMarchingCubesHashSDF::saveMesh; oracle/_ref/libref_mesh_host.so built by oracle/build_ref.py
build_mesh_host) on the triangle soups of tests/test_mesh_reference_host.py: merged vertices, colours, faces and the PLY file bytes.

    python oracle/build_ref.py && python scripts/make_golden_mesh_host.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_mesh_reference_host import GOLDEN, TRANSFORM, reference_save, soups              # noqa: E402


def main():
    out = {}
    d = tempfile.mkdtemp()
    S = soups()
    for name, transform in (("a", None), ("b", None), ("b", TRANSFORM)):
        key = name + ("_t" if transform is not None else "")
        path = os.path.join(d, key + ".ply")
        p, c, f = reference_save(S[name], transform, path)
        out["soup_" + name] = S[name]
        out["pos_" + key], out["col_" + key], out["faces_" + key] = p, c, f
        out["ply_" + key] = np.frombuffer(open(path, "rb").read(), np.uint8)
        print(key, len(S[name]), "triangles ->", len(p), "vertices,", len(f), "faces,", os.path.getsize(path), "bytes")
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
