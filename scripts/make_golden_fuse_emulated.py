"""Generates tests/golden/fuse_reference_emulated.npz: the fused keyframe (keys, descriptors) the REFERENCE's SIFTImageManager::fuseToGlobal produces
(FL/SiftGPU/SIFTImageManager.cpp:366-476, compiled with the manager class against the CUDA emulation: oracle/_ref/libref_fuse_emulated.so, oracle/build_ref.py
build_fuse_emulated) on the solved chunks of tests/test_fuse_reference_emulated.py.

    python oracle/build_ref.py && python scripts/make_golden_fuse_emulated.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bundlefusion_b200 import synth                                                            # noqa: E402
from tests.test_fuse_reference_emulated import CASES, GOLDEN, reference_filter_frames, reference_fuse                   # noqa: E402


def main():
    out = {}
    for c, kw in enumerate(CASES):
        k, d = reference_fuse(synth.make_fuse_problem(**kw))
        out[f"keys_{c}"], out[f"descs_{c}"] = k, d
        print(kw, "->", len(k), "fused keys")
    out["ff_last"], out["ff_valid"] = reference_filter_frames()
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
