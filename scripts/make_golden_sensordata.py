"""Generates tests/golden/sens_reference_sensordata.npz: two `.sens` files WRITTEN by the reference's own container class (ml::SensorData::initDefault / addFrame /
saveToFile, external/mLib/include/ext-depthcamera/sensorData.h, compiled by oracle/build_ref.py build_sensordata_host into oracle/_ref/libref_sensordata_host.so) from
the sequence of tests/test_sens_reference_sensordata.py: raw colour with raw depth, raw colour with zlib depth (the reference's stb deflate).

    python oracle/build_ref.py && python scripts/make_golden_sensordata.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_sens_reference_sensordata import GOLDEN, RefSensorData, sequence              # noqa: E402


def main():
    R = RefSensorData()
    K, rgb, depth, poses, ts = sequence()
    d = tempfile.mkdtemp()
    out = {}
    for zl in (0, 1):
        p = os.path.join(d, f"r{zl}.sens")
        R.write(p, K, rgb, depth, poses, ts, zl)
        out[f"file_depth{zl}"] = np.frombuffer(open(p, "rb").read(), np.uint8)
        print(p, os.path.getsize(p), "bytes")
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
