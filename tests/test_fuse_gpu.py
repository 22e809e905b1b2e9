"""bfSiftFuseToGlobal (csrc/sift_fuse.cu: SIFTImageManager::fuseToGlobal on the device, one launch) on the GPU against oracle/fuse_oracle.c,
bit for bit -- fused key points (position, scale, depth), descriptors, count, order."""
import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from bundlefusion_b200 import synth
from oracle import oracle as orc
from tests._cudart import DevBuf, device_count
from tests.test_fuse_oracle import run_fuse

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    if device_count() == 0:
        pytest.skip("no CUDA device")
    import ctypes as C
    L = capi.lib()
    vp, u = C.c_void_p, C.c_uint
    L.bfSiftFuseToGlobal.argtypes = [vp, vp, vp, vp, u, vp, vp, vp, u, C.POINTER(C.c_float), u, vp, vp, vp, u, vp]
    L.bfSetStream(None)
    return L


@pytest.mark.parametrize("seed,n_images,n_points,stride", [(0, 6, 140, 256), (1, 11, 150, 1024), (2, 3, 40, 64), (5, 11, 160, 1024)])
def test_fuse_to_global_bit_exact(gpu, seed, n_images, n_points, stride):
    pb = synth.make_fuse_problem(seed=seed, n_images=n_images, n_points=n_points, key_stride=stride)
    ko, do = orc.sift_fuse_to_global(pb["corr"], pb["keyIdx"], pb["transforms"], pb["keys"], pb["descs"], pb["numKeys"], pb["keyStride"], pb["K"])
    kg, dg, st = run_fuse(gpu, pb, to_dev=DevBuf, from_dev=lambda b: b.get())
    assert st == 0 and len(kg) == len(ko) > 10
    assert np.array_equal(kg.view(np.uint32), ko.view(np.uint32)) and np.array_equal(dg, do)


def test_fuse_to_global_rejects_lists_beyond_its_capacity(gpu):
    pb = synth.make_fuse_problem(seed=5, n_images=11, n_points=400, key_stride=1024)      # > 4096 correspondences: more than any chunk can hold (25 x 55)
    assert len(pb["corr"]) > 4096
    with pytest.raises(AssertionError):
        run_fuse(gpu, pb, to_dev=DevBuf, from_dev=lambda b: b.get())


def test_fuse_to_global_empty_and_capped(gpu):
    pb = synth.make_fuse_problem(seed=3, n_images=4, n_points=60, key_stride=128)
    ko, do = orc.sift_fuse_to_global(pb["corr"], pb["keyIdx"], pb["transforms"], pb["keys"], pb["descs"], pb["numKeys"], pb["keyStride"], pb["K"])
    kg, dg, _ = run_fuse(gpu, pb, max_keys=9, to_dev=DevBuf, from_dev=lambda b: b.get())
    assert np.array_equal(kg.view(np.uint32), ko[:9].view(np.uint32)) and np.array_equal(dg, do[:9])
    empty = dict(pb, corr=pb["corr"][:0], keyIdx=pb["keyIdx"][:0])
    kg, dg, _ = run_fuse(gpu, empty, to_dev=DevBuf, from_dev=lambda b: b.get())
    assert len(kg) == 0
