"""ctypes view of the C-ABI declared in include/*.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (``make -C
bundlefusion_b200/csrc``).  There is NO CPU fallback: ``lib()`` raises if the library is
missing, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbundlefusion_b200.so")

BF_SDF_BLOCK_SIZE = 8
BF_SDF_BLOCK_VOXELS = 512
BF_HASH_BUCKET_SIZE = 4
BF_LOCK_ENTRY = -1
BF_FREE_ENTRY = -2


class BFFloat4x4(C.Structure):
    _fields_ = [("m", C.c_float * 16)]


class BFHashEntry(C.Structure):
    _fields_ = [("pos", C.c_int32 * 3), ("ptr", C.c_int32), ("offset", C.c_uint32)]


class BFVoxel(C.Structure):
    _fields_ = [("sdf", C.c_float), ("weight", C.c_float), ("color", C.c_uint8 * 4)]


class _Dummy2(C.Structure):
    _pack_ = 8
    _fields_ = [("v", C.c_uint32 * 2)]


class BFHashParams(C.Structure):
    _fields_ = [
        ("m_rigidTransform", BFFloat4x4),
        ("m_rigidTransformInverse", BFFloat4x4),
        ("m_hashNumBuckets", C.c_uint32),
        ("m_hashBucketSize", C.c_uint32),
        ("m_hashMaxCollisionLinkedListSize", C.c_uint32),
        ("m_numSDFBlocks", C.c_uint32),
        ("m_SDFBlockSize", C.c_int32),
        ("m_virtualVoxelSize", C.c_float),
        ("m_numOccupiedBlocks", C.c_uint32),
        ("m_maxIntegrationDistance", C.c_float),
        ("m_truncScale", C.c_float),
        ("m_truncation", C.c_float),
        ("m_integrationWeightSample", C.c_uint32),
        ("m_integrationWeightMax", C.c_uint32),
        ("m_streamingVoxelExtents", C.c_float * 3),
        ("m_streamingGridDimensions", C.c_int32 * 3),
        ("m_streamingMinGridPos", C.c_int32 * 3),
        ("m_streamingInitialChunkListSize", C.c_uint32),
        ("m_dummy", C.c_uint64),  # uint2, 8-byte aligned
    ]


class BFDepthCameraParams(C.Structure):
    _fields_ = [
        ("fx", C.c_float), ("fy", C.c_float), ("mx", C.c_float), ("my", C.c_float),
        ("m_imageWidth", C.c_uint32), ("m_imageHeight", C.c_uint32),
        ("m_sensorDepthWorldMin", C.c_float), ("m_sensorDepthWorldMax", C.c_float),
    ]


class BFDepthCameraData(C.Structure):
    _fields_ = [("d_depthData", C.c_void_p), ("d_colorData", C.c_void_p)]


class BFHashDataStruct(C.Structure):
    _fields_ = [
        ("d_heap", C.c_void_p),
        ("d_heapCounter", C.c_void_p),
        ("d_hashDecision", C.c_void_p),
        ("d_hashDecisionPrefix", C.c_void_p),
        ("d_hash", C.c_void_p),
        ("d_hashCompactified", C.c_void_p),
        ("d_hashCompactifiedCounter", C.c_void_p),
        ("d_SDFBlocks", C.c_void_p),
        ("d_hashBucketMutex", C.c_void_p),
        ("m_bIsOnGPU", C.c_uint8),
    ]


assert C.sizeof(BFHashParams) == 224, C.sizeof(BFHashParams)
assert BFHashParams.m_dummy.offset == 216
assert C.sizeof(BFHashEntry) == 20 and C.sizeof(BFVoxel) == 12
assert C.sizeof(BFDepthCameraParams) == 32 and C.sizeof(BFHashDataStruct) == 80

# every symbol include/bf_tsdf.h declares (the "library exports what the header says" test walks this)
TSDF_SYMBOLS = [
    "updateConstantHashParams", "updateConstantDepthCameraParams", "bindInputDepthColorTextures",
    "resetCUDA", "resetHashBucketMutexCUDA", "allocCUDA", "fillDecisionArrayCUDA", "compactifyHashCUDA",
    "compactifyHashAllInOneCUDA", "integrateDepthMapCUDA", "deIntegrateDepthMapCUDA", "starveVoxelsKernelCUDA",
    "garbageCollectIdentifyCUDA", "garbageCollectFreeCUDA",
    "bfSetStream", "bfGetStream", "bfGetLastErrorString", "bfTsdfAuxBytes", "bfTsdfReset", "bfTsdfIntegrateFrame",
    "bfTsdfGarbageCollect", "bfTsdfGetHeapFreeCount", "bfTsdfGetNumOccupiedBlocks", "bfTsdfGetLastFrameStats",
    "bfTsdfReleaseAux",
]

HOST_SYMBOLS = ["bfMat4Inverse", "bfTsdfRunOps"]


class BFTsdfOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("frame", C.c_int32), ("pose", C.c_float * 16)]


BF_TSDF_OP_INTEGRATE, BF_TSDF_OP_DEINTEGRATE, BF_TSDF_OP_GARBAGE_COLLECT = 0, 1, 2

_lib = None


def lib() -> C.CDLL:
    """Load libbundlefusion_b200.so (raises RuntimeError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C bundlefusion_b200/csrc`). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp = C.c_void_p
    # reference-named stubs
    L.updateConstantHashParams.argtypes = [P(BFHashParams)]
    L.updateConstantDepthCameraParams.argtypes = [P(BFDepthCameraParams)]
    L.bindInputDepthColorTextures.argtypes = [P(BFDepthCameraData), C.c_uint, C.c_uint]
    for name in ("resetCUDA", "resetHashBucketMutexCUDA", "fillDecisionArrayCUDA", "compactifyHashCUDA",
                 "starveVoxelsKernelCUDA", "garbageCollectIdentifyCUDA", "garbageCollectFreeCUDA"):
        getattr(L, name).argtypes = [P(BFHashDataStruct), P(BFHashParams)]
        getattr(L, name).restype = None
    L.allocCUDA.argtypes = [P(BFHashDataStruct), P(BFHashParams), P(BFDepthCameraData), P(BFDepthCameraParams), vp]
    L.allocCUDA.restype = None
    L.compactifyHashAllInOneCUDA.argtypes = [P(BFHashDataStruct), P(BFHashParams)]
    L.compactifyHashAllInOneCUDA.restype = C.c_uint
    for name in ("integrateDepthMapCUDA", "deIntegrateDepthMapCUDA"):
        getattr(L, name).argtypes = [P(BFHashDataStruct), P(BFHashParams), P(BFDepthCameraData), P(BFDepthCameraParams)]
        getattr(L, name).restype = None
    # extension
    L.bfSetStream.argtypes = [vp]
    L.bfSetStream.restype = None
    L.bfGetStream.restype = vp
    L.bfGetLastErrorString.restype = C.c_char_p
    L.bfTsdfAuxBytes.argtypes = [P(BFHashParams)]
    L.bfTsdfAuxBytes.restype = C.c_size_t
    L.bfTsdfReset.argtypes = [P(BFHashDataStruct), P(BFHashParams)]
    L.bfTsdfIntegrateFrame.argtypes = [P(BFHashDataStruct), P(BFHashParams), P(BFDepthCameraData), P(BFDepthCameraParams), C.c_int]
    L.bfTsdfGarbageCollect.argtypes = [P(BFHashDataStruct), P(BFHashParams)]
    L.bfTsdfGetHeapFreeCount.argtypes = [P(BFHashDataStruct), P(C.c_uint)]
    L.bfTsdfGetNumOccupiedBlocks.argtypes = [P(BFHashDataStruct), P(C.c_uint)]
    L.bfTsdfGetLastFrameStats.argtypes = [P(BFHashDataStruct), C.c_ulonglong * 4]
    L.bfTsdfReleaseAux.argtypes = [P(BFHashDataStruct)]
    L.bfMat4Inverse.argtypes = [P(C.c_float), P(C.c_float)]
    L.bfMat4Inverse.restype = None
    L.bfTsdfRunOps.argtypes = [P(BFHashDataStruct), P(BFHashParams), P(BFDepthCameraParams), P(BFTsdfOp), C.c_int, P(vp), P(vp)]
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().bfGetLastErrorString().decode()
        raise RuntimeError(f"bundlefusion_b200: {what} failed with CUDA error {rc}: {msg}")
