/*
 * bf_host.h -- host-side helpers of the drop-in library (C-ABI).
 *
 * These are the pieces of FriedLiver's HOST code that sit directly on top of the kernel stubs and
 * that a frame loop needs at device rate: the 4x4 pose inverse the reference computes on the host
 * before every (de)integration (FL/DepthSensing/CUDASceneRepHashSDF.h:128-134 ->
 * FL/SiftGPU/cuda_SimpleMatrixUtil.h:980-1100) and the re-integration batch of
 * FL/DepthSensing/DepthSensing.cpp:854-902 (up to s_maxFrameFixes x {deIntegrate(old pose);
 * integrate(new pose)} followed by garbageCollect), replayed from a command list so that the host
 * issues a whole frame's TSDF work without returning to the caller.
 */
#ifndef BF_HOST_H
#define BF_HOST_H

#include "bf_tsdf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* general 4x4 inverse in fp32 as the reference's host forms it: float4x4::getInverse (cuda_SimpleMatrixUtil.h:980-1100) = mat4f::getInverse (mLib
 * core-math/matrix4x4.h:587-710), every product and sum in their order, bit for bit (tests/test_mat4_inverse_reference.py); row-major in/out */
void bfMat4Inverse(const float* m16, float* out16);

enum { BF_TSDF_OP_INTEGRATE = 0, BF_TSDF_OP_DEINTEGRATE = 1, BF_TSDF_OP_GARBAGE_COLLECT = 2 };

/* one TSDF operation of the frame loop */
typedef struct BFTsdfOp {
    int32_t kind;        /* BF_TSDF_OP_*                                      */
    int32_t frame;       /* index into the frame arrays (ignored for GC)      */
    float   pose[16];    /* camera-to-world, row-major (ignored for GC)       */
} BFTsdfOp;

/* Replays `numOps` operations on the library stream, asynchronously (no host sync):
 * integrate / deIntegrate exactly as CUDASceneRepHashSDF::integrate / ::deIntegrate, GC as ::garbageCollect.
 * d_depthFrames[i] / d_colorFrames[i] are DEVICE pointers to frame i (W*H float / W*H uchar4).
 * hashParams is updated with the last pose (as setLastRigidTransform would leave it). */
int bfTsdfRunOps(BFHashDataStruct* hashData, BFHashParams* hashParams, const BFDepthCameraParams* cam,
                 const BFTsdfOp* ops, int numOps,
                 const float* const* d_depthFrames, const uint8_t* const* d_colorFrames);

#ifdef __cplusplus
}
#endif
#endif
