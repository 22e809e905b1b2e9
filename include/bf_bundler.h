/*
 * bf_bundler.h -- C-ABI of the trajectory glue between bundling and reconstruction.
 *
 * Re-exports, by NAME and machine-level signature, the `extern "C"` stubs FL/OnlineBundler.cpp:18-27 declares and FL/OnlineBundler.cu
 * defines (float4x4 = 16 row-major floats; C++ references are pointers at the ABI level):
 *   computeSiftTransformCU     FL/OnlineBundler.cu:6-69     pose used to integrate the CURRENT frame: last known pose chained with the
 *                                                            Kabsch transform of the most recent earlier frame of the chunk it matched
 *   updateTrajectoryCU         FL/OnlineBundler.cu:71-110   complete[k] = global[k / submap] * local[k / submap][k % submap], or -inf if invalid
 *   initNextGlobalTransformCU  FL/OnlineBundler.cu:114-140  next keyframe's initial global pose
 * (FL/ = /root/reference/FriedLiver/Source/.)  All three launch asynchronously on the library stream (bfSetStream).
 */
#ifndef BF_BUNDLER_H
#define BF_BUNDLER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void computeSiftTransformCU(const float* d_currFilteredTransformsInv, const int* d_currNumFilteredMatchesPerImagePair,
                            const float* d_completeTrajectory, unsigned int lastValidCompleteTransform,
                            float* d_siftTrajectory, unsigned int curFrameIndexAll, unsigned int curFrameIndex, float* d_currIntegrateTrans);
void initNextGlobalTransformCU(float* d_globalTrajectory, unsigned int numGlobalTransforms, unsigned int initGlobalIdx,
                               float* d_localTrajectories, unsigned int lastValidLocal, unsigned int numLocalTransformsPerTrajectory);
void updateTrajectoryCU(const float* d_globalTrajectory, unsigned int numGlobalTransforms, float* d_completeTrajectory, unsigned int numCompleteTransforms,
                        const float* d_localTrajectories, unsigned int numLocalTransformsPerTrajectory, unsigned int numLocalTrajectories,
                        int* d_imageInvalidateList);

#ifdef __cplusplus
}
#endif
#endif /* BF_BUNDLER_H */
