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

/* ---- B200-native extension: the re-integration choice of TrajectoryManager::generateUpdateLists (FL/TrajectoryManager.cpp:45-108) on
 * the device.  The reference copies the whole optimised trajectory to the host every frame, converts every pose with MatrixToPose, and
 * std::sorts all frames to take the top N.  Here one launch computes, per frame, dist = |(s w, t)_integrated - (s w, t)_optimised|^2
 * ((w, t) = the SE(3) logarithm, s = rescaleRotToTrans = 2) and selects the up to topN INTEGRATED frames of largest dist > minPoseDistSqrt, in
 * descending order (ties: lower frame index first; the reference's std::sort leaves ties unspecified).  A frame whose optimised transform
 * is invalid (first entry -inf) is never selected (the reference routes it to the de-integration list).
 * d_frameState[i] != 0 <=> frame i is currently integrated.  Outputs: d_dist[numFrames], d_list[topN] (frame indices), d_count[1].
 * Asynchronous on the library stream; returns 0 or a cudaError_t. */
int bfTrajectorySelectReintegration(const float* d_optimizedTransforms, const float* d_integratedTransforms, const int* d_frameState,
                                    unsigned int numFrames, unsigned int topN, float minPoseDistSqrt, float rescaleRotToTrans,
                                    float* d_dist, int* d_list, int* d_count);

#ifdef __cplusplus
}
#endif
#endif /* BF_BUNDLER_H */
