/*
 * bf_frameloop.h -- C-ABI of the per-frame sequencing object: ONE call per sensor frame runs what FriedLiver's frame callback runs.
 *
 * The reference's frame loop (single-threaded build, no RUN_MULTITHREADED) is
 *   OnD3D11FrameRender                       FL/DepthSensing/DepthSensing.cpp:966-1129
 *     CUDAImageManager::process              FL/CUDAImageManager.cpp:22-158      frame ingest (erode, filter, resample), frame store
 *     OnlineBundler::processInput            FL/OnlineBundler.cpp:167-227        SIFT detect, dense cache, match + filters vs the chunk, SIFT pose
 *       Bundler::detectFeatures / storeCachedFrame / matchAndFilter              FL/Bundler.cpp:91-249
 *       OnlineBundler::computeCurrentSiftTransform / prepareLocalSolve           FL/OnlineBundler.cpp:118-165
 *     reintegrate()                          FL/DepthSensing/DepthSensing.cpp:854-902   <= s_maxFrameFixes re-integrations + GC
 *     integrate(current frame, SIFT pose) + TrajectoryManager::addFrame          :1033-1061
 *     OnlineBundler::process                 FL/OnlineBundler.cpp:410-416
 *       optimizeLocal  -> Bundler::optimize -> SBA::align -> CUDASolverBundling::solve (+ VerifyTrajectoryCU)   :229-262, FL/Bundler.cpp:251-283
 *       processGlobal  -> SIFTImageManager::fuseToGlobal, global matchAndFilter, re-validation                  :271-358
 *       optimizeGlobal -> global solve (+ max-residual removal), updateTrajectoryCU, TrajectoryManager update   :369-408
 * (FL/ = /root/reference/FriedLiver/Source/).  bfFrameLoopStep is that sequence behind one call: every stage is a kernel sequence of this
 * library (bf_ingest.h, bf_sift.h, bf_cache.h, bf_solver.h, bf_bundler.h, bf_tsdf.h / bf_host.h); the host state machines (Bundler /
 * OnlineBundler / SIFTImageManager bookkeeping, TrajectoryManager) are C++ inside the library (csrc/frame_loop.cu, csrc/trajectory_host.cu).
 * The host looks at the device three times per frame (key-point count after detection; match verdict + correspondence count; and, on chunk
 * boundaries, the solve verdicts) -- the reference does so dozens of times (every filter stage copies counts back).
 *
 * What is NOT here: sensors / .sens decoding, rendering (ray cast, D3D11), the multi-threaded hand-over between the reconstruction and
 * bundling threads (the call runs the reference's single-threaded order), end-of-sequence extras (USE_GLOBAL_DENSE_AT_END, marching cubes).
 * Threading: one BFFrameLoop per process / device, calls from one thread (the stages share the library's stream and workspaces).
 */
#ifndef BF_FRAMELOOP_H
#define BF_FRAMELOOP_H

#include <stddef.h>
#include <stdint.h>

#include "bf_tsdf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* GlobalAppState / GlobalBundlingState values the loop uses (FriedLiver/zParametersDefault.txt, zParametersBundlingDefault.txt) */
typedef struct BFFrameLoopParams {
    uint32_t depthWidth, depthHeight, colorWidth, colorHeight;          /* sensor images */
    uint32_t integrationWidth, integrationHeight;                      /* s_integrationWidth / Height */
    uint32_t siftWidth, siftHeight;                                     /* s_widthSIFT / s_heightSIFT */
    float    depthIntrinsics[16], colorIntrinsics[16];                  /* sensor calibration, 4x4 row-major */
    uint32_t submapSize;                                                /* s_submapSize (10) */
    uint32_t maxNumImages;                                              /* s_maxNumImages: keyframes (1200) */
    uint32_t maxNumKeysPerImage;                                        /* s_maxNumKeysPerImage (1024) */
    uint32_t maxNumFrames;                                              /* frames the device frame store / trajectories hold (<= maxNumImages * submapSize) */
    uint32_t maxGlobalResiduals;                                        /* capacity of the keyframe correspondence list; 0: 25 N (N - 1) / 2 as the reference, capped at 16 M */
    uint32_t numLocalNonLinIterations, numLocalLinIterations;           /* 2, 100 */
    uint32_t numGlobalNonLinIterations, numGlobalLinIterations;         /* 3, 150 */
    uint32_t numOptPerResidualRemoval;                                  /* 1 */
    float    sensorDepthMin, sensorDepthMax;                            /* 0.1, 4.0 */
    float    minKeyScale;                                               /* 3.0 */
    int32_t  featureCountThreshold;                                     /* 150 (FL/Bundler.cpp:61) */
    float    siftMatchThresh, siftMatchRatioMaxLocal, siftMatchRatioMaxGlobal;   /* 0.7, 0.8, 0.8 */
    uint32_t minNumMatchesLocal, minNumMatchesGlobal;                   /* 5, 5 */
    float    maxKabschResidual2;                                        /* 0.0004 */
    float    surfAreaPcaThresh;                                         /* 0.032 */
    float    projCorrDistThres, projCorrNormalThres, projCorrColorThresh;        /* 0.15, 0.97, 0.1 */
    float    verifySiftErrThresh, verifySiftCorrThresh;                 /* 0.075, 0.02 */
    float    verifyOptErrThresh, verifyOptCorrThresh;                   /* 0.05, 0.001 */
    float    optMaxResThresh;                                           /* 0.08 */
    int32_t  useLocalVerify, useLocalDense, useComprehensiveFrameInvalidation;   /* 1, 1, 1 */
    uint32_t downsampledWidth, downsampledHeight;                       /* 80, 60 */
    float    colorDownSigma, depthDownSigmaD, depthDownSigmaR;          /* 2.5, 1.0, 0.05 */
    int32_t  erodeSIFTdepth, depthFilter; float depthSigmaD, depthSigmaR;        /* 1, 1, 2.0, 0.05 */
    uint32_t maxFrameFixes, topNActive; float minPoseDistSqrt;          /* 10, 30, 0.0 */
    int32_t  reconstructionEnabled;                                     /* s_reconstructionEnabled (1) */
    BFHashParams hash;                                                  /* voxel hash parameters (poses are set per operation) */
    float    renderDepthMin, renderDepthMax;                            /* frustum test range (0.1, 4.0) */
} BFFrameLoopParams;

/* fills *p with the defaults of zParametersDefault.txt / zParametersBundlingDefault.txt for a width x height sensor with the synthetic pinhole
 * of SURVEY.md section 8d (fx = fy = 525 * width / 640, principal point at the image centre); integration at sensor resolution */
void bfFrameLoopDefaultParams(BFFrameLoopParams* p, uint32_t width, uint32_t height);

typedef struct BFFrameLoopStatus {
    uint32_t frame;                  /* index of the frame just processed */
    int32_t  validTransform;         /* OnlineBundler::getCurrentIntegrationFrame: the frame got a SIFT pose and was integrated */
    int32_t  globalTrackingLost;
    float    transform[16];          /* the pose it was integrated with (all -inf when invalid) */
    uint32_t numKeyPoints;           /* features detected in the frame */
    int32_t  lastMatchedFrame;       /* chunk-local index of the last frame it matched, -1 if none */
    uint32_t numLocalCorrespondences;
    uint32_t numReintegrated;        /* re-integration operations executed this frame */
    int32_t  localSolved;            /* chunk index solved this frame, -1 if none;  localValid: the solve + verification accepted it */
    int32_t  localValid;
    uint32_t numKeyframes;           /* keyframes in the global manager */
    uint32_t numGlobalCorrespondences;
    int32_t  globalSolved;           /* a global solve ran */
    int32_t  globalRemoved;          /* it removed a max-residual image pair */
    uint32_t numOptimizedFrames;     /* frames the optimised trajectory covers */
} BFFrameLoopStatus;

typedef struct BFFrameLoop BFFrameLoop;

/* allocates every device buffer of the loop (frame store, SIFT managers, caches, solvers, trajectories, voxel hash); returns 0 or a cudaError_t */
int bfFrameLoopCreate(const BFFrameLoopParams* params, BFFrameLoop** out);
void bfFrameLoopDestroy(BFFrameLoop* loop);

/* One sensor frame through the whole loop.  depth: float [depthHeight][depthWidth], -inf = invalid (FL/SensorDataReader.cpp:104-106); color:
 * uchar4 [colorHeight][colorWidth].  onHost != 0: the pointers are (pinned) host memory and are copied to the device inside the call, as
 * CUDAImageManager::process uploads on arrival; else they are device pointers.  Returns 0 or a cudaError_t; *status (optional) is filled. */
int bfFrameLoopStep(BFFrameLoop* loop, const float* depth, const uint8_t* color, int onHost, BFFrameLoopStatus* status);

/* The same step with look-ahead: (nextDepth, nextColor) is the frame the NEXT call will pass as (depth, color) -- a recorded stream knows it, a live sensor
 * driver knows it once the following frame has arrived.  While this frame is matched, filtered, solved and fused, the work of the next frame that depends
 * on nothing but the frame itself (upload, CUDAImageManager::process, SIFT detection, dense cache: FL/CUDAImageManager.cpp:22-158, FL/Bundler.cpp:91-101,
 * FL/CUDACache.cpp:45-86) is queued on a third stream of the loop -- the overlap the reference gets from its reconstruction and bundling threads
 * (RUN_MULTITHREADED: the image manager processes frame f + 1 while the bundler works on frame f).  Results are identical to bfFrameLoopStep: the same kernels
 * on the same inputs, their destinations (frame-store slot, key / descriptor slot and cache slot of the chunk the frame will belong to) being fixed by
 * host state one frame early.  The next call MUST pass the announced pointers (else cudaErrorInvalidValue); the buffers must stay valid and unchanged until
 * that call returns.  nextDepth == NULL: plain bfFrameLoopStep.  Look-ahead is suspended while the stage profile is on. */
int bfFrameLoopStepAhead(BFFrameLoop* loop, const float* depth, const uint8_t* color, const float* nextDepth, const uint8_t* nextColor, int onHost, BFFrameLoopStatus* status);

/* After the last frame: the reference keeps calling processInput / process with no new frame (FL/OnlineBundler.cpp:170-197) so that the last,
 * partial chunk is solved and re-integration drains.  One such turn per call. */
int bfFrameLoopStepPastEnd(BFFrameLoop* loop, BFFrameLoopStatus* status);

/* optimised camera-to-world poses of the first n frames (host, [n][16]; invalid frames all -inf); returns the number written */
unsigned int bfFrameLoopGetTrajectory(BFFrameLoop* loop, float* h_out, unsigned int maxFrames);
/* the voxel hash the loop integrates into (caller may read it; bf_tsdf.h) */
const BFHashDataStruct* bfFrameLoopGetHashData(const BFFrameLoop* loop);
const BFHashParams* bfFrameLoopGetHashParams(const BFFrameLoop* loop);
/* counters for measurement: out[0] frames, out[1] integrations, out[2] re-integrations, out[3] local solves, out[4] global solves,
 * out[5] PCG iterations (global, last solve), out[6] host synchronisations, out[7] keyframes */
void bfFrameLoopGetCounters(const BFFrameLoop* loop, unsigned long long out[8]);

/* Two streams: with overlap on, the re-integration + integration of frame f are queued on a second, lower-priority stream of the loop and run beside
 * the solves of frame f and the feature work of frame f + 1 -- what the reference's reconstruction and bundling threads do -- with the data
 * dependencies of the single-threaded order kept by events (same results).  Off by default (everything on the library stream).
 * bfFrameLoopJoin makes the library stream wait (asynchronously) for the reconstruction stream: call it before recording an event that should
 * cover the fused model, or before reading the voxel hash on the library stream (bfFrameLoopGetHashData does it itself).  Returns the previous setting. */
int bfFrameLoopSetOverlap(BFFrameLoop* loop, int enable);
int bfFrameLoopJoin(BFFrameLoop* loop);

/* Stage profile (measurement only; adds one host synchronisation per step while on).  Stages of a step, in order:
 *   0 upload + ingest      1 SIFT detection      2 dense cache (+ wait for the key-point count)      3 match + filters + SIFT pose
 *   4 re-integration + GC + integration      5 local solve (+ verification)      6 fuse to keyframe + keyframe matching      7 global solve + trajectory update
 * bfFrameLoopGetStageTimes: milliseconds on the device time line summed per stage since bfFrameLoopSetProfiling(loop, 1); returns the steps covered. */
#define BF_FRAMELOOP_STAGES 8
int bfFrameLoopSetProfiling(BFFrameLoop* loop, int enable);
unsigned long long bfFrameLoopGetStageTimes(const BFFrameLoop* loop, double outMs[BF_FRAMELOOP_STAGES]);

#ifdef __cplusplus
}
#endif
#endif /* BF_FRAMELOOP_H */
