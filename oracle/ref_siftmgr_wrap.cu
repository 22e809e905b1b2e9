// ref_siftmgr_wrap.cu -- TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own match-manager kernels
// (FL/SiftGPU/SIFTImageManager.cu: sort, Kabsch filter, surface-area filter, dense verification, residual assembly), for the
// compat build oracle/build_ref.py makes in /tmp (-> oracle/_ref/libref_siftmgr.so).  This file contains no reference code: it
// includes the patched scratch copy as one translation unit and launches the reference's __global__ functions with the grids its
// SIFTImageManager methods use (file:line cited per call), on raw device pointers instead of the class's members.
#include "SiftGPU/SIFTImageManager.cu"

#define REF_API extern "C" __attribute__((visibility("default")))

static float4x4 to_m(const float* m) { float4x4 r; for (int i = 0; i < 16; ++i) r.entries[i] = m[i]; return r; }

// SIFTImageManager::SortKeyPointMatchesCU, SIFTImageManager.cu:143-160
REF_API int refSortKeyPointMatches(unsigned curFrame, unsigned startFrame, unsigned numFrames, const int* d_num, float* d_dist, uint2* d_idx) {
    if (numFrames <= startFrame) return 0;
    SortKeyPointMatchesCU_Kernel<<<dim3(numFrames - startFrame), dim3(SORT_NUM_BLOCK_THREADS_X)>>>(curFrame, startFrame, d_num, d_dist, d_idx);
    return (int)cudaDeviceSynchronize();
}

// SIFTImageManager::FilterKeyPointMatchesCU, SIFTImageManager.cu:265-292
REF_API int refFilterKeyPointMatches(unsigned curFrame, unsigned startFrame, unsigned numFrames, const SIFTKeyPoint* d_keys, const int* d_num, const float* d_dist,
                                     const uint2* d_idx, int* d_numF, float* d_distF, uint2* d_idxF, float4x4* d_T, float4x4* d_Tinv, const float* siftIntrinsicsInv,
                                     unsigned minNumMatches, float maxKabschRes2) {
    if (numFrames <= startFrame) return 0;
    FilterKeyPointMatchesCU_Kernel<<<dim3(numFrames - startFrame), dim3(FILTER_NUM_BLOCK_THREADS_X)>>>(curFrame, startFrame, d_keys, d_num, d_dist, d_idx, d_numF, d_distF,
                                                                                                     d_idxF, d_T, d_Tinv, to_m(siftIntrinsicsInv), minNumMatches, maxKabschRes2);
    return (int)cudaDeviceSynchronize();
}

// SIFTImageManager::FilterMatchesBySurfaceAreaCU, SIFTImageManager.cu:391-407
REF_API int refFilterMatchesBySurfaceArea(unsigned curFrame, unsigned startFrame, unsigned numFrames, const SIFTKeyPoint* d_keys, int* d_numF, const uint2* d_idxF,
                                          const float* colorIntrinsicsInv, float areaThresh) {
    if (numFrames <= startFrame) return 0;
    const unsigned threads = ((MAX_MATCHES_PER_IMAGE_PAIR_FILTERED + 31) / 32) * 32;
    FilterMatchesBySurfaceAreaCU_Kernel<<<dim3(numFrames - startFrame), dim3(threads)>>>(curFrame, startFrame, d_keys, d_numF, d_idxF, to_m(colorIntrinsicsInv), areaThresh);
    return (int)cudaDeviceSynchronize();
}

// SIFTImageManager::FilterMatchesByDenseVerifyCU, SIFTImageManager.cu:587-608
REF_API int refFilterMatchesByDenseVerify(unsigned curFrame, unsigned startFrame, unsigned numFrames, unsigned w, unsigned h, const float* intrinsics, int* d_numF,
                                          const float4x4* d_T, const float4x4* d_Tinv, const CUDACachedFrame* d_frames, float distThresh, float normalThresh,
                                          float colorThresh, float errThresh, float corrThresh, float dMin, float dMax) {
    if (numFrames <= startFrame) return 0;
    dim3 block(w, (h + FILTER_DENSE_VERIFY_THREAD_SPLIT - 1) / FILTER_DENSE_VERIFY_THREAD_SPLIT);
    FilterMatchesByDenseVerifyCU_Kernel<<<dim3(numFrames - startFrame), block>>>(curFrame, startFrame, w, h, to_m(intrinsics), d_numF, d_T, d_Tinv, d_frames, distThresh,
                                                                               normalThresh, colorThresh, errThresh, corrThresh, dMin, dMax);
    return (int)cudaDeviceSynchronize();
}

// SIFTImageManager::AddCurrToResidualsCU, SIFTImageManager.cu:657-685
REF_API int refAddCurrToResiduals(unsigned curFrame, unsigned startFrame, unsigned numFrames, EntryJ* d_glob, uint2* d_globIdx, int* d_globNum, const int* d_numF,
                                  const uint2* d_idxF, const SIFTKeyPoint* d_keys, unsigned maxKeyPointsPerImage, const float* colorIntrinsicsInv) {
    if (numFrames <= startFrame) return 0;
    const unsigned threads = ((MAX_MATCHES_PER_IMAGE_PAIR_FILTERED + 31) / 32) * 32;
    AddCurrToResidualsCU_Kernel<<<dim3(numFrames - startFrame), dim3(threads)>>>(curFrame, startFrame, d_glob, d_globIdx, d_globNum, d_numF, d_idxF, d_keys,
                                                                                maxKeyPointsPerImage, to_m(colorIntrinsicsInv));
    return (int)cudaDeviceSynchronize();
}
