// ref_imageutil_wrap.cu -- TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own image kernels (FL/CUDAImageUtil.cu) and
// trajectory kernels (FL/OnlineBundler.cu) for the compat build oracle/build_ref.py makes in /tmp (-> oracle/_ref/libref_imageutil.so).
// It contains no reference code: it includes the scratch copies as one translation unit and calls the reference's CUDAImageUtil statics
// in the ORDER its two callers use them -- CUDACache::storeFrame (FL/CUDACache.cpp:45-86) and CUDAImageManager::process
// (FL/CUDAImageManager.cpp:88-137, 44-61) -- on raw device pointers, so that rows a20 / a21 can be compared with the reference on a GPU.
// The three extern "C" stubs of OnlineBundler.cu (row a22) are exported by the included file itself.
#include "CUDAImageUtil.cu"
#include "OnlineBundler.cu"
#include "CUDACacheUtil.h"

#define REF_API extern "C" __attribute__((visibility("default")))

// CUDACache::storeFrame, FL/CUDACache.cpp:45-86 (both normal formats, as FL/CUDACacheUtil.h:7-8 defines both macros).  Helper buffers are the caller's:
// d_filterHelper [H][W] float, d_helperCamPos / d_helperNormals [H][W] float4, d_intensityHelper [h][w] float.
REF_API int refCacheStoreFrame(const float* d_depth, unsigned W, unsigned H, const uchar4* d_color, unsigned CW, unsigned CH, unsigned w, unsigned h,
                               const float* inputIntrinsicsInv, float filterIntensitySigma, float filterDepthSigmaD, float filterDepthSigmaR,
                               CUDACachedFrame frame, float* d_filterHelper, float4* d_helperCamPos, float4* d_helperNormals, float* d_intensityHelper) {
    float4x4 Ki; for (int i = 0; i < 16; ++i) Ki.entries[i] = inputIntrinsicsInv[i];
    const float* d_inputDepth = d_depth;
    if (filterDepthSigmaD > 0.0f) { CUDAImageUtil::gaussFilterDepthMap(d_filterHelper, d_depth, filterDepthSigmaD, filterDepthSigmaR, W, H); d_inputDepth = d_filterHelper; }
    CUDAImageUtil::convertDepthFloatToCameraSpaceFloat4(d_helperCamPos, d_inputDepth, Ki, W, H);
    CUDAImageUtil::resampleFloat4(frame.d_cameraposDownsampled, w, h, d_helperCamPos, W, H);
    CUDAImageUtil::computeNormals(d_helperNormals, d_helperCamPos, W, H);
    CUDAImageUtil::resampleFloat4(frame.d_normalsDownsampled, w, h, d_helperNormals, W, H);
    CUDAImageUtil::convertNormalsFloat4ToUCHAR4(frame.d_normalsDownsampledUCHAR4, frame.d_normalsDownsampled, w, h);
    CUDAImageUtil::resampleFloat(frame.d_depthDownsampled, w, h, d_inputDepth, W, H);
    CUDAImageUtil::resampleToIntensity(d_intensityHelper, w, h, d_color, CW, CH);
    if (filterIntensitySigma > 0.0f) CUDAImageUtil::gaussFilterIntensity(frame.d_intensityDownsampled, d_intensityHelper, filterIntensitySigma, w, h);
    else cudaMemcpy(frame.d_intensityDownsampled, d_intensityHelper, sizeof(float) * w * h, cudaMemcpyDeviceToDevice);      // the reference swaps the two pointers
    CUDAImageUtil::computeIntensityDerivatives(frame.d_intensityDerivsDownsampled, frame.d_intensityDownsampled, w, h);
    return (int)cudaDeviceSynchronize();
}

// CUDAImageManager::process, depth part (FL/CUDAImageManager.cpp:88-137): erosions (ping-pong between the raw and the filtered buffer, raw is
// overwritten as in the reference), range-gated Gaussian or copy, copy / resample to the integration resolution; colour part (:44-61)
REF_API int refIngestFrame(float* d_depthRaw, float* d_depthFiltered, unsigned W, unsigned H, const uchar4* d_colorRaw, unsigned CW, unsigned CH, unsigned w, unsigned h,
                           int erode, int structureSize, float erodeDThresh, float erodeFracReq, float depthSigmaD, float depthSigmaR,
                           float* d_depthIntegration, uchar4* d_colorIntegration) {
    if (erode) {
        unsigned numIter = 2; numIter = 2 * ((numIter + 1) / 2);
        for (unsigned i = 0; i < numIter; ++i) {
            if (i % 2 == 0) CUDAImageUtil::erodeDepthMap(d_depthFiltered, d_depthRaw, structureSize, W, H, erodeDThresh, erodeFracReq);
            else CUDAImageUtil::erodeDepthMap(d_depthRaw, d_depthFiltered, structureSize, W, H, erodeDThresh, erodeFracReq);
        }
    }
    if (depthSigmaD > 0.0f) CUDAImageUtil::gaussFilterDepthMap(d_depthFiltered, d_depthRaw, depthSigmaD, depthSigmaR, W, H);
    else CUDAImageUtil::copy<float>(d_depthFiltered, d_depthRaw, W, H);
    if (W == w && H == h) CUDAImageUtil::copy<float>(d_depthIntegration, d_depthFiltered, w, h);
    else CUDAImageUtil::resampleFloat(d_depthIntegration, w, h, d_depthFiltered, W, H);
    if (CW == w && CH == h) CUDAImageUtil::copy<uchar4>(d_colorIntegration, const_cast<uchar4*>(d_colorRaw), w, h);
    else CUDAImageUtil::resampleUCHAR4(d_colorIntegration, w, h, d_colorRaw, CW, CH);
    return (int)cudaDeviceSynchronize();
}
