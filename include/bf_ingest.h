/*
 * bf_ingest.h -- C-ABI of the per-frame sensor-image preparation.
 *
 * Replaces the device part of CUDAImageManager::process (FL/CUDAImageManager.cpp:22-158) for frames already on the device:
 *   depth : erodeDepthMap x2 (7x7 window, dThresh 0.05, fracReq 0.3; FL/CUDAImageUtil.cu:701-741, CUDAImageManager.cpp:88-101)
 *           -> gaussFilterDepthMap(s_depthSigmaD, s_depthSigmaR) (:759-794, CUDAImageManager.cpp:102-109)
 *           -> copy / resampleFloat to the integration resolution (:93-110, CUDAImageManager.cpp:119-137)
 *   colour: copy / resampleUCHAR4 to the integration resolution (:160-177, CUDAImageManager.cpp:44-61)
 * The reference runs this as 4-5 full-image launches with three intermediate images; here it is ONE launch: a CTA stages a raw
 * depth tile with a 10-pixel halo in shared memory, erodes twice and filters in place, and writes the integration-resolution
 * pixels that fall into its tile; the same CTA resamples the colour pixels of its tile.  (FL/ = /root/reference/FriedLiver/Source/.)
 * Out of scope: the host -> device copies of the sensor buffers (the caller's) and the D3D11 colour-space re-projection
 * (s_bUseCameraCalibration, a DirectX path).
 */
#ifndef BF_INGEST_H
#define BF_INGEST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct BFIngestParams {
    uint32_t depthWidth, depthHeight;        /* sensor depth image */
    uint32_t colorWidth, colorHeight;        /* sensor colour image */
    uint32_t widthIntegration, heightIntegration;
    int32_t  erodeIterations;                /* s_erodeSIFTdepth ? 2 : 0 (the reference rounds to an even count, CUDAImageManager.cpp:89-90) */
    int32_t  erodeStructureSize;             /* 3 */
    float    erodeDThresh, erodeFracReq;     /* 0.05, 0.3 */
    float    depthSigmaD, depthSigmaR;       /* s_depthFilter ? (2.0, 0.05) : sigmaD <= 0 (copy) */
} BFIngestParams;

/* d_depthRaw float [depthHeight][depthWidth] (-inf invalid), d_colorRaw uchar4 [colorHeight][colorWidth];
 * outputs at the integration resolution.  Inputs are not modified.  Asynchronous on the library stream. */
int bfIngestFrame(const BFIngestParams* params, const float* d_depthRaw, const uint8_t* d_colorRaw, float* d_depthIntegration, uint8_t* d_colorIntegration);

#ifdef __cplusplus
}
#endif
#endif /* BF_INGEST_H */
