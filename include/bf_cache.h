/*
 * bf_cache.h -- C-ABI of the dense-cache frame builder.
 *
 * Replaces the body of  CUDACache::storeFrame(const float* d_depth, w, h, const uchar4* d_color, cw, ch)
 * (FL/CUDACache.cpp:45-86), which the reference runs as eight launches over FULL-resolution intermediates
 * (FL/CUDAImageUtil.cu: gaussFilterDepthMap :759, convertDepthFloatToCameraSpaceFloat4 :367, computeNormals :404,
 * resampleFloat4 :126 x2, convertNormalsFloat4ToUCHAR4 :497, resampleFloat :93, resampleToIntensity :224,
 * gaussFilterIntensity :811, computeIntensityDerivatives :260) to fill one 80x60 CUDACachedFrame
 * (FL/CUDACacheUtil.h:41-53, = BFCUDACachedFrame in bf_solver.h).  (FL/ = /root/reference/FriedLiver/Source/.)
 * Here it is ONE launch that evaluates everything at cache resolution (the filters are evaluated only where a cache pixel
 * needs them); outputs are the reference's formats.
 */
#ifndef BF_CACHE_H
#define BF_CACHE_H

#include <stddef.h>
#include <stdint.h>

#include "bf_solver.h"      /* BFCUDACachedFrame */

#ifdef __cplusplus
extern "C" {
#endif

/* what CUDACache's constructor latches (FL/CUDACache.cpp:14-40) */
typedef struct BFCacheParams {
    uint32_t inputDepthWidth, inputDepthHeight;     /* m_inputDepthWidth / Height */
    uint32_t inputColorWidth, inputColorHeight;
    uint32_t width, height;                         /* m_width, m_height: the cache resolution (80 x 60) */
    float    inputIntrinsicsInv[16];                /* m_inputIntrinsicsInv, row-major 4x4 */
    float    filterIntensitySigma;                  /* s_colorDownSigma  (2.5); <= 0: no intensity filter */
    float    filterDepthSigmaD;                     /* s_depthDownSigmaD (1.0); <= 0: no depth filter */
    float    filterDepthSigmaR;                     /* s_depthDownSigmaR (0.05) */
} BFCacheParams;

/* Fills *frame (a HOST struct of six device pointers) from a full-resolution depth (float, -inf = invalid) and colour (uchar4)
 * image.  Asynchronous on the library stream.  Returns 0 or a cudaError_t. */
int bfCacheStoreFrame(const BFCacheParams* params, const float* d_depth, const uint8_t* d_color, const BFCUDACachedFrame* frame);

#ifdef __cplusplus
}
#endif
#endif /* BF_CACHE_H */
