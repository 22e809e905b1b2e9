/*
 * cache_oracle.c -- CPU restatement of CUDACache::storeFrame, kernel by kernel, with the
 * reference's full-resolution intermediates (the CUDA library evaluates the same values at cache resolution only).
 *
 * TEST INFRASTRUCTURE ONLY (see tsdf_oracle.c header).  PARITY STATUS: PINNED against the reference's own kernels executed on the CPU -- FL/CUDAImageUtil.cu, called in the order of CUDACache::storeFrame, compiled by g++ against the CUDA emulation (oracle/build_ref.py build_mgr_emulated -> oracle/_ref/libref_mgr_emulated.so), outputs committed as tests/golden/manager_reference_emulated.npz, tests/test_manager_reference_emulated.py: depth, camera positions, float and uchar4 normals bit for bit; intensity and its derivatives within 3e-7 (this file places the fused multiply-adds nvcc emits for the reference's expressions, the emulation is built without contraction).
 * Also pinned by the known-answer tests in
 * tests/test_cache_oracle.py (analytic plane, hand-computed filter cases, agreement with the independent numpy generator
 * bundlefusion_b200/synth.py: make_cache_frame).
 *
 * Restates (FL/ = FriedLiver/Source/):
 *   gaussFilterDepthMapDevice                FL/CUDAImageUtil.cu:759-794   (range-gated Gaussian, window ceil(2 sigmaD))
 *   convertDepthFloatToCameraSpaceFloat4     FL/CUDAImageUtil.cu:367-386
 *   computeNormals_Kernel                    FL/CUDAImageUtil.cu:404-431
 *   resampleFloat / resampleFloat4           FL/CUDAImageUtil.cu:93-150    (nearest, scale (in-1)/(out-1))
 *   convertNormalsFloat4ToUCHAR4_Kernel      FL/CUDAImageUtil.cu:497-513
 *   resampleToIntensity_Kernel               FL/CUDAImageUtil.cu:224-243, convertToIntensity :196-199
 *   gaussFilterIntensityDevice               FL/CUDAImageUtil.cu:811-848
 *   computeIntensityDerivatives_Kernel       FL/CUDAImageUtil.cu:260-296
 *   CUDACache::storeFrame                    FL/CUDACache.cpp:45-86
 * Arithmetic contract shared with bundlefusion_b200/csrc/cache.cu (bit-exact comparison): binary32, every operation individually
 * rounded except the two places written as fmaf below (x * scale + 0.5, the intensity dot product), Gaussian weights
 * exp(-(dx^2 + dy^2) / (2 sigma^2)) evaluated ONCE per offset by the host's expf (the library does the same on the host), sums in
 * the reference's loop order (x outer, y inner).  Against the reference's own GPU expf / FMA contraction that is a last-bit matter.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/bf_cache.h"

#define ORC_API __attribute__((visibility("default")))
#define MINF (-INFINITY)

static inline float gaussD(float sigma, int x, int y) { return expf(-((float)(x * x + y * y) / (2.0f * sigma * sigma))); }
static inline unsigned src_index(unsigned o, float scale) { return (unsigned)fmaf((float)o, scale, 0.5f); }

/* host arrays in, host arrays out (each [h][w] of the cache): depth f32, campos f32x4, normals f32x4, normalsU u8x4, intensity f32, derivs f32x2 */
ORC_API void orc_cache_store_frame(const BFCacheParams* P, const float* depth, const uint8_t* color,
                                   float* oDepth, float* oCampos, float* oNormals, uint8_t* oNormalsU, float* oIntensity, float* oDerivs) {
    const int W = (int)P->inputDepthWidth, H = (int)P->inputDepthHeight, w = (int)P->width, h = (int)P->height;
    const float* M = P->inputIntrinsicsInv;
    /* depth (optionally range-gated Gaussian) */
    float* dF = (float*)malloc(sizeof(float) * (size_t)W * H);
    if (P->filterDepthSigmaD > 0.0f) {
        const float sD = P->filterDepthSigmaD, sR = P->filterDepthSigmaR;
        const int r = (int)ceil(2.0 * sD);
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            float sum = 0.0f, sumW = 0.0f, out = MINF;
            const float c = depth[y * W + x];
            if (c != MINF)
                for (int m = x - r; m <= x + r; ++m) for (int n = y - r; n <= y + r; ++n)
                    if (m >= 0 && n >= 0 && m < W && n < H) {
                        const float cur = depth[n * W + m];
                        if (cur != MINF && fabsf(c - cur) < sR) { const float wgt = gaussD(sD, m - x, n - y); sumW += wgt; sum += wgt * cur; }
                    }
            if (sumW > 0.0f) out = sum / sumW;
            dF[y * W + x] = out;
        }
    } else memcpy(dF, depth, sizeof(float) * (size_t)W * H);
    /* camera-space positions and normals at full resolution */
    float* cam = (float*)malloc(sizeof(float) * 4 * (size_t)W * H);
    float* nrm = (float*)malloc(sizeof(float) * 4 * (size_t)W * H);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        float* o = &cam[4 * (y * W + x)];
        o[0] = o[1] = o[2] = o[3] = MINF;
        const float d = dF[y * W + x];
        if (d != MINF) {
            const float v[4] = { (float)x * d, (float)y * d, d, d };
            float r4[4];
            for (int k = 0; k < 4; ++k) r4[k] = ((M[4 * k] * v[0] + M[4 * k + 1] * v[1]) + M[4 * k + 2] * v[2]) + M[4 * k + 3] * v[3];
            o[0] = r4[0]; o[1] = r4[1]; o[2] = r4[3]; o[3] = 1.0f;
        }
    }
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        float* o = &nrm[4 * (y * W + x)];
        o[0] = o[1] = o[2] = o[3] = MINF;
        if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
            const float *CC = &cam[4 * (y * W + x)], *PC = &cam[4 * ((y + 1) * W + x)], *CP = &cam[4 * (y * W + x + 1)],
                        *MC = &cam[4 * ((y - 1) * W + x)], *CM = &cam[4 * (y * W + x - 1)];
            if (CC[0] != MINF && PC[0] != MINF && CP[0] != MINF && MC[0] != MINF && CM[0] != MINF) {
                const float a[3] = { PC[0] - MC[0], PC[1] - MC[1], PC[2] - MC[2] }, b[3] = { CP[0] - CM[0], CP[1] - CM[1], CP[2] - CM[2] };
                const float n[3] = { a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0] };
                const float l = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
                if (l > 0.0f) { o[0] = n[0] / -l; o[1] = n[1] / -l; o[2] = n[2] / -l; o[3] = 0.0f; }
            }
        }
    }
    /* nearest-neighbour resample to the cache, normals -> uchar4 */
    const float sw = (float)(W - 1) / (float)(w - 1), sh = (float)(H - 1) / (float)(h - 1);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
        const unsigned xi = src_index((unsigned)x, sw), yi = src_index((unsigned)y, sh);
        const int o = y * w + x;
        if (xi < (unsigned)W && yi < (unsigned)H) {
            oDepth[o] = dF[yi * W + xi];
            memcpy(&oCampos[4 * o], &cam[4 * (yi * W + xi)], 16);
            memcpy(&oNormals[4 * o], &nrm[4 * (yi * W + xi)], 16);
        }
        uint8_t* u = &oNormalsU[4 * o];
        u[0] = u[1] = u[2] = u[3] = 0;
        if (oNormals[4 * o] != MINF)
            for (int k = 0; k < 3; ++k) { const float p = (oNormals[4 * o + k] + 1.0f) / 2.0f; u[k] = (uint8_t)roundf(p * 255.0f); }
    }
    /* intensity: resample, Gaussian, Sobel / 8 */
    const int CW = (int)P->inputColorWidth, CH = (int)P->inputColorHeight;
    const float cw = (float)(CW - 1) / (float)(w - 1), chh = (float)(CH - 1) / (float)(h - 1);
    float* ih = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
        const unsigned xi = src_index((unsigned)x, cw), yi = src_index((unsigned)y, chh);
        ih[y * w + x] = 0.0f;
        if (xi < (unsigned)CW && yi < (unsigned)CH) {
            const uint8_t* c = &color[4 * (yi * CW + xi)];
            ih[y * w + x] = fmaf(0.114f, (float)c[2], fmaf(0.299f, (float)c[0], 0.587f * (float)c[1])) / 255.0f;
        }
    }
    if (P->filterIntensitySigma > 0.0f) {
        const float s = P->filterIntensitySigma;
        const int r = (int)ceil(2.0 * s);
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
            float sum = 0.0f, sumW = 0.0f;
            for (int m = x - r; m <= x + r; ++m) for (int n = y - r; n <= y + r; ++n)
                if (m >= 0 && n >= 0 && m < w && n < h) { const float wgt = gaussD(s, m - x, n - y); sumW += wgt; sum += wgt * ih[n * w + m]; }
            oIntensity[y * w + x] = (sumW > 0.0f) ? sum / sumW : 0.0f;
        }
    } else memcpy(oIntensity, ih, sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
        float* o = &oDerivs[2 * (y * w + x)];
        o[0] = o[1] = MINF;
        if (x > 0 && x < w - 1 && y > 0 && y < h - 1) {
            const float* I = oIntensity;
            const float p00 = I[(y - 1) * w + x - 1], p01 = I[y * w + x - 1], p02 = I[(y + 1) * w + x - 1], p10 = I[(y - 1) * w + x],
                        p12 = I[(y + 1) * w + x], p20 = I[(y - 1) * w + x + 1], p21 = I[y * w + x + 1], p22 = I[(y + 1) * w + x + 1];
            if (p00 == MINF || p01 == MINF || p02 == MINF || p10 == MINF || p12 == MINF || p20 == MINF || p21 == MINF || p22 == MINF) continue;
            float rU = (-1.0f) * p00 + (1.0f) * p20 + (-2.0f) * p01 + (2.0f) * p21 + (-1.0f) * p02 + (1.0f) * p22;
            float rV = (-1.0f) * p00 + (-2.0f) * p10 + (-1.0f) * p20 + (1.0f) * p02 + (2.0f) * p12 + (1.0f) * p22;
            o[0] = rU / 8.0f; o[1] = rV / 8.0f;
        }
    }
    free(dF); free(cam); free(nrm); free(ih);
}
