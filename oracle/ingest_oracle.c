/*
 * ingest_oracle.c -- CPU restatement of the device part of CUDAImageManager::process, pass by pass
 * with the reference's full-image intermediates.
 *
 * TEST INFRASTRUCTURE ONLY (see tsdf_oracle.c header).  PARITY STATUS: PINNED against the reference's own kernels executed on the CPU -- FL/CUDAImageUtil.cu, called in the order of CUDAImageManager::process, compiled by g++ against the CUDA emulation (oracle/build_ref.py build_mgr_emulated -> oracle/_ref/libref_mgr_emulated.so), outputs committed as tests/golden/manager_reference_emulated.npz, tests/test_manager_reference_emulated.py: depth and colour at the integration resolution bit for bit (two image sizes x four configurations: eroded / filtered / copied / resampled).
 * Also pinned by the known-answer tests in tests/test_ingest_oracle.py.
 *
 * Restates: erodeDepthMapDevice FL/CUDAImageUtil.cu:701-741, gaussFilterDepthMapDevice :759-794, resampleFloat_Kernel :93-110,
 * resampleUCHAR4_Kernel :160-177, sequencing FL/CUDAImageManager.cpp:44-61, 88-137.
 * Arithmetic contract shared with bundlefusion_b200/csrc/ingest.cu (bit-exact): as oracle/cache_oracle.c -- host expf weights once
 * per offset, sums in the reference's loop order (x outer, y inner), fmaf only in x * scale + 0.5.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/bf_ingest.h"

#define ORC_API __attribute__((visibility("default")))
#define MINF (-INFINITY)

static void erode(float* out, const float* in, int s, int W, int H, float dThresh, float fracReq) {
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        unsigned count = 0;
        const float old = in[y * W + x];
        for (int i = -s; i <= s; ++i) for (int j = -s; j <= s; ++j)
            if (x + j >= 0 && x + j < W && y + i >= 0 && y + i < H) {
                const float d = in[(y + i) * W + x + j];
                if (d == MINF || d == 0.0f || fabsf(d - old) > dThresh) ++count;
            }
        const unsigned sum = (unsigned)((2 * s + 1) * (2 * s + 1));
        out[y * W + x] = ((float)count / (float)sum >= fracReq) ? MINF : old;
    }
}
static inline unsigned src_index(unsigned o, float scale) { return (unsigned)fmaf((float)o, scale, 0.5f); }

ORC_API void orc_ingest_frame(const BFIngestParams* P, const float* depthRaw, const uint8_t* colorRaw, float* depthOut, uint8_t* colorOut) {
    const int W = (int)P->depthWidth, H = (int)P->depthHeight, w = (int)P->widthIntegration, h = (int)P->heightIntegration;
    float* a = (float*)malloc(sizeof(float) * (size_t)W * H);
    float* b = (float*)malloc(sizeof(float) * (size_t)W * H);
    memcpy(a, depthRaw, sizeof(float) * (size_t)W * H);
    const int iters = 2 * ((P->erodeIterations + 1) / 2);
    for (int i = 0; i < iters; ++i) { erode(b, a, P->erodeStructureSize, W, H, P->erodeDThresh, P->erodeFracReq); float* t = a; a = b; b = t; }
    if (P->depthSigmaD > 0.0f) {
        const float sD = P->depthSigmaD, sR = P->depthSigmaR;
        const int r = (int)ceil(2.0 * sD);
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            float sum = 0.0f, sumW = 0.0f, out = MINF;
            const float c = a[y * W + x];
            if (c != MINF)
                for (int m = x - r; m <= x + r; ++m) for (int n = y - r; n <= y + r; ++n)
                    if (m >= 0 && n >= 0 && m < W && n < H) {
                        const float cur = a[n * W + m];
                        if (cur != MINF && fabsf(c - cur) < sR) { const float wgt = expf(-((float)((m - x) * (m - x) + (n - y) * (n - y)) / (2.0f * sD * sD))); sumW += wgt; sum += wgt * cur; }
                    }
            if (sumW > 0.0f) out = sum / sumW;
            b[y * W + x] = out;
        }
        float* t = a; a = b; b = t;
    }
    if (W == w && H == h) memcpy(depthOut, a, sizeof(float) * (size_t)W * H);
    else {
        const float sw = (float)(W - 1) / (float)(w - 1), sh = (float)(H - 1) / (float)(h - 1);
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
            const unsigned xi = src_index((unsigned)x, sw), yi = src_index((unsigned)y, sh);
            if (xi < (unsigned)W && yi < (unsigned)H) depthOut[y * w + x] = a[yi * W + xi];
        }
    }
    const int CW = (int)P->colorWidth, CH = (int)P->colorHeight;
    if (CW == w && CH == h) memcpy(colorOut, colorRaw, 4 * (size_t)CW * CH);
    else {
        const float sw = (float)(CW - 1) / (float)(w - 1), sh = (float)(CH - 1) / (float)(h - 1);
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
            const unsigned xi = src_index((unsigned)x, sw), yi = src_index((unsigned)y, sh);
            if (xi < (unsigned)CW && yi < (unsigned)CH) memcpy(&colorOut[4 * (y * w + x)], &colorRaw[4 * (yi * CW + xi)], 4);
        }
    }
    free(a); free(b);
}
