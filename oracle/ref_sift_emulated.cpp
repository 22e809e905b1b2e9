// ref_sift_emulated.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own SiftGPU (FL/SiftGPU: SiftGPU, SiftPyramid,
// CuTexImage, GlobalUtil, ProgramCU.cu kernels, SiftMatchGPU), compiled by g++ against the CPU emulation of oracle/ref_emu/ref_emu_cuda.h
// (oracle/build_ref.py -> oracle/_ref/libref_sift_emulated.so).  The calls below are the ones Bundler makes (FL/Bundler.cpp:55-100, 117-136).
// This file contains no reference code.
#include "SiftGPU/SiftGPU.h"
#include "SiftGPU/SiftMatch.h"
#include "SiftGPU/SiftCameraParams.h"
#include "SiftGPU/GlobalUtil.h"

extern "C" void updateConstantSiftCameraParams(const SiftCameraParams& params);

#define REF_API extern "C" __attribute__((visibility("default")))

struct RefSiftParams { unsigned width, height, depthWidth, depthHeight; float depthMin, depthMax, minKeyScale; int featureCountThreshold; unsigned maxKeyPoints; };

// SiftGPU::SetParams + InitSiftGPU + RunSIFT + GetKeyPointsAndDescriptorsCUDA (FL/Bundler.cpp:55-100).  All pointers are "device" = host here.
REF_API int refEmuSiftDetect(const float* intensity, const float* depth, const RefSiftParams* P, float* keyPoints, unsigned char* descriptors) {
    SiftCameraParams cp;
    std::memset(&cp, 0, sizeof cp);
    cp.m_depthWidth = P->depthWidth; cp.m_depthHeight = P->depthHeight; cp.m_intensityWidth = P->width; cp.m_intensityHeight = P->height;
    cp.m_minKeyScale = P->minKeyScale;
    updateConstantSiftCameraParams(cp);
    if (getenv("EMU_TRACE")) fprintf(stderr, "[ref] constants set\n");
    SiftGPU* sift = new SiftGPU;
    sift->SetParams(P->width, P->height, false, (unsigned)P->featureCountThreshold, P->depthMin, P->depthMax);
    if (getenv("EMU_TRACE")) fprintf(stderr, "[ref] params set\n");
    sift->InitSiftGPU();
    if (getenv("EMU_TRACE")) fprintf(stderr, "[ref] initialised\n");
    const int ok = sift->RunSIFT(const_cast<float*>(intensity), depth);
    int n = -1;
    if (ok) {
        SIFTImageGPU img;
        img.d_keyPoints = reinterpret_cast<SIFTKeyPoint*>(keyPoints);
        img.d_keyPointDescs = reinterpret_cast<SIFTKeyPointDesc*>(descriptors);
        n = (int)sift->GetKeyPointsAndDescriptorsCUDA(img, depth, P->maxKeyPoints);
    }
    delete sift;
    return n;
}

// SiftMatchGPU::SetDescriptors x2 + GetSiftMatch (FL/Bundler.cpp:117-136): returns the match counter; idx [128][2], dist [128]
REF_API int refEmuSiftMatch(unsigned char* des1, int n1, unsigned char* des2, int n2, float distmax, float ratiomax, unsigned* idx, float* dist, unsigned offX, unsigned offY) {
    static SiftMatchGPU* m = nullptr;
    if (!m) { m = new SiftMatchGPU(4096); m->InitSiftMatch(); }
    int num = 0;
    ImagePairMatch ipm; ipm.d_numMatches = &num; ipm.d_distances = dist; ipm.d_keyPointIndices = reinterpret_cast<uint2*>(idx);
    m->SetDescriptors(0, n1, des1);
    m->SetDescriptors(1, n2, des2);
    m->GetSiftMatch(n1, ipm, make_uint2(offX, offY), distmax, ratiomax);
    return num;
}
