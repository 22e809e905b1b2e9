// ref_fuse_wrap.cpp -- TEST INFRASTRUCTURE ONLY (oracle/build_ref.py: build_fuse_emulated).  Entry point around the reference's OWN chunk -> keyframe fusion:
// SIFTImageManager::fuseToGlobal / computeTracks / findTrack (FL/SiftGPU/SIFTImageManager.cpp:366-476), host code of the reference's manager class, compiled with the
// class (SIFTImageManager.cpp + SIFTImageManager.cu) from the scratch copy against the CPU emulation of CUDA, so that oracle/fuse_oracle.c can be pinned.
// The manager is filled through its own public interface (createSIFTImageGPU / finalizeSIFTImageGPU, the global correspondence arrays it exposes).
#include "stdafx.h"
#ifndef SAFE_DELETE
#define SAFE_DELETE(p) { if (p) { delete (p); (p) = NULL; } }          // mLib core-base/common.h
#endif
#define private public                                       // the key-index array of the global correspondences has no setter (SIFTImageManager.h:340)
#include "SiftGPU/SIFTImageManager.h"
#undef private
#include "SiftGPU/SIFTImageManager.cu"
#include "SiftGPU/SIFTImageManager.cpp"

// keys [numImages][keysPerImage (rows used: numKeys[i])][4], descs likewise [128 bytes]; corr: EntryJ [numCorr], corrKeys uint2 [numCorr] (global key indices);
// transforms float4x4 [numImages]; out: the fused image's keys / descriptors; returns its key count
extern "C" int ref_fuse_to_global(unsigned numImages, const unsigned* numKeys, unsigned keysPerImage, const float* keys, const unsigned char* descs, unsigned numCorr, const void* corr,
                                  const unsigned* corrKeys, const float* transforms, const float* colorIntrinsics, unsigned maxKeysGlobal, float* outKeys, unsigned char* outDescs) {
    SIFTImageManager local(numImages + 1, keysPerImage), global(4, maxKeysGlobal);
    for (unsigned i = 0; i < numImages; ++i) {
        SIFTImageGPU& img = local.createSIFTImageGPU();
        memcpy(img.d_keyPoints, keys + (size_t)4 * keysPerImage * i, sizeof(SIFTKeyPoint) * numKeys[i]);
        memcpy(img.d_keyPointDescs, descs + (size_t)128 * keysPerImage * i, sizeof(SIFTKeyPointDesc) * numKeys[i]);
        local.finalizeSIFTImageGPU(numKeys[i]);
    }
    memcpy(local.d_globMatches, corr, sizeof(EntryJ) * numCorr);
    memcpy(local.d_globMatchesKeyPointIndices, corrKeys, sizeof(uint2) * numCorr);
    local.m_globNumResiduals = numCorr;
    float4x4 K, Kinv;
    for (int k = 0; k < 16; ++k) K.entries[k] = colorIntrinsics[k];
    Kinv = K.getInverse();
    local.fuseToGlobal(&global, K, (const float4x4*)transforms, Kinv);
    const unsigned n = global.getNumKeyPointsPerImage(0);
    memcpy(outKeys, global.getImageGPU(0).d_keyPoints, sizeof(SIFTKeyPoint) * n);
    memcpy(outDescs, global.getImageGPU(0).d_keyPointDescs, sizeof(SIFTKeyPointDesc) * n);
    return (int)n;
}

// SIFTImageManager::filterFrames (SIFTImageManager.cpp:551-575): numFiltered [numFrames] (the current frame's filtered match counts per earlier frame), validImages
// [>= max(numFrames, curFrame + 1)] in / out; returns the last matched frame or (unsigned)-1
extern "C" unsigned ref_filter_frames(unsigned curFrame, unsigned startFrame, unsigned numFrames, const int* numFiltered, int* validImages, unsigned numValid) {
    SIFTImageManager m(numValid + 2, 16);
    std::vector<int> v(validImages, validImages + numValid);
    v.resize(numValid + 2, 0);
    m.setValidImagesDEBUG(v);
    memcpy(m.d_currNumFilteredMatchesPerImagePair, numFiltered, sizeof(int) * numFrames);
    const unsigned last = m.filterFrames(curFrame, startFrame, numFrames);
    for (unsigned i = 0; i < numValid; ++i) validImages[i] = m.m_validImages[i];
    return last;
}
