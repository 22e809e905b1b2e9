// bf_reference_classes.hpp -- C++ shim: the reference's SIFT class surface over this library's C-ABI, so that the call sites of
// FL/Bundler.cpp compile against it unchanged (FL/ = /root/reference/FriedLiver/Source/).
//
//   class SiftGPU            FL/SiftGPU/SiftGPU.h:73-114      SetParams / InitSiftGPU / RunSIFT / GetKeyPointsAndDescriptorsCUDA / GetFeatureNum
//   class SiftMatchGPU       FL/SiftGPU/SiftMatch.h:15-47      InitSiftMatch / SetDescriptors / GetSiftMatch
//   class SIFTImageManager   FL/SiftGPU/SIFTImageManager.h:62-330   storage, bookkeeping and the *CU members (Sort / Filter... / AddCurrToResiduals /
//                                                                Invalidate / CheckForInvalidFrames / VerifyTrajectory), filterFrames, fuseToGlobal
//   structs SIFTKeyPoint, SIFTKeyPointDesc, SIFTImageGPU, ImagePairMatch, EntryJ, float4x4 (row-major 16 floats)
//
// Header-only; needs <cuda_runtime.h> (the reference's classes own their device buffers through cudaMalloc) and links against
// libbundlefusion_b200.so.  Semantics are the reference's (same method names, argument meaning, key points packed by a prefix sum, counts copied
// to the host where the reference copies them); what differs is inside the calls: one launch pair per GetSiftMatch instead of three plus a
// memset, detection without a mid-pipeline host round trip, fuseToGlobal on the device.  The sync-free path that never leaves the device is the
// frame loop (bf_frameloop.h); this shim is the drop-in for code that keeps the reference's host sequencing.
// tests/test_reference_classes_shim.py compiles and links a translation unit that uses every method (g++, no GPU needed) and, on a GPU, runs
// Bundler::detectFeatures + matchAndFilter written against these classes.
#pragma once
#include <cuda_runtime.h>

#include <cassert>
#include <cstring>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

#include "bf_sift.h"
#include "bf_solver.h"

#ifndef MAX_MATCHES_PER_IMAGE_PAIR_RAW
#define MAX_MATCHES_PER_IMAGE_PAIR_RAW BF_MAX_MATCHES_PER_IMAGE_PAIR_RAW            // FL/GlobalDefines.h:8
#define MAX_MATCHES_PER_IMAGE_PAIR_FILTERED BF_MAX_MATCHES_PER_IMAGE_PAIR_FILTERED  // :9
#endif

namespace bfref {

inline void cuda_check(cudaError_t e, const char* what) { if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e)); }   // MLIB_CUDA_SAFE_CALL
inline void bf_check(int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString((cudaError_t)rc)); }

struct float4x4 { float m[16]; float& operator()(int r, int c) { return m[4 * r + c]; } float operator()(int r, int c) const { return m[4 * r + c]; } };
typedef BFSIFTKeyPoint SIFTKeyPoint;                              // { float2 pos; float scale; float depth; }
struct SIFTKeyPointDesc { unsigned char feature[128]; };
struct SIFTImageGPU { SIFTKeyPoint* d_keyPoints; SIFTKeyPointDesc* d_keyPointDescs; };
struct ImagePairMatch { int* d_numMatches; float* d_distances; uint2* d_keyPointIndices; };
typedef BFEntryJ EntryJ;
typedef BFCUDACachedFrame CUDACachedFrame;

// ---- SiftGPU (detection + description of one image) ----
class SiftGPU {
public:
    SiftGPU() { std::memset(&m_p, 0, sizeof(m_p)); cuda_check(cudaMalloc(&d_num, sizeof(int)), "SiftGPU"); }
    ~SiftGPU() { cudaFree(d_num); }
    void SetParams(unsigned int siftWidth, unsigned int siftHeight, bool /*enableTiming*/, unsigned int featureCountThreshold, float siftDepthMin, float siftDepthMax) {
        m_p.width = siftWidth; m_p.height = siftHeight; m_p.depthWidth = siftWidth; m_p.depthHeight = siftHeight;
        m_p.featureCountThreshold = (int)featureCountThreshold; m_p.depthMin = siftDepthMin; m_p.depthMax = siftDepthMax;
    }
    // c_siftCameraParams fields the detector reads (updateConstantSiftCameraParams, FL/OnlineBundler.cpp:46-57)
    void SetCameraParams(unsigned int depthWidth, unsigned int depthHeight, float minKeyScale) { m_p.depthWidth = depthWidth; m_p.depthHeight = depthHeight; m_p.minKeyScale = minKeyScale; }
    void InitSiftGPU() {}
    int RunSIFT(float* d_colorData, const float* d_depthData) { m_intensity = d_colorData; m_depth = d_depthData; m_num = -1; return 1; }
    unsigned int GetKeyPointsAndDescriptorsCUDA(SIFTImageGPU& siftImage, const float* d_depthData, unsigned int maxNumKeyPoints = (unsigned int)-1) {
        m_p.maxKeyPoints = maxNumKeyPoints;
        bf_check(bfSiftDetect(&m_p, m_intensity, d_depthData ? d_depthData : m_depth, siftImage.d_keyPoints, reinterpret_cast<uint8_t*>(siftImage.d_keyPointDescs), d_num, nullptr), "bfSiftDetect");
        cuda_check(cudaMemcpy(&m_num, d_num, sizeof(int), cudaMemcpyDeviceToHost), "GetKeyPointsAndDescriptorsCUDA");
        if ((unsigned)m_num > maxNumKeyPoints) m_num = (int)maxNumKeyPoints;
        return (unsigned int)m_num;
    }
    int GetFeatureNum() const { return m_num; }
private:
    BFSiftDetectParams m_p; float* m_intensity = nullptr; const float* m_depth = nullptr; int* d_num = nullptr; int m_num = -1;
};

// ---- SiftMatchGPU (one image pair per call, as the reference drives it) ----
class SiftMatchGPU {
public:
    explicit SiftMatchGPU(int max_sift = 4096) : m_max(max_sift) { m_num[0] = m_num[1] = 0; m_des[0] = m_des[1] = nullptr; }
    void InitSiftMatch() {}
    void SetDescriptors(int index, int num, unsigned char* d_descriptors, int /*id*/ = -1) { if (num > m_max) num = m_max; m_num[index] = num; m_des[index] = d_descriptors; }
    void GetSiftMatch(int /*max_match*/, ImagePairMatch& imagePairMatch, uint2 keyPointOffset, float distmax = 0.7f, float ratiomax = 0.8f, int /*mutual_best_match*/ = 1) {
        BFSiftMatchJob j;
        j.d_des1 = m_des[0]; j.num1 = m_num[0]; j.d_des2 = m_des[1]; j.num2 = m_num[1];
        j.out.d_numMatches = imagePairMatch.d_numMatches; j.out.d_distances = imagePairMatch.d_distances; j.out.d_keyPointIndices = reinterpret_cast<uint32_t*>(imagePairMatch.d_keyPointIndices);
        j.keyPointOffset[0] = keyPointOffset.x; j.keyPointOffset[1] = keyPointOffset.y;
        bf_check(bfSiftMatchBatch(&j, 1, distmax, ratiomax), "bfSiftMatchBatch");
    }
private:
    int m_max; int m_num[2]; unsigned char* m_des[2];
};

// ---- SIFTImageManager ----
class SIFTImageManager {
public:
    SIFTImageManager(unsigned int maxImages = 500, unsigned int maxKeyPointsPerImage = 4096) : m_maxNumImages(maxImages), m_maxKeyPointsPerImage(maxKeyPointsPerImage) { alloc(); }
    ~SIFTImageManager() { for (void* p : m_owned) cudaFree(p); }

    SIFTImageGPU& getImageGPU(unsigned int i) { return m_SIFTImagesGPU[i]; }
    const SIFTImageGPU& getImageGPU(unsigned int i) const { return m_SIFTImagesGPU[i]; }
    unsigned int getNumImages() const { return (unsigned int)m_SIFTImagesGPU.size(); }
    unsigned int getNumKeyPointsPerImage(unsigned int i) const { return m_numKeyPointsPerImage[i]; }
    unsigned int getMaxNumKeyPointsPerImage() const { return m_maxKeyPointsPerImage; }
    unsigned int getTotalNumKeyPoints() const { return m_numKeyPoints; }
    SIFTImageGPU& createSIFTImageGPU() {                            // cpp:44-60
        assert(m_SIFTImagesGPU.size() < m_maxNumImages);
        SIFTImageGPU g; g.d_keyPoints = d_keyPoints + m_numKeyPoints; g.d_keyPointDescs = d_keyPointDescs + m_numKeyPoints;
        m_SIFTImagesGPU.push_back(g); m_bFinalizedGPUImage = false;
        return m_SIFTImagesGPU.back();
    }
    void finalizeSIFTImageGPU(unsigned int numKeyPoints) {          // cpp:62-75
        m_numKeyPointsPerImagePrefixSum.push_back(m_numKeyPoints); m_numKeyPoints += numKeyPoints; m_numKeyPointsPerImage.push_back(numKeyPoints);
        m_bFinalizedGPUImage = true; m_currentImage = (unsigned int)m_SIFTImagesGPU.size() - 1;
    }
    ImagePairMatch& getImagePairMatch(unsigned int prevImageIdx, unsigned int curImageIdx, uint2& keyPointOffset) {      // cpp:77-83
        keyPointOffset = make_uint2(m_numKeyPointsPerImagePrefixSum[prevImageIdx], m_numKeyPointsPerImagePrefixSum[curImageIdx]);
        return m_currImagePairMatches[prevImageIdx];
    }
    void reset() {                                                   // h:112-124
        m_SIFTImagesGPU.clear(); m_numKeyPointsPerImage.clear(); m_numKeyPointsPerImagePrefixSum.clear(); m_numKeyPoints = 0; m_globNumResiduals = 0; m_bFinalizedGPUImage = false;
        cuda_check(cudaMemset(d_globNumResiduals, 0, sizeof(int)), "reset");
        m_validImages.assign(m_maxNumImages, 0); m_validImages[0] = 1;
    }

    void SortKeyPointMatchesCU(unsigned int curFrame, unsigned int startFrame, unsigned int numFrames) {
        bf_check(bfSiftSortKeyPointMatches(curFrame, startFrame, numFrames, d_currNumMatchesPerImagePair, d_currMatchDistances, reinterpret_cast<uint32_t*>(d_currMatchKeyPointIndices)), "SortKeyPointMatchesCU");
    }
    void FilterKeyPointMatchesCU(unsigned int curFrame, unsigned int startFrame, unsigned int numFrames, const float4x4& siftIntrinsicsInv, unsigned int minNumMatches, float maxKabschRes2) {
        bf_check(bfSiftFilterKeyPointMatches(curFrame, startFrame, numFrames, d_keyPoints, d_currNumMatchesPerImagePair, d_currMatchDistances, reinterpret_cast<const uint32_t*>(d_currMatchKeyPointIndices),
                                             d_currNumFilteredMatchesPerImagePair, d_currFilteredMatchDistances, reinterpret_cast<uint32_t*>(d_currFilteredMatchKeyPointIndices),
                                             reinterpret_cast<float*>(d_currFilteredTransforms), reinterpret_cast<float*>(d_currFilteredTransformsInv), siftIntrinsicsInv.m, minNumMatches, maxKabschRes2), "FilterKeyPointMatchesCU");
    }
    void FilterMatchesBySurfaceAreaCU(unsigned int curFrame, unsigned int startFrame, unsigned int numFrames, const float4x4& colorIntrinsicsInv, float areaThresh) {
        bf_check(bfSiftFilterMatchesBySurfaceArea(curFrame, startFrame, numFrames, d_keyPoints, d_currNumFilteredMatchesPerImagePair, reinterpret_cast<const uint32_t*>(d_currFilteredMatchKeyPointIndices),
                                                  colorIntrinsicsInv.m, areaThresh, nullptr), "FilterMatchesBySurfaceAreaCU");
    }
    void FilterMatchesByDenseVerifyCU(unsigned int curFrame, unsigned int startFrame, unsigned int numFrames, unsigned int imageWidth, unsigned int imageHeight, const float4x4& intrinsics,
                                      const CUDACachedFrame* d_cachedFrames, float distThresh, float normalThresh, float colorThresh, float errThresh, float corrThresh, float sensorDepthMin, float sensorDepthMax) {
        bf_check(bfSiftFilterMatchesByDenseVerify(curFrame, startFrame, numFrames, imageWidth, imageHeight, intrinsics.m, d_currNumFilteredMatchesPerImagePair, reinterpret_cast<const float*>(d_currFilteredTransforms),
                                                  d_cachedFrames, distThresh, normalThresh, colorThresh, errThresh, corrThresh, sensorDepthMin, sensorDepthMax, nullptr), "FilterMatchesByDenseVerifyCU");
    }
    int VerifyTrajectoryCU(unsigned int numImages, float4x4* d_trajectory, unsigned int imageWidth, unsigned int imageHeight, const float4x4& intrinsics, const CUDACachedFrame* d_cachedFrames,
                           float distThresh, float normalThresh, float colorThresh, float errThresh, float corrThresh, float sensorDepthMin, float sensorDepthMax) {
        if (numImages < 2) return 0;
        updateGPUValidImages();
        bf_check(bfSiftVerifyTrajectory(numImages, d_validImages, reinterpret_cast<const float*>(d_trajectory), imageWidth, imageHeight, intrinsics.m, d_cachedFrames, distThresh, normalThresh, colorThresh,
                                        errThresh, corrThresh, sensorDepthMin, sensorDepthMax, d_validOpt, nullptr), "VerifyTrajectoryCU");
        int valid = 0;
        cuda_check(cudaMemcpy(&valid, d_validOpt, sizeof(int), cudaMemcpyDeviceToHost), "VerifyTrajectoryCU");
        return valid;
    }
    void AddCurrToResidualsCU(unsigned int curFrame, unsigned int startFrame, unsigned int numFrames, const float4x4& colorIntrinsicsInv) {
        bf_check(bfSiftAddCurrToResiduals(curFrame, startFrame, numFrames, d_globMatches, reinterpret_cast<uint32_t*>(d_globMatchesKeyPointIndices), d_globNumResiduals, d_currNumFilteredMatchesPerImagePair,
                                          reinterpret_cast<const uint32_t*>(d_currFilteredMatchKeyPointIndices), d_keyPoints, colorIntrinsicsInv.m), "AddCurrToResidualsCU");
        cuda_check(cudaMemcpy(&m_globNumResiduals, d_globNumResiduals, sizeof(unsigned int), cudaMemcpyDeviceToHost), "AddCurrToResidualsCU");       // SIFTImageManager.cu:680-683
    }
    void InvalidateImageToImageCU(const uint2& imageToImageIdx) { bf_check(bfSiftInvalidateImageToImage(d_globMatches, m_globNumResiduals, imageToImageIdx.x, imageToImageIdx.y), "InvalidateImageToImageCU"); }
    void CheckForInvalidFramesSimpleCU(const int* d_varToCorrNumEntriesPerRow, unsigned int numVars) { checkInvalid(d_varToCorrNumEntriesPerRow, numVars, 0); }
    void CheckForInvalidFramesCU(const int* d_varToCorrNumEntriesPerRow, unsigned int numVars) { checkInvalid(d_varToCorrNumEntriesPerRow, numVars, 1); }

    unsigned int filterFrames(unsigned int curFrame, unsigned int startFrame, unsigned int numFrames) {     // cpp:551-575
        if (numFrames == 0) return (unsigned int)-1;
        updateGPUValidImages();
        bf_check(bfSiftFilterFrames(curFrame, startFrame, numFrames, d_currNumFilteredMatchesPerImagePair, d_validImages, d_lastMatched), "filterFrames");
        int last = -1;
        cuda_check(cudaMemcpy(&last, d_lastMatched, sizeof(int), cudaMemcpyDeviceToHost), "filterFrames");
        m_validImages[curFrame] = last >= 0 ? 1 : 0;
        return (unsigned int)last;
    }
    // cpp:413-476 (device version: nothing but the new key count returns to the host)
    void fuseToGlobal(SIFTImageManager* global, const float4x4& colorIntrinsics, const float4x4* d_transforms, const float4x4& /*colorIntrinsicsInv*/) const {
        const unsigned int n = getNumImages();
        std::vector<int> counts(m_numKeyPointsPerImage.begin(), m_numKeyPointsPerImage.end());
        // this class packs key points by a prefix sum; the fuse kernel addresses image * stride + key, so hand it a stride layout view when every image sits at its
        // stride already (stride = capacity) or re-pack into scratch otherwise
        SIFTKeyPoint* d_k = nullptr; SIFTKeyPointDesc* d_d = nullptr; int* d_cnt = nullptr; uint2* d_idx = nullptr;
        unsigned int stride = 1;                                     // the kernel bounds numImages * stride by 16384: use the chunk's largest image, not the capacity
        for (int c : counts) if ((unsigned)c > stride) stride = (unsigned)c;
        if (m_globNumResiduals > 4096) throw std::runtime_error("fuseToGlobal: more than 4096 chunk correspondences");
        cuda_check(cudaMalloc(&d_k, sizeof(SIFTKeyPoint) * (size_t)n * stride), "fuseToGlobal"); cuda_check(cudaMalloc(&d_d, sizeof(SIFTKeyPointDesc) * (size_t)n * stride), "fuseToGlobal");
        cuda_check(cudaMalloc(&d_cnt, sizeof(int) * n), "fuseToGlobal"); cuda_check(cudaMalloc(&d_idx, sizeof(uint2) * (m_globNumResiduals + 1)), "fuseToGlobal");
        std::vector<uint2> idx(m_globNumResiduals);
        if (m_globNumResiduals) cuda_check(cudaMemcpy(idx.data(), d_globMatchesKeyPointIndices, sizeof(uint2) * m_globNumResiduals, cudaMemcpyDeviceToHost), "fuseToGlobal");
        auto restride = [&](unsigned int g) { unsigned int i = 0; while (i + 1 < n && m_numKeyPointsPerImagePrefixSum[i + 1] <= g) ++i; return i * stride + (g - m_numKeyPointsPerImagePrefixSum[i]); };
        for (auto& p : idx) { p.x = restride(p.x); p.y = restride(p.y); }
        if (m_globNumResiduals) cuda_check(cudaMemcpy(d_idx, idx.data(), sizeof(uint2) * m_globNumResiduals, cudaMemcpyHostToDevice), "fuseToGlobal");
        for (unsigned int i = 0; i < n; ++i) {
            cuda_check(cudaMemcpy(d_k + (size_t)i * stride, m_SIFTImagesGPU[i].d_keyPoints, sizeof(SIFTKeyPoint) * counts[i], cudaMemcpyDeviceToDevice), "fuseToGlobal");
            cuda_check(cudaMemcpy(d_d + (size_t)i * stride, m_SIFTImagesGPU[i].d_keyPointDescs, sizeof(SIFTKeyPointDesc) * counts[i], cudaMemcpyDeviceToDevice), "fuseToGlobal");
        }
        cuda_check(cudaMemcpy(d_cnt, counts.data(), sizeof(int) * n, cudaMemcpyHostToDevice), "fuseToGlobal");
        SIFTImageGPU& cur = global->createSIFTImageGPU();
        int* d_out = nullptr; cuda_check(cudaMalloc(&d_out, sizeof(int)), "fuseToGlobal");
        bf_check(bfSiftFuseToGlobal(d_globMatches, reinterpret_cast<const uint32_t*>(d_idx), d_globNumResiduals, reinterpret_cast<const float*>(d_transforms), n, d_k, reinterpret_cast<const uint8_t*>(d_d), d_cnt,
                                    stride, colorIntrinsics.m, m_globNumResiduals ? m_globNumResiduals : 1, cur.d_keyPoints, reinterpret_cast<uint8_t*>(cur.d_keyPointDescs), d_out,
                                    global->getMaxNumKeyPointsPerImage(), nullptr), "bfSiftFuseToGlobal");
        int numKeys = 0;
        cuda_check(cudaMemcpy(&numKeys, d_out, sizeof(int), cudaMemcpyDeviceToHost), "fuseToGlobal");
        global->finalizeSIFTImageGPU((unsigned int)numKeys);
        cudaFree(d_k); cudaFree(d_d); cudaFree(d_cnt); cudaFree(d_idx); cudaFree(d_out);
    }

    const std::vector<int>& getValidImages() const { return m_validImages; }
    void invalidateFrame(unsigned int frame) { m_validImages[frame] = 0; }
    void updateGPUValidImages() { if (getNumImages()) cuda_check(cudaMemcpy(d_validImages, m_validImages.data(), sizeof(int) * getNumImages(), cudaMemcpyHostToDevice), "updateGPUValidImages"); }
    const int* getValidImagesGPU() const { return d_validImages; }
    unsigned int getNumGlobalCorrespondences() const { return m_globNumResiduals; }
    EntryJ* getGlobalCorrespondencesGPU() { return d_globMatches; }
    const float4x4* getFiltTransformsToWorldGPU() const { return d_currFilteredTransformsInv; }
    const int* getNumFiltMatchesGPU() const { return d_currNumFilteredMatchesPerImagePair; }
    bool getTopRetryImage(unsigned int& idx) { if (m_imagesToRetry.empty()) return false; idx = m_imagesToRetry.front(); m_imagesToRetry.pop_front(); return true; }
    void addToRetryList(unsigned int idx) { m_imagesToRetry.push_front(idx); }
    unsigned int getCurrentFrame() const { return m_currentImage; }
    void setCurrentFrame(unsigned int idx) { m_currentImage = idx; }

private:
    template <class T> T* dev(size_t n) { T* p = nullptr; cuda_check(cudaMalloc(&p, sizeof(T) * (n ? n : 1)), "SIFTImageManager::alloc"); m_owned.push_back(p); return p; }
    void alloc() {                                                    // cpp:262-311
        m_numKeyPoints = 0;
        d_keyPoints = dev<SIFTKeyPoint>((size_t)m_maxNumImages * m_maxKeyPointsPerImage); d_keyPointDescs = dev<SIFTKeyPointDesc>((size_t)m_maxNumImages * m_maxKeyPointsPerImage);
        m_currImagePairMatches.resize(m_maxNumImages);
        d_currNumMatchesPerImagePair = dev<int>(m_maxNumImages); d_currMatchDistances = dev<float>((size_t)m_maxNumImages * MAX_MATCHES_PER_IMAGE_PAIR_RAW);
        d_currMatchKeyPointIndices = dev<uint2>((size_t)m_maxNumImages * MAX_MATCHES_PER_IMAGE_PAIR_RAW);
        d_currNumFilteredMatchesPerImagePair = dev<int>(m_maxNumImages); d_currFilteredMatchDistances = dev<float>((size_t)m_maxNumImages * MAX_MATCHES_PER_IMAGE_PAIR_FILTERED);
        d_currFilteredMatchKeyPointIndices = dev<uint2>((size_t)m_maxNumImages * MAX_MATCHES_PER_IMAGE_PAIR_FILTERED);
        d_currFilteredTransforms = dev<float4x4>(m_maxNumImages); d_currFilteredTransformsInv = dev<float4x4>(m_maxNumImages);
        m_validImages.assign(m_maxNumImages, 0); m_validImages[0] = 1;
        d_validImages = dev<int>(m_maxNumImages);
        cuda_check(cudaMemcpy(d_validImages, m_validImages.data(), sizeof(int), cudaMemcpyHostToDevice), "alloc");
        size_t maxResiduals = (size_t)MAX_MATCHES_PER_IMAGE_PAIR_FILTERED * ((size_t)m_maxNumImages * (m_maxNumImages - 1)) / 2;
        m_globNumResiduals = 0;
        d_globNumResiduals = dev<int>(1); cuda_check(cudaMemset(d_globNumResiduals, 0, sizeof(int)), "alloc");
        d_globMatches = dev<EntryJ>(maxResiduals); d_globMatchesKeyPointIndices = dev<uint2>(maxResiduals);
        d_validOpt = dev<int>(1); d_lastMatched = dev<int>(1);
        for (unsigned int r = 0; r < m_maxNumImages; ++r) {           // initializeMatching, cpp:356-364
            m_currImagePairMatches[r].d_numMatches = d_currNumMatchesPerImagePair + r;
            m_currImagePairMatches[r].d_distances = d_currMatchDistances + (size_t)r * MAX_MATCHES_PER_IMAGE_PAIR_RAW;
            m_currImagePairMatches[r].d_keyPointIndices = d_currMatchKeyPointIndices + (size_t)r * MAX_MATCHES_PER_IMAGE_PAIR_RAW;
        }
    }
    void checkInvalid(const int* d_rows, unsigned int numVars, int comprehensive) {      // SIFTImageManager.cu:724-790: flags round-trip through the host around the kernel
        updateGPUValidImages();
        bf_check(bfSiftCheckForInvalidFrames(d_rows, d_validImages, numVars, d_globMatches, m_globNumResiduals, comprehensive), "CheckForInvalidFrames");
        cuda_check(cudaMemcpy(m_validImages.data(), d_validImages, sizeof(int) * numVars, cudaMemcpyDeviceToHost), "CheckForInvalidFrames");
    }

    std::vector<SIFTImageGPU> m_SIFTImagesGPU; bool m_bFinalizedGPUImage = false;
    unsigned int m_numKeyPoints = 0; std::vector<unsigned int> m_numKeyPointsPerImage, m_numKeyPointsPerImagePrefixSum;
    SIFTKeyPoint* d_keyPoints = nullptr; SIFTKeyPointDesc* d_keyPointDescs = nullptr;
    std::vector<ImagePairMatch> m_currImagePairMatches;
    int* d_currNumMatchesPerImagePair = nullptr; float* d_currMatchDistances = nullptr; uint2* d_currMatchKeyPointIndices = nullptr;
    int* d_currNumFilteredMatchesPerImagePair = nullptr; float* d_currFilteredMatchDistances = nullptr; uint2* d_currFilteredMatchKeyPointIndices = nullptr;
    float4x4* d_currFilteredTransforms = nullptr; float4x4* d_currFilteredTransformsInv = nullptr;
    std::vector<int> m_validImages; int* d_validImages = nullptr;
    unsigned int m_globNumResiduals = 0; int* d_globNumResiduals = nullptr; EntryJ* d_globMatches = nullptr; uint2* d_globMatchesKeyPointIndices = nullptr;
    int* d_validOpt = nullptr; int* d_lastMatched = nullptr;
    unsigned int m_maxNumImages, m_maxKeyPointsPerImage, m_currentImage = 0;
    std::list<unsigned int> m_imagesToRetry;
    std::vector<void*> m_owned;
};

}  // namespace bfref
