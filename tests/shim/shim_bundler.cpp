// A chunk of Bundler written against the reference's class surface (bf_reference_classes.hpp): what FL/Bundler.cpp:91-249 does per frame
// (detectFeatures, matchAndFilter) and FL/Bundler.cpp:384-390 per chunk (fuseToGlobal), with the call sequence a maintainer's code has.
// tests/test_reference_classes_shim.py compiles and links it with g++ (CPU: "does the surface compile and resolve against the library") and,
// on a GPU, runs it on synthetic frames and compares every output with the same sequence driven through the C-ABI from Python.
//
// usage: shim_bundler <in.bin> <out.bin>
//   in : u32 W, H, nFrames, maxKeys; f32 colorIntrinsics[16], colorIntrinsicsInv[16]; per frame f32 intensity[H*W], f32 depth[H*W]
//   out: per frame i32 numKeys, i32 lastMatched, u32 numGlobalCorr; then per frame keys (numKeys * 16 B) + descriptors (numKeys * 128 B);
//        EntryJ[numGlobalCorr]; i32 fusedKeys; fused keys + descriptors; i32 verifyTrajectory verdict is NOT run here (needs a cache; see touch_rest)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "bf_reference_classes.hpp"

using namespace bfref;

struct Frames {
    uint32_t W, H, n, maxKeys;
    float4x4 K, Kinv;
    std::vector<float> intensity, depth;
};

static bool load(const char* path, Frames& f) {
    FILE* fp = std::fopen(path, "rb");
    if (!fp) return false;
    uint32_t hdr[4];
    if (std::fread(hdr, 4, 4, fp) != 4) return false;
    f.W = hdr[0]; f.H = hdr[1]; f.n = hdr[2]; f.maxKeys = hdr[3];
    if (std::fread(f.K.m, 4, 16, fp) != 16 || std::fread(f.Kinv.m, 4, 16, fp) != 16) return false;
    size_t px = (size_t)f.W * f.H;
    f.intensity.resize(px * f.n); f.depth.resize(px * f.n);
    for (uint32_t i = 0; i < f.n; ++i) {
        if (std::fread(&f.intensity[px * i], 4, px, fp) != px || std::fread(&f.depth[px * i], 4, px, fp) != px) return false;
    }
    std::fclose(fp);
    return true;
}

// Bundler::detectFeatures (FL/Bundler.cpp:91-100)
static unsigned int detectFeatures(SiftGPU& sift, SIFTImageManager& manager, float* d_intensitySift, const float* d_inputDepthFilt, unsigned int maxKeys) {
    SIFTImageGPU& cur = manager.createSIFTImageGPU();
    int success = sift.RunSIFT(d_intensitySift, d_inputDepthFilt);
    if (!success) throw std::runtime_error("Error running SIFT detection");
    unsigned int numKeypoints = sift.GetKeyPointsAndDescriptorsCUDA(cur, d_inputDepthFilt, maxKeys);
    manager.finalizeSIFTImageGPU(numKeypoints);
    return numKeypoints;
}

// Bundler::matchAndFilter (FL/Bundler.cpp:141-249) without the dense-verify stage (touch_rest below has it)
static int matchAndFilter(SiftMatchGPU& matcher, SIFTImageManager& siftManager, const float4x4& siftIntrinsicsInv, float matchThresh, float ratioMax, unsigned int minNumMatches,
                          float maxKabschRes2, float areaThresh) {
    const unsigned int numFrames = siftManager.getNumImages();
    const unsigned int curFrame = siftManager.getCurrentFrame();
    const unsigned int startFrame = 0;
    if (numFrames <= 1) return -1;
    int num2 = (int)siftManager.getNumKeyPointsPerImage(curFrame);
    if (num2 == 0) return -1;
    for (unsigned int prev = startFrame; prev < numFrames; ++prev) {
        if (prev == curFrame) continue;
        uint2 keyPointOffset = make_uint2(0, 0);
        ImagePairMatch& imagePairMatch = siftManager.getImagePairMatch(prev, curFrame, keyPointOffset);
        SIFTImageGPU& prevImage = siftManager.getImageGPU(prev);
        SIFTImageGPU& curImage = siftManager.getImageGPU(curFrame);
        int num1 = (int)siftManager.getNumKeyPointsPerImage(prev);
        if (num1 == 0 || num2 == 0) { cudaMemset(imagePairMatch.d_numMatches, 0, sizeof(int)); continue; }
        matcher.SetDescriptors(0, num1, (unsigned char*)prevImage.d_keyPointDescs);
        matcher.SetDescriptors(1, num2, (unsigned char*)curImage.d_keyPointDescs);
        matcher.GetSiftMatch(num1, imagePairMatch, keyPointOffset, matchThresh, ratioMax);
    }
    siftManager.SortKeyPointMatchesCU(curFrame, startFrame, numFrames);
    siftManager.FilterKeyPointMatchesCU(curFrame, startFrame, numFrames, siftIntrinsicsInv, minNumMatches, maxKabschRes2);
    siftManager.FilterMatchesBySurfaceAreaCU(curFrame, startFrame, numFrames, siftIntrinsicsInv, areaThresh);
    unsigned int lastMatchedFrame = siftManager.filterFrames(curFrame, startFrame, numFrames);
    if (lastMatchedFrame != (unsigned int)-1) siftManager.AddCurrToResidualsCU(curFrame, startFrame, numFrames, siftIntrinsicsInv);
    return (int)lastMatchedFrame;
}

// the members the run above does not reach: compiled and linked by the CPU test, never called without a cache
void touch_rest(SIFTImageManager& m, const CUDACachedFrame* d_cached, float4x4* d_traj, const int* d_rows) {
    float4x4 K{};
    m.FilterMatchesByDenseVerifyCU(1, 0, 2, 80, 60, K, d_cached, 0.15f, 0.97f, 0.1f, 0.075f, 0.02f, 0.1f, 4.0f);
    (void)m.VerifyTrajectoryCU(2, d_traj, 80, 60, K, d_cached, 0.15f, 0.97f, 0.1f, 0.05f, 0.001f, 0.1f, 4.0f);
    m.InvalidateImageToImageCU(make_uint2(0, 1));
    m.CheckForInvalidFramesSimpleCU(d_rows, 2);
    m.CheckForInvalidFramesCU(d_rows, 2);
    m.invalidateFrame(1); m.addToRetryList(1);
    unsigned int idx = 0; (void)m.getTopRetryImage(idx);
    (void)m.getValidImagesGPU(); (void)m.getFiltTransformsToWorldGPU(); (void)m.getNumFiltMatchesGPU(); (void)m.getTotalNumKeyPoints();
    m.reset();
}

template <class T> static void put(FILE* fp, const T* p, size_t n) { if (n && std::fwrite(p, sizeof(T), n, fp) != n) throw std::runtime_error("short write"); }
template <class T> static std::vector<T> fetch(const T* d, size_t n) {
    std::vector<T> h(n);
    if (n) cuda_check(cudaMemcpy(h.data(), d, sizeof(T) * n, cudaMemcpyDeviceToHost), "fetch");
    return h;
}

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    try {
        Frames f;
        if (!load(argv[1], f)) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
        const size_t px = (size_t)f.W * f.H;
        float *d_int = nullptr, *d_dep = nullptr;
        cuda_check(cudaMalloc(&d_int, 4 * px), "malloc"); cuda_check(cudaMalloc(&d_dep, 4 * px), "malloc");

        SiftGPU sift;
        sift.SetParams(f.W, f.H, false, 150, 0.1f, 4.0f);
        sift.SetCameraParams(f.W, f.H, 3.0f);
        sift.InitSiftGPU();
        SiftMatchGPU matcher((int)f.maxKeys);
        matcher.InitSiftMatch();
        SIFTImageManager local(f.n + 1, f.maxKeys), global(4, f.maxKeys);

        FILE* out = std::fopen(argv[2], "wb");
        if (!out) return 2;
        for (uint32_t i = 0; i < f.n; ++i) {
            cuda_check(cudaMemcpy(d_int, &f.intensity[px * i], 4 * px, cudaMemcpyHostToDevice), "upload");
            cuda_check(cudaMemcpy(d_dep, &f.depth[px * i], 4 * px, cudaMemcpyHostToDevice), "upload");
            int numKeys = (int)detectFeatures(sift, local, d_int, d_dep, f.maxKeys);
            int last = matchAndFilter(matcher, local, f.Kinv, 0.7f, 0.8f, 5, 0.0004f, 0.032f);
            unsigned int nCorr = local.getNumGlobalCorrespondences();
            put(out, &numKeys, 1); put(out, &last, 1); put(out, &nCorr, 1);
        }
        for (uint32_t i = 0; i < f.n; ++i) {
            unsigned int k = local.getNumKeyPointsPerImage(i);
            auto keys = fetch(local.getImageGPU(i).d_keyPoints, k);
            auto des = fetch(local.getImageGPU(i).d_keyPointDescs, k);
            put(out, keys.data(), k); put(out, des.data(), k);
        }
        auto corr = fetch(local.getGlobalCorrespondencesGPU(), local.getNumGlobalCorrespondences());
        put(out, corr.data(), corr.size());

        // Bundler::fuseToGlobal with the chunk's poses (identity here: the frames are generated from one camera position when the test says so)
        std::vector<float4x4> poses(f.n);
        for (auto& p : poses) { std::memset(p.m, 0, sizeof(p.m)); p(0, 0) = p(1, 1) = p(2, 2) = p(3, 3) = 1.0f; }
        float4x4* d_poses = nullptr;
        cuda_check(cudaMalloc(&d_poses, sizeof(float4x4) * f.n), "malloc");
        cuda_check(cudaMemcpy(d_poses, poses.data(), sizeof(float4x4) * f.n, cudaMemcpyHostToDevice), "upload");
        local.fuseToGlobal(&global, f.K, d_poses, f.Kinv);
        int fused = (int)global.getNumKeyPointsPerImage(0);
        put(out, &fused, 1);
        auto gk = fetch(global.getImageGPU(0).d_keyPoints, (size_t)fused);
        auto gd = fetch(global.getImageGPU(0).d_keyPointDescs, (size_t)fused);
        put(out, gk.data(), gk.size()); put(out, gd.data(), gd.size());
        std::fclose(out);
        cudaFree(d_int); cudaFree(d_dep); cudaFree(d_poses);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "shim_bundler: %s\n", e.what());
        return 1;
    }
    return 0;
}
