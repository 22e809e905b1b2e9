// trajectory_host.cu -- host-side TrajectoryManager behind the C-ABI of include/bf_bundler.h (row a22, host part).
// Behavioural source (what, not how): FL/TrajectoryManager.{h,cpp}.  Pure host code (no kernel): the reference class is host C++ too;
// it lives in a .cu only to share the SE(3) logarithm of se3.cuh with the device code (same operation order on both sides; this TU
// is built -fmad=false like the other bit-comparable ones).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>
#include <mutex>
#include <vector>

#include "../../include/bf_bundler.h"
#include "bf_common.cuh"
#include "se3.cuh"

struct BFTrajectoryManager {
    struct Frame { int type; unsigned frameIdx; float integrated[16]; float dist; };
    std::mutex mtx;
    std::vector<float> optimized;           // [numMax][16]   (m_optmizedTransforms)
    std::vector<Frame> frames;              // m_frames
    std::vector<Frame*> sorted;             // m_framesSort
    unsigned numAdded = 0, numOptimized = 0;
    std::list<Frame*> toDeIntegrate, toIntegrate, toReIntegrate;
    unsigned topN = 0; float minPoseDist = 0.0f, rescale = 2.0f;
    float* opt(unsigned i) { return &optimized[16 * (size_t)i]; }
    const float* opt(unsigned i) const { return &optimized[16 * (size_t)i]; }
};
using TM = BFTrajectoryManager;

static void invalidate_frame(TM* tm, unsigned idx) {                      // TrajectoryManager.cpp:201-210
    TM::Frame& f = tm->frames[idx];
    if (f.type == BF_TRAJ_INVALID) return;
    const int before = f.type;
    f.type = BF_TRAJ_INVALID;
    if (before == BF_TRAJ_INTEGRATED) tm->toDeIntegrate.push_back(&f);
}

BF_API BFTrajectoryManager* bfTrajectoryCreate(unsigned int numMaxImage, unsigned int topNActive, float minPoseDistSqrt) {   // cpp:7-21
    TM* tm = new TM();
    tm->optimized.assign(16 * (size_t)numMaxImage, 0.0f);
    tm->frames.resize(numMaxImage);
    for (unsigned i = 0; i < numMaxImage; ++i) {
        TM::Frame& f = tm->frames[i];
        f.type = BF_TRAJ_NOT_INTEGRATED_NO_TRANSFORM; f.frameIdx = 0xFFFFFFFFu; f.dist = 0.0f;
        for (int k = 0; k < 16; ++k) f.integrated[k] = -INFINITY;
    }
    tm->sorted.reserve(numMaxImage);
    tm->topN = topNActive; tm->minPoseDist = minPoseDistSqrt; tm->rescale = 2.0f;
    return tm;
}
BF_API void bfTrajectoryDestroy(BFTrajectoryManager* tm) { delete tm; }

BF_API void bfTrajectoryAddFrame(BFTrajectoryManager* tm, int type, const float* transform, unsigned int idx) {             // cpp:23-31
    TM::Frame& f = tm->frames[idx];
    f.type = type; f.frameIdx = idx;
    std::memcpy(f.integrated, transform, sizeof f.integrated);
    std::memcpy(tm->opt(idx), transform, 16 * sizeof(float));
    tm->sorted.push_back(&f);
    tm->numAdded++;
}

BF_API void bfTrajectoryUpdateOptimizedTransform(BFTrajectoryManager* tm, const float* h_trajectory, unsigned int numFrames) {    // cpp:33-43
    std::lock_guard<std::mutex> lk(tm->mtx);
    tm->numOptimized = numFrames;
    numFrames = std::min(numFrames, tm->numAdded);
    std::memcpy(tm->optimized.data(), h_trajectory, sizeof(float) * 16 * (size_t)numFrames);
}

BF_API void bfTrajectoryGenerateUpdateLists(BFTrajectoryManager* tm) {                                                       // cpp:45-108
    std::lock_guard<std::mutex> lk(tm->mtx);
    const unsigned numFrames = std::min(tm->numOptimized, tm->numAdded);
    for (unsigned i = 0; i < numFrames; ++i) {
        TM::Frame& f = tm->frames[i];
        if (tm->opt(i)[0] == -INFINITY) { invalidate_frame(tm, i); continue; }
        if (f.type == BF_TRAJ_NOT_INTEGRATED_NO_TRANSFORM || f.type == BF_TRAJ_INVALID) {
            f.type = BF_TRAJ_NOT_INTEGRATED_WITH_TRANSFORM;
            tm->toIntegrate.push_back(&f);
        }
        bf::V3 ro, to, ri, ti;
        bf::matrix_to_pose(tm->opt(i), ro, to);
        bf::matrix_to_pose(f.integrated, ri, ti);
        // PoseHelper::MatrixToPose packs (translation, rotation) (FL/PoseHelper.h:355-358) and components 0..2 are the ones scaled
        // (FL/TrajectoryManager.cpp:67-74): m_featureRescaleRotToTrans ends up on the Lie TRANSLATION
        const bf::V3 dt = ti * tm->rescale - to * tm->rescale, dr = ri - ro;
        // point6d operator| (mLib core-math/point6d.h, un-vendored submodule; published form: the six products summed left to right)
        f.dist = dt.x * dt.x + dt.y * dt.y + dt.z * dt.z + dr.x * dr.x + dr.y * dr.y + dr.z * dr.z;
    }
    std::stable_sort(tm->sorted.begin(), tm->sorted.begin() + numFrames, [](const TM::Frame* l, const TM::Frame* r) {
        if (l->type == BF_TRAJ_INTEGRATED && r->type != BF_TRAJ_INTEGRATED) return true;
        if (l->type != BF_TRAJ_INTEGRATED) return false;
        // a frame integrated with an invalid pose has a NaN distance; ordered last so that the comparator stays a strict weak order
        const float dl = std::isnan(l->dist) ? -INFINITY : l->dist, dr = std::isnan(r->dist) ? -INFINITY : r->dist;
        return dl > dr;
    });
    for (unsigned i = (unsigned)tm->toReIntegrate.size(); i < tm->topN && i < numFrames; ++i) {
        TM::Frame* f = tm->sorted[i];
        if (f->dist > tm->minPoseDist && f->type == BF_TRAJ_INTEGRATED) { f->type = BF_TRAJ_REINTEGRATION; tm->toReIntegrate.push_back(f); }
        else break;
    }
}

BF_API void bfTrajectoryConfirmIntegration(BFTrajectoryManager* tm, unsigned int frameIdx) { tm->frames[frameIdx].type = BF_TRAJ_INTEGRATED; }   // cpp:110-114

BF_API int bfTrajectoryGetTopFromReIntegrateList(BFTrajectoryManager* tm, float* oldT, float* newT, unsigned int* frameIdx) {     // cpp:116-136
    if (tm->toReIntegrate.empty()) return 0;
    std::lock_guard<std::mutex> lk(tm->mtx);
    while (!tm->toReIntegrate.empty()) {
        TM::Frame* f = tm->toReIntegrate.front();
        std::memcpy(newT, tm->opt(f->frameIdx), 16 * sizeof(float));
        *frameIdx = f->frameIdx;
        std::memcpy(oldT, f->integrated, 16 * sizeof(float));
        tm->toReIntegrate.pop_front();
        if (newT[0] != -INFINITY) { std::memcpy(f->integrated, newT, 16 * sizeof(float)); break; }
    }
    return 1;
}
BF_API int bfTrajectoryGetTopFromIntegrateList(BFTrajectoryManager* tm, float* T, unsigned int* frameIdx) {                   // cpp:138-153
    if (tm->toIntegrate.empty()) return 0;
    std::lock_guard<std::mutex> lk(tm->mtx);
    TM::Frame* f = tm->toIntegrate.front();
    std::memcpy(T, tm->opt(f->frameIdx), 16 * sizeof(float));
    *frameIdx = f->frameIdx;
    std::memcpy(f->integrated, T, 16 * sizeof(float));
    tm->toIntegrate.pop_front();
    return 1;
}
BF_API int bfTrajectoryGetTopFromDeIntegrateList(BFTrajectoryManager* tm, float* T, unsigned int* frameIdx) {                 // cpp:155-166
    if (tm->toDeIntegrate.empty()) return 0;
    std::lock_guard<std::mutex> lk(tm->mtx);
    TM::Frame* f = tm->toDeIntegrate.front();
    std::memcpy(T, f->integrated, 16 * sizeof(float));
    *frameIdx = f->frameIdx;
    tm->toDeIntegrate.pop_front();
    return 1;
}
BF_API unsigned int bfTrajectoryGetNumOptimizedFrames(const BFTrajectoryManager* tm) { return tm->numOptimized; }
BF_API unsigned int bfTrajectoryGetNumAddedFrames(const BFTrajectoryManager* tm) { return tm->numAdded; }
BF_API unsigned int bfTrajectoryGetNumActiveOperations(const BFTrajectoryManager* tm) {                                        // cpp:193-199
    return (unsigned)(tm->toDeIntegrate.size() + tm->toIntegrate.size() + tm->toReIntegrate.size());
}
BF_API int bfTrajectoryGetFrameType(const BFTrajectoryManager* tm, unsigned int frameIdx) { return tm->frames[frameIdx].type; }
BF_API float bfTrajectoryGetFrameDist(const BFTrajectoryManager* tm, unsigned int frameIdx) { return tm->frames[frameIdx].dist; }
BF_API unsigned int bfTrajectoryGetOptimizedTransforms(BFTrajectoryManager* tm, float* h_out) {                                // h:49-67
    std::lock_guard<std::mutex> lk(tm->mtx);
    const unsigned n = std::min(tm->numAdded, tm->numOptimized);
    for (unsigned i = 0; i < n; ++i) {
        if (tm->frames[i].type == BF_TRAJ_INVALID) for (int k = 0; k < 16; ++k) h_out[16 * (size_t)i + k] = -INFINITY;
        else std::memcpy(&h_out[16 * (size_t)i], tm->opt(i), 16 * sizeof(float));
    }
    return n;
}
