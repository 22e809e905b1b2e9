// ref_trajectory_host.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own TrajectoryManager (FL/TrajectoryManager.{h,cpp}) and
// the Lie part of FL/PoseHelper.h, compiled by g++ from the scratch copy against mLib's own math types (oracle/ref_traj_stubs/stdafx.h)
// (oracle/build_ref.py -> oracle/_ref/libref_trajectory_host.so).  Runs on the CPU.  This file contains no reference code.
#include "stdafx.h"
#include "TrajectoryManager.h"
#include "GlobalAppState.h"

#define REF_API extern "C" __attribute__((visibility("default")))

static mat4f to_m(const float* p) { mat4f m; std::memcpy(m.matrix, p, sizeof m.matrix); return m; }

REF_API void* refTrajCreate(unsigned numMaxImage, unsigned topNActive, float minPoseDistSqrt) {
    GlobalAppState::get().s_topNActive = topNActive; GlobalAppState::get().s_minPoseDistSqrt = minPoseDistSqrt;
    return new TrajectoryManager(numMaxImage);
}
REF_API void refTrajDestroy(void* h) { delete static_cast<TrajectoryManager*>(h); }
REF_API void refTrajAddFrame(void* h, int type, const float* T, unsigned idx) { static_cast<TrajectoryManager*>(h)->addFrame((TrajectoryManager::TrajectoryFrame::TYPE)type, to_m(T), idx); }
REF_API void refTrajUpdateOptimizedTransform(void* h, const float* traj, unsigned n) { static_cast<TrajectoryManager*>(h)->updateOptimizedTransform(reinterpret_cast<const float4x4*>(traj), n); }
REF_API void refTrajGenerateUpdateLists(void* h) { static_cast<TrajectoryManager*>(h)->generateUpdateLists(); }
REF_API void refTrajConfirmIntegration(void* h, unsigned i) { static_cast<TrajectoryManager*>(h)->confirmIntegration(i); }
REF_API int refTrajGetTopFromReIntegrateList(void* h, float* o, float* n, unsigned* idx) {
    mat4f a, b; const bool r = static_cast<TrajectoryManager*>(h)->getTopFromReIntegrateList(a, b, *idx);
    if (r) { std::memcpy(o, a.matrix, 64); std::memcpy(n, b.matrix, 64); } return r;
}
REF_API int refTrajGetTopFromIntegrateList(void* h, float* t, unsigned* idx) { mat4f a; const bool r = static_cast<TrajectoryManager*>(h)->getTopFromIntegrateList(a, *idx); if (r) std::memcpy(t, a.matrix, 64); return r; }
REF_API int refTrajGetTopFromDeIntegrateList(void* h, float* t, unsigned* idx) { mat4f a; const bool r = static_cast<TrajectoryManager*>(h)->getTopFromDeIntegrateList(a, *idx); if (r) std::memcpy(t, a.matrix, 64); return r; }
REF_API unsigned refTrajGetNumActiveOperations(void* h) { return static_cast<TrajectoryManager*>(h)->getNumActiveOperations(); }
REF_API int refTrajGetFrameType(void* h, unsigned i) { return (int)static_cast<TrajectoryManager*>(h)->getFrames()[i].type; }
REF_API float refTrajGetFrameDist(void* h, unsigned i) { return static_cast<TrajectoryManager*>(h)->getFrames()[i].dist; }
REF_API unsigned refTrajGetOptimizedTransforms(void* h, float* out) {
    std::vector<mat4f> t; static_cast<TrajectoryManager*>(h)->getOptimizedTransforms(t);
    for (size_t i = 0; i < t.size(); ++i) std::memcpy(out + 16 * i, t[i].matrix, 64);
    return (unsigned)t.size();
}
// PoseHelper::MatrixToPose (Lie build): (translation part, rotation part)
REF_API void refMatrixToPose(const float* T, float* pose6) { const Pose p = PoseHelper::MatrixToPose(to_m(T)); for (int i = 0; i < 6; ++i) pose6[i] = p[i]; }
