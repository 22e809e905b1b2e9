// ref_kabsch_host.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own host-callable code: filterKeyPointMatches /
// kabsch / covarianceSVD (FL/SiftGPU/cuda_kabsch.h -- `__host__` when not compiled by nvcc, :417-421), MYEIGEN::eigenSystem
// (FL/SiftGPU/cuda_SVD.h:17-20) and computeEigenValues (FL/SiftGPU/cuda_EigenValue.h:9-39).  Compiled with g++ by oracle/build_ref.py from
// the scratch copy (-> oracle/_ref/libref_kabsch_host.so); runs on the CPU, so the oracle's Kabsch filter and eigen code can be pinned
// against the reference's code without a GPU.  This file contains no reference code.
#include <cstring>
#include <xmmintrin.h>          // _mm_rsqrt_ss: the reference's host rsqrt (cuda_svd3.h:35-41) expects MSVC's <intrin.h> to have brought it in

#include "SiftGPU/cuda_kabsch.h"

#define REF_API extern "C" __attribute__((visibility("default")))

// filterKeyPointMatches, cuda_kabsch.h:417-502.  idx [numRaw][2], dist [numRaw]: in (sorted by distance) / out (filtered first); T: 16 floats out
REF_API unsigned refHostFilterKeyPointMatches(const float* keyPoints /*[K][4]*/, unsigned* idx, float* dist, unsigned numRaw, float* T, const float* colorIntrinsicsInv,
                                              unsigned minNumMatches, float maxKabschRes2) {
    float4x4 Ki, Tm;
    for (int i = 0; i < 16; ++i) Ki.entries[i] = colorIntrinsicsInv[i];
    Tm.setIdentity();
    const unsigned n = filterKeyPointMatches(reinterpret_cast<const SIFTKeyPoint*>(keyPoints), reinterpret_cast<volatile uint2*>(idx), dist, numRaw, Tm, Ki, minNumMatches, maxKabschRes2);
    for (int i = 0; i < 16; ++i) T[i] = Tm.entries[i];
    return n;
}

// kabsch, cuda_kabsch.h:73-211
REF_API void refHostKabsch(const float* src, const float* tgt, unsigned n, float* T, float* evs) {
    float3 s[MAX_MATCHES_PER_IMAGE_PAIR_FILTERED], t[MAX_MATCHES_PER_IMAGE_PAIR_FILTERED];
    for (unsigned i = 0; i < n; ++i) { s[i] = make_float3(src[3 * i], src[3 * i + 1], src[3 * i + 2]); t[i] = make_float3(tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2]); }
    float3 e;
    const float4x4 Tm = kabsch(s, t, n, e);
    for (int i = 0; i < 16; ++i) T[i] = Tm.entries[i];
    evs[0] = e.x; evs[1] = e.y; evs[2] = e.z;
}

// MYEIGEN::eigenSystem, cuda_SVD.h:17-20: returns 1 on convergence; evs[3], ev[3][3] as the reference hands them back
REF_API int refHostEigenSystem(const float* m9, float* evs, float* ev9) {
    float3x3 m;
    for (int i = 0; i < 9; ++i) m.entries[i] = m9[i];
    float3 e, e0, e1, e2;
    const bool ok = MYEIGEN::eigenSystem(m, e, e0, e1, e2);
    evs[0] = e.x; evs[1] = e.y; evs[2] = e.z;
    const float3 v[3] = { e0, e1, e2 };
    for (int i = 0; i < 3; ++i) { ev9[3 * i] = v[i].x; ev9[3 * i + 1] = v[i].y; ev9[3 * i + 2] = v[i].z; }
    return ok ? 1 : 0;
}

// computeEigenValues (symmetric 3x3, closed form), cuda_EigenValue.h:9-39
REF_API void refHostEigenValues3(const float* m9, float* evs) {
    float3x3 m;
    for (int i = 0; i < 9; ++i) m.entries[i] = m9[i];
    const float3 e = computeEigenValues(m);
    evs[0] = e.x; evs[1] = e.y; evs[2] = e.z;
}

// float4x4::getInverse, cuda_SimpleMatrixUtil.h:980-1100 (what setLastRigidTransform calls on the host, FL/DepthSensing/CUDASceneRepHashSDF.h:130)
REF_API void refHostFloat4x4Inverse(const float* m16, float* out16) {
    float4x4 m;
    for (int i = 0; i < 16; ++i) m.entries[i] = m16[i];
    const float4x4 r = m.getInverse();
    for (int i = 0; i < 16; ++i) out16[i] = r.entries[i];
}
