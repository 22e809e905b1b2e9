/*
 * trajectory_oracle.c -- CPU restatement of the trajectory glue kernels: FL/OnlineBundler.cu:6-140.
 * TEST INFRASTRUCTURE ONLY (see tsdf_oracle.c header).  PARITY STATUS: PINNED against the reference's own kernels executed on the CPU -- FL/OnlineBundler.cu compiled by g++ against the CUDA emulation (oracle/build_ref.py build_mgr_emulated -> oracle/_ref/libref_mgr_emulated.so), outputs committed as tests/golden/manager_reference_emulated.npz, tests/test_manager_reference_emulated.py: same -inf patterns and slots written, values within 2 units in the last place (the fused products below against the emulation's unfused ones).
 * Also pinned by tests/test_trajectory_oracle.py (numpy float64
 * products, identity / inverse cases).  Arithmetic contract shared with bundlefusion_b200/csrc/trajectory.cu: the 4x4 product fused
 * as fma(a4,b4, fma(a3,b3, fma(a1,b1, a2*b2))) (the way nvcc fuses a1*b1 + a2*b2 + a3*b3 + a4*b4, cf. tsdf_oracle.c header), the
 * inverse by the cofactor formula of bundlefusion_b200/csrc/mat4.cuh, every other operation individually rounded.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

void orc_mat4_inverse(const float* m, float* out);          /* solver_oracle.c */

static void mul(const float* a, const float* b, float* r) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r[4 * i + j] = fmaf(a[4 * i + 3], b[12 + j], fmaf(a[4 * i + 2], b[8 + j], fmaf(a[4 * i], b[j], a[4 * i + 1] * b[4 + j])));
}

/* getSiftTransformCU_Kernel, OnlineBundler.cu:6-53 */
ORC_API void orc_compute_sift_transform(const float* filteredInv, const int* numFiltered, const float* complete, unsigned lastValidComplete,
                                        float* siftTraj, unsigned curAll, unsigned cur, float* currIntegrate) {
    if (cur == 0) return;
    for (int i = (int)cur - 1; i >= 0; --i) {
        if (numFiltered[i] <= 0) continue;
        const unsigned prev = curAll - (cur - (unsigned)i);
        float T[16], R[16];
        mul(&siftTraj[16 * prev], &filteredInv[16 * i], T);
        memcpy(&siftTraj[16 * curAll], T, sizeof T);
        if (lastValidComplete == 0) memcpy(R, T, sizeof R);
        else if (prev < lastValidComplete) mul(&complete[16 * prev], &filteredInv[16 * i], R);
        else {
            float inv[16], off[16], t2[16];
            orc_mat4_inverse(&siftTraj[16 * lastValidComplete], inv);
            mul(inv, &siftTraj[16 * prev], off);
            mul(&complete[16 * lastValidComplete], off, t2);
            mul(t2, &filteredInv[16 * i], R);
        }
        memcpy(currIntegrate, R, sizeof R);
        break;
    }
}
/* updateTrajectoryCU_Kernel, OnlineBundler.cu:71-90 */
ORC_API void orc_update_trajectory(const float* global, float* complete, unsigned numComplete, const float* local, unsigned perTraj, const int* invalidate) {
    const unsigned submap = perTraj - 1;
    for (unsigned idx = 0; idx < numComplete; ++idx) {
        if (invalidate[idx] == 0) { for (int k = 0; k < 16; ++k) complete[16 * idx + k] = -INFINITY; }
        else mul(&global[16 * (idx / submap)], &local[16 * ((idx / submap) * perTraj + idx % submap)], &complete[16 * idx]);
    }
}
/* initNextGlobalTransformCU_Kernel, OnlineBundler.cu:114-126 */
ORC_API void orc_init_next_global(float* global, unsigned numGlobal, unsigned initIdx, const float* local, unsigned lastValidLocal, unsigned perTraj) {
    float R[16];
    mul(&global[16 * initIdx], &local[16 * (size_t)(numGlobal * perTraj - (perTraj - lastValidLocal))], R);
    memcpy(&global[16 * numGlobal], R, sizeof R);
}

void orc_matrix_to_pose(const float* M, float* rot, float* trans);          /* solver_oracle.c: SE(3) log, LieDerivUtil.h:135-158 */

/* TrajectoryManager::generateUpdateLists, the re-integration part (FL/TrajectoryManager.cpp:45-108): dist per frame, then the up to topN
 * integrated frames of largest dist > minDist, descending (ties: lower index). */
ORC_API int orc_select_reintegration(const float* opt, const float* integ, const int* state, unsigned n, unsigned topN, float minDist, float scale,
                                     float* dist, int* list) {
    for (unsigned i = 0; i < n; ++i) {
        dist[i] = -1.0f;
        if (state[i] != 0 && opt[16 * i] != -INFINITY) {
            float ro[3], to[3], ri[3], ti[3];
            orc_matrix_to_pose(&opt[16 * i], ro, to);
            orc_matrix_to_pose(&integ[16 * i], ri, ti);
            /* the host's PoseHelper::MatrixToPose packs (translation, rotation) -- FL/PoseHelper.h:355-358 -- and the scale factor is applied to
             * components 0..2 (FL/TrajectoryManager.cpp:67-74): m_featureRescaleRotToTrans multiplies the Lie TRANSLATION, not the rotation */
            float d = 0.0f, e = 0.0f;
            for (int k = 0; k < 3; ++k) { const float a = ti[k] * scale - to[k] * scale; d += a * a; }
            for (int k = 0; k < 3; ++k) { const float a = ri[k] - ro[k]; e += a * a; }
            dist[i] = d + e;
        }
    }
    int found = 0;
    char* taken = (char*)calloc(n ? n : 1, 1);
    for (; (unsigned)found < topN; ++found) {
        int bi = -1; float best = -1.0f;
        for (unsigned i = 0; i < n; ++i) if (!taken[i] && dist[i] > minDist && dist[i] >= 0.0f && dist[i] > best) { best = dist[i]; bi = (int)i; }
        if (bi < 0) break;
        taken[bi] = 1; list[found] = bi;
    }
    free(taken);
    return found;
}
