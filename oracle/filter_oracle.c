/*
 * filter_oracle.c -- CPU restatement of the Kabsch match filter (first filter):
 * FilterKeyPointMatchesCU_Kernel FL/SiftGPU/SIFTImageManager.cu:186-263 and filterKeyPointMatches / ComputeReprojection /
 * sortKabschResiduals / addMatch / getKeySourceAndTargetPoints / kabsch / covarianceSVD FL/SiftGPU/cuda_kabsch.h:110-502,
 * computeEigenValues FL/SiftGPU/cuda_EigenValue.h:9-39.   (FL/ = FriedLiver/Source/.)
 *
 * TEST INFRASTRUCTURE ONLY (see tsdf_oracle.c header).  PARITY STATUS: "parity unpinned" against the reference itself (no tests /
 * golden vectors; SIFTImageManager.cu is not rebuilt by oracle/build_ref.py); pinned by tests/test_filter_oracle.py (float64 numpy
 * Kabsch via np.linalg.svd on planted inlier / outlier sets).
 *
 * What is restated literally: the greedy sequential control flow (which raw match is tried when, the 5-pixel proximity rule, the
 * residual sort by pairwise exchange, the remove-until-below-threshold loop with its "removing made it worse" escape, the validity
 * rule on three condition numbers, the 25-match cap) and the closed-form symmetric eigenvalues.
 * What is NOT the reference's code: the 3x3 SVD inside kabsch().  The reference calls a Numerical-Recipes-style svdcmp (cuda_SVD.h);
 * any correct SVD yields the same rotation when det(W U^T) > 0, so this file (and bundlefusion_b200/csrc/sift_filter.cu, operation for
 * operation) uses a cyclic-Jacobi SVD.  One documented difference: for a reflection (det < 0) the reference flips the THIRD column of
 * its unsorted decomposition, here the column of the SMALLEST singular value is flipped (the textbook Kabsch); such pairs fail the
 * residual test in practice.
 * Arithmetic: binary32, every operation individually rounded (-ffp-contract=off here, -fmad=false there); sqrtf exact; acosf / cosf
 * come from libm here and from CUDA there (last-bit differences, visible only in a condition number sitting exactly at 100).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define MAX_RAW 128
#define MAX_FILTERED 25                 /* MAX_MATCHES_PER_IMAGE_PAIR_FILTERED, FL/GlobalDefines.h:9 */
#define KABSCH_CONDITION_THRESH 100.0f  /* cuda_kabsch.h:231 */

typedef struct { float px, py, scale, depth; } KeyPoint;       /* SIFTKeyPoint, FL/SiftGPU/SIFTImageManager.h:22-26 */
typedef struct { float x, y, z; } f3;

/* cuda_EigenValue.h:9-39: eigenvalues of a symmetric 3x3, e0 >= e1 >= e2 */
static void sym_eigenvalues(const float a[9], float e[3]) {
    const float PI = 3.14159265f;
    float p = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    if (p == 0.0f) { e[0] = a[0]; e[1] = a[4]; e[2] = a[8]; return; }
    const float q = (a[0] + a[4] + a[8]) / 3.0f;
    p = (a[0] - q) * (a[0] - q) + (a[4] - q) * (a[4] - q) + (a[8] - q) * (a[8] - q) + 2.0f * p;
    p = sqrtf(p / 6.0f);
    float B[9];
    for (int k = 0; k < 9; ++k) B[k] = (a[k] - ((k % 4 == 0) ? q : 0.0f)) * (1.0f / p);
    const float det = B[0] * (B[4] * B[8] - B[5] * B[7]) - B[1] * (B[3] * B[8] - B[5] * B[6]) + B[2] * (B[3] * B[7] - B[4] * B[6]);
    const float r = det / 2.0f;
    float phi;
    if (r <= -1.0f) phi = PI / 3.0f; else if (r >= 1.0f) phi = 0.0f; else phi = acosf(r) / 3.0f;
    e[0] = q + 2.0f * p * cosf(phi);
    e[2] = q + 2.0f * p * cosf(phi + PI * (2.0f / 3.0f));
    e[1] = 3.0f * q - e[0] - e[2];
}

/* cyclic-Jacobi SVD of a 3x3 (row-major): H = U diag(s) V^T, s descending, U and V orthonormal (see header) */
static void svd3(const float H[9], float U[9], float s[3], float V[9]) {
    float A[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[3 * i + j] = H[i] * H[j] + H[3 + i] * H[3 + j] + H[6 + i] * H[6 + j];      /* H^T H */
    for (int k = 0; k < 9; ++k) V[k] = (k % 4 == 0) ? 1.0f : 0.0f;
    static const int P[3] = { 0, 0, 1 }, Q[3] = { 1, 2, 2 };
    for (int sweep = 0; sweep < 10; ++sweep)
        for (int r = 0; r < 3; ++r) {
            const int p = P[r], q = Q[r];
            const float apq = A[3 * p + q];
            if (fabsf(apq) <= 1e-30f) continue;
            const float theta = (A[3 * q + q] - A[3 * p + p]) / (2.0f * apq);
            const float t = ((theta >= 0.0f) ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
            const float c = 1.0f / sqrtf(t * t + 1.0f), sn = t * c;
            for (int k = 0; k < 3; ++k) { const float akp = A[3 * k + p], akq = A[3 * k + q]; A[3 * k + p] = c * akp - sn * akq; A[3 * k + q] = sn * akp + c * akq; }
            for (int k = 0; k < 3; ++k) { const float apk = A[3 * p + k], aqk = A[3 * q + k]; A[3 * p + k] = c * apk - sn * aqk; A[3 * q + k] = sn * apk + c * aqk; }
            for (int k = 0; k < 3; ++k) { const float vkp = V[3 * k + p], vkq = V[3 * k + q]; V[3 * k + p] = c * vkp - sn * vkq; V[3 * k + q] = sn * vkp + c * vkq; }
        }
    float ev[3] = { A[0], A[4], A[8] };
    int ord[3] = { 0, 1, 2 };
    for (int i = 0; i < 2; ++i) for (int j = i + 1; j < 3; ++j) if (ev[ord[j]] > ev[ord[i]]) { const int tmp = ord[i]; ord[i] = ord[j]; ord[j] = tmp; }
    float Vs[9];
    for (int c = 0; c < 3; ++c) { s[c] = sqrtf(fmaxf(ev[ord[c]], 0.0f)); for (int k = 0; k < 3; ++k) Vs[3 * k + c] = V[3 * k + ord[c]]; }
    memcpy(V, Vs, sizeof Vs);
    /* U columns: H v / s where s is significant, completed to a right-handed orthonormal basis otherwise */
    float u[3][3];
    const float tiny = 1e-7f * s[0];
    for (int c = 0; c < 3; ++c) {
        if (s[c] > tiny && s[c] > 0.0f) {
            for (int k = 0; k < 3; ++k) u[c][k] = (H[3 * k] * V[c] + H[3 * k + 1] * V[3 + c] + H[3 * k + 2] * V[6 + c]) / s[c];
        } else if (c == 2) {
            u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1]; u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2]; u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
        } else if (c == 1) {       /* any unit vector orthogonal to u0 */
            const float ax = fabsf(u[0][0]), ay = fabsf(u[0][1]), az = fabsf(u[0][2]);
            float e[3] = { 0, 0, 0 }; e[(ax <= ay && ax <= az) ? 0 : ((ay <= az) ? 1 : 2)] = 1.0f;
            float w[3] = { u[0][1] * e[2] - u[0][2] * e[1], u[0][2] * e[0] - u[0][0] * e[2], u[0][0] * e[1] - u[0][1] * e[0] };
            const float l = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            for (int k = 0; k < 3; ++k) u[1][k] = w[k] / l;
        } else { u[0][0] = 1.0f; u[0][1] = 0.0f; u[0][2] = 0.0f; }
    }
    for (int c = 0; c < 3; ++c) for (int k = 0; k < 3; ++k) U[3 * k + c] = u[c][k];
}

static float det3(const float m[9]) { return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]); }

/* kabsch(), cuda_kabsch.h:110-176: T (4x4 row-major) with T src ~ tgt; evs = singular values of the covariance, descending */
static void kabsch(const f3* src, const f3* tgt, unsigned n, float T[16], float evs[3]) {
    float p0[3] = { 0, 0, 0 }, q0[3] = { 0, 0, 0 };
    for (unsigned i = 0; i < n; ++i) { p0[0] += src[i].x; p0[1] += src[i].y; p0[2] += src[i].z; q0[0] += tgt[i].x; q0[1] += tgt[i].y; q0[2] += tgt[i].z; }
    for (int k = 0; k < 3; ++k) { p0[k] /= (float)n; q0[k] /= (float)n; }
    float H[9] = { 0 };
    for (unsigned i = 0; i < n; ++i) {
        const float p[3] = { src[i].x - p0[0], src[i].y - p0[1], src[i].z - p0[2] }, q[3] = { tgt[i].x - q0[0], tgt[i].y - q0[1], tgt[i].z - q0[2] };
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] += p[r] * q[c];
    }
    for (int k = 0; k < 9; ++k) H[k] /= (float)n;
    float U[9], V[9];
    svd3(H, U, evs, V);
    /* R = V D U^T, D = diag(1, 1, det(V U^T)) */
    float VUt[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) VUt[3 * r + c] = V[3 * r] * U[3 * c] + V[3 * r + 1] * U[3 * c + 1] + V[3 * r + 2] * U[3 * c + 2];
    const float d = (det3(VUt) < 0.0f) ? -1.0f : 1.0f;
    float R[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[3 * r + c] = V[3 * r] * U[3 * c] + V[3 * r + 1] * U[3 * c + 1] + (V[3 * r + 2] * d) * U[3 * c + 2];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[4 * r + c] = R[3 * r + c];
        T[4 * r + 3] = q0[r] - (R[3 * r] * p0[0] + R[3 * r + 1] * p0[1] + R[3 * r + 2] * p0[2]);
    }
    T[12] = T[13] = T[14] = 0.0f; T[15] = 1.0f;
}
/* covarianceSVD(), cuda_kabsch.h:178-198 */
static void covariance_eigs(const f3* pts, unsigned n, float e[3]) {
    float p0[3] = { 0, 0, 0 };
    for (unsigned i = 0; i < n; ++i) { p0[0] += pts[i].x; p0[1] += pts[i].y; p0[2] += pts[i].z; }
    for (int k = 0; k < 3; ++k) p0[k] /= (float)n;
    float C[9] = { 0 };
    for (unsigned i = 0; i < n; ++i) {
        const float p[3] = { pts[i].x - p0[0], pts[i].y - p0[1], pts[i].z - p0[2] };
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C[3 * r + c] += p[r] * p[c];
    }
    for (int k = 0; k < 9; ++k) C[k] /= (float)n;
    sym_eigenvalues(C, e);
}
/* ComputeReprojection(), cuda_kabsch.h:381-414 */
static int compute_reprojection(f3* src, f3* tgt, unsigned n, float* res, float T[16], uint32_t* idx /*[.][2]*/, float* dist) {
    float evs[3];
    kabsch(src, tgt, n, T, evs);
    for (unsigned i = 0; i < n; ++i) {
        const float dx = (T[0] * src[i].x + T[1] * src[i].y + T[2] * src[i].z + T[3]) - tgt[i].x;
        const float dy = (T[4] * src[i].x + T[5] * src[i].y + T[6] * src[i].z + T[7]) - tgt[i].y;
        const float dz = (T[8] * src[i].x + T[9] * src[i].y + T[10] * src[i].z + T[11]) - tgt[i].z;
        res[i] = dx * dx + dy * dy + dz * dz;
    }
    for (unsigned i = 0; i < n; ++i)                /* sortKabschResiduals, :368-377 */
        for (unsigned j = i; j < n; ++j)
            if (res[i] > res[j]) {
                float t = res[i]; res[i] = res[j]; res[j] = t;
                f3 s = src[i]; src[i] = src[j]; src[j] = s;
                s = tgt[i]; tgt[i] = tgt[j]; tgt[j] = s;
                uint32_t a = idx[2 * i], b = idx[2 * i + 1]; idx[2 * i] = idx[2 * j]; idx[2 * i + 1] = idx[2 * j + 1]; idx[2 * j] = a; idx[2 * j + 1] = b;
                t = dist[i]; dist[i] = dist[j]; dist[j] = t;
            }
    const float c1 = evs[0] / evs[1];
    float e[3];
    covariance_eigs(src, n, e); const float cp = e[0] / e[1];
    covariance_eigs(tgt, n, e); const float cq = e[0] / e[1];
    if (c1 != c1 || cp != cp || cq != cq || fabsf(c1) > KABSCH_CONDITION_THRESH || fabsf(cp) > KABSCH_CONDITION_THRESH || fabsf(cq) > KABSCH_CONDITION_THRESH) return 0;
    return 1;
}
static int add_match(uint32_t ax, uint32_t ay, const KeyPoint* kp, const uint32_t* idx, unsigned cur) {         /* addMatch, :233-247 */
    for (unsigned i = 0; i < cur; ++i) {
        const float dix = kp[ax].px - kp[idx[2 * i]].px, diy = kp[ax].py - kp[idx[2 * i]].py;
        const float djx = kp[ay].px - kp[idx[2 * i + 1]].px, djy = kp[ay].py - kp[idx[2 * i + 1]].py;
        if (sqrtf(dix * dix + diy * diy) <= 5.0f || sqrtf(djx * djx + djy * djy) <= 5.0f) return 0;
    }
    return 1;
}
static void key_points_3d(const KeyPoint* kp, const uint32_t* idx, unsigned n, f3* src, f3* tgt, const float* Ki) {   /* getKeySourceAndTargetPoints, :249-321 */
    for (unsigned i = 0; i < n; ++i)
        for (int s = 0; s < 2; ++s) {
            const KeyPoint* k = &kp[idx[2 * i + s]];
            const float v[3] = { k->depth * k->px, k->depth * k->py, k->depth * 1.0f };
            f3 o = { Ki[0] * v[0] + Ki[1] * v[1] + Ki[2] * v[2] + Ki[3], Ki[4] * v[0] + Ki[5] * v[1] + Ki[6] * v[2] + Ki[7], Ki[8] * v[0] + Ki[9] * v[1] + Ki[10] * v[2] + Ki[11] };
            if (s == 0) src[i] = o; else tgt[i] = o;
        }
}

/* filterKeyPointMatches, cuda_kabsch.h:417-502.  idx / dist: the pair's raw matches (sorted by distance), modified in place; returns the
 * number of filtered matches (their indices / distances in the first slots), T = the transform estimate. */
static unsigned filter_pair(const KeyPoint* kp, uint32_t* idx, float* dist, unsigned numRaw, float T[16], const float* Ki, unsigned minNum, float maxRes2) {
    f3 src[MAX_FILTERED], tgt[MAX_FILTERED];
    float res[MAX_FILTERED];
    unsigned i0 = 0, cur = 0;
    float curMax = 100.0f;
    int valid = 0;
    for (int k = 0; k < 16; ++k) T[k] = (k % 5 == 0) ? 1.0f : 0.0f;
    for (;;) {
        if (i0 == numRaw || cur >= MAX_FILTERED) {
            if (cur < minNum || curMax >= maxRes2 || !valid) cur = 0;
            break;
        } else if (add_match(idx[2 * i0], idx[2 * i0 + 1], kp, idx, cur)) {
            idx[2 * cur] = idx[2 * i0]; idx[2 * cur + 1] = idx[2 * i0 + 1]; dist[cur] = dist[i0];
            ++cur;
            if (cur >= 3) {
                key_points_3d(kp, idx, cur, src, tgt, Ki);
                valid = compute_reprojection(src, tgt, cur, res, T, idx, dist);
                const int b = valid;
                float prevT[16]; memcpy(prevT, T, sizeof prevT);
                curMax = res[cur - 1];
                if (curMax > maxRes2) {
                    float lastRes = -1.0f;
                    for (int i = (int)cur - 1; i >= 3; --i) {
                        lastRes = res[i];
                        --cur;
                        valid = compute_reprojection(src, tgt, cur, res, T, idx, dist);
                        curMax = res[cur - 1];
                        if (cur == 3 && (curMax > maxRes2 || (b && !valid))) { ++cur; curMax = lastRes; valid = b; memcpy(T, prevT, sizeof prevT); break; }
                        if (curMax < maxRes2) break;
                    }
                }
            }
        }
        ++i0;
    }
    return cur;
}

void orc_mat4_inverse(const float* m, float* out);          /* solver_oracle.c */

/* FilterKeyPointMatchesCU_Kernel over pairs [startFrame, numFrames) \\ {curFrame}; arrays in the manager's layout */
ORC_API void orc_sift_filter_matches(unsigned curFrame, unsigned startFrame, unsigned numFrames, const KeyPoint* kp, const int32_t* numMatches,
                                     const float* dists, const uint32_t* idxs, int32_t* numFiltered, float* fDists, uint32_t* fIdxs, float* fT, float* fTinv,
                                     const float* siftIntrinsicsInv, unsigned minNum, float maxRes2) {
    for (unsigned p = startFrame; p < numFrames; ++p) {
        if (p == curFrame) continue;
        unsigned n = (unsigned)(numMatches[p] < MAX_RAW ? numMatches[p] : MAX_RAW);
        if (numMatches[p] <= 0) { numFiltered[p] = 0; continue; }
        uint32_t idx[2 * MAX_RAW]; float dist[MAX_RAW];
        memcpy(idx, idxs + 2 * (size_t)p * MAX_RAW, sizeof(uint32_t) * 2 * n); memcpy(dist, dists + (size_t)p * MAX_RAW, sizeof(float) * n);
        float T[16];
        const unsigned c = filter_pair(kp, idx, dist, n, T, siftIntrinsicsInv, minNum, maxRes2);
        numFiltered[p] = (int32_t)c;
        memcpy(fT + 16 * (size_t)p, T, sizeof T);
        orc_mat4_inverse(T, fTinv + 16 * (size_t)p);
        for (unsigned k = 0; k < MAX_FILTERED; ++k) {
            if (k < c) { fDists[p * MAX_FILTERED + k] = dist[k]; fIdxs[2 * (p * MAX_FILTERED + k)] = idx[2 * k]; fIdxs[2 * (p * MAX_FILTERED + k) + 1] = idx[2 * k + 1]; }
            else { fDists[p * MAX_FILTERED + k] = 999.0f; fIdxs[2 * (p * MAX_FILTERED + k)] = 0xFFFFFFFFu; fIdxs[2 * (p * MAX_FILTERED + k) + 1] = 0xFFFFFFFFu; }
        }
    }
}

/* AddCurrToResidualsCU_Kernel (FL/SiftGPU/SIFTImageManager.cu:610-655) with the pairs appended in ascending order (the reference's
 * atomicAdd makes the pair order race-dependent).  entries: EntryJ = { uint32 i, uint32 j, float pos_i[3], float pos_j[3] } (32 bytes). */
ORC_API int orc_sift_add_residuals(unsigned curFrame, unsigned startFrame, unsigned numFrames, uint8_t* entries, uint32_t* entryIdx, int numResiduals,
                                   const int32_t* numFiltered, const uint32_t* fIdxs, const KeyPoint* kp, const float* Ki) {
    for (unsigned p = startFrame; p < numFrames; ++p) {
        if (p == curFrame) continue;
        for (int k = 0; k < numFiltered[p]; ++k) {
            const uint32_t* ij = &fIdxs[2 * ((size_t)p * MAX_FILTERED + k)];
            f3 s, t;
            key_points_3d(kp, ij, 1, &s, &t, Ki);
            uint8_t* e = entries + 32 * (size_t)numResiduals;
            const uint32_t ii = p, jj = curFrame;
            memcpy(e, &ii, 4); memcpy(e + 4, &jj, 4); memcpy(e + 8, &s, 12); memcpy(e + 20, &t, 12);
            entryIdx[2 * numResiduals] = ij[0]; entryIdx[2 * numResiduals + 1] = ij[1];
            ++numResiduals;
        }
    }
    return numResiduals;
}
