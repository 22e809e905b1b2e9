/*
 * filter_oracle.c -- CPU restatement of the Kabsch match filter (first filter):
 * FilterKeyPointMatchesCU_Kernel FL/SiftGPU/SIFTImageManager.cu:186-263 and filterKeyPointMatches / ComputeReprojection /
 * sortKabschResiduals / addMatch / getKeySourceAndTargetPoints / kabsch / covarianceSVD FL/SiftGPU/cuda_kabsch.h:110-502,
 * computeEigenValues FL/SiftGPU/cuda_EigenValue.h:9-39.   (FL/ = FriedLiver/Source/.)
 *
 * TEST INFRASTRUCTURE ONLY (see tsdf_oracle.c header).  PARITY STATUS: PINNED against the reference's own code for kabsch(), the complete
 * filterKeyPointMatches() and MYEIGEN::eigenSystem -- those functions are __host__-callable as written, oracle/build_ref.py compiles them
 * with g++ from the sources under /root/reference (oracle/ref_kabsch_host.cpp -> oracle/_ref/libref_kabsch_host.so), their outputs on
 * seeded inputs are committed as tests/golden/kabsch_reference_host.npz, and this file reproduces them bit for bit
 * (tests/test_kabsch_reference_host.py) when its rsqrt is switched to the reference's host flavour.  The reference's DEVICE build of the
 * same code differs from its host build only in that primitive (CUDA's rsqrtf, 2 ulp, instead of the 12-bit SSE estimate); the oracle's
 * default, and the CUDA path's, is 1 / sqrtf.  Also pinned by a float64 numpy Kabsch on planted inlier / outlier sets
 * (tests/test_filter_oracle.py).  Not pinned against a run of the reference's kernels on a GPU (oracle/_ref/libref_siftmgr.so is built,
 * scripts/ref_siftmgr_compare.py is ready; no hardware run yet).
 *
 * What is restated literally: the greedy sequential control flow (which raw match is tried when, the 5-pixel proximity rule, the
 * residual sort by pairwise exchange, the remove-until-below-threshold loop with its "removing made it worse" escape, the validity
 * rule on three condition numbers, the 25-match cap), the closed-form symmetric eigenvalues, and the reference's 3x3 SVD: the fast
 * approximate one of cuda_svd3.h (four fixed Jacobi sweeps with approximate Givens angles, column sort with its rho2 quirk, Givens QR),
 * with the reflection fix always on the third column.
 * Arithmetic: binary32, every operation individually rounded (-ffp-contract=off here, -fmad=false there); sqrtf exact; acosf / cosf
 * come from libm here and from CUDA there (last-bit differences, visible only in a condition number sitting exactly at 100).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
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

/* ---- the 3x3 SVD the reference's kabsch() really uses: svd() of FL/SiftGPU/cuda_svd3.h (E. Jang's CUDA version of McAdams, Selle, Tamstorf,
 * Teran, Sifakis, "Computing the Singular Value Decomposition of 3x3 matrices with minimal branching and elementary floating point
 * operations", TR1690, 2011): four fixed sweeps of Jacobi rotations with the APPROXIMATE Givens angle on A^T A, the rotation accumulated as a
 * quaternion, a sort of the columns of A V by their norms, and a Givens QR of the result.  Restated operation for operation, quirks included:
 * the norm of the second column is taken over (b12, b22, b23) -- b23 where b32 is meant (cuda_svd3.h:245) --, only four sweeps, no
 * convergence test: for rank-deficient input (three points) the result can be far from a singular value decomposition, and so is the
 * reference's.  rsqrt: the reference calls CUDA's rsqrtf on the device (2 ulp) and an SSE estimate on the host (12 bits); here it is
 * 1 / sqrtf, individually rounded -- except in the pinning mode below, where the host estimate is used so that this restatement can be
 * compared bit for bit with the reference's own host build (oracle/ref_kabsch_host.cpp, tests/test_kabsch_reference_host.py). ---- */
#include <xmmintrin.h>
static int g_rsqrtHostEstimate = 0;
ORC_API void orc_set_rsqrt_host_estimate(int on) { g_rsqrtHostEstimate = on; }
static float rsqrt_(float x) {
    if (g_rsqrtHostEstimate) { float out; _mm_store_ss(&out, _mm_rsqrt_ss(_mm_load_ss(&x))); return out; }      /* cuda_svd3.h:35-41 */
    return 1.0f / sqrtf(x);
}
static void cond_swap(int c, float* X, float* Y) { const float Z = *X; *X = c ? *Y : *X; *Y = c ? Z : *Y; }
static void cond_neg_swap(int c, float* X, float* Y) { const float Z = -*X; *X = c ? *Y : *X; *Y = c ? Z : *Y; }

static void jacobi_conjugation(int x, int y, int z, float* s11, float* s21, float* s22, float* s31, float* s32, float* s33, float* qV) {    /* cuda_svd3.h:149-204 */
    float ch = 2.0f * (*s11 - *s22), sh = *s21;                                  /* approximateGivensQuaternion, :134-147 */
    const int big = 5.828427124f * sh * sh < ch * ch;
    const float w = rsqrt_(ch * ch + sh * sh);
    ch = big ? w * ch : 0.923879532f;
    sh = big ? w * sh : 0.3826834323f;
    const float scale = ch * ch + sh * sh;
    const float a = (ch * ch - sh * sh) / scale, b = (2.0f * sh * ch) / scale;
    float t11 = *s11, t21 = *s21, t22 = *s22, t31 = *s31, t32 = *s32, t33 = *s33;
    *s11 = a * (a * t11 + b * t21) + b * (a * t21 + b * t22);
    *s21 = a * (-b * t11 + a * t21) + b * (-b * t21 + a * t22);
    *s22 = -b * (-b * t11 + a * t21) + a * (-b * t21 + a * t22);
    *s31 = a * t31 + b * t32; *s32 = -b * t31 + a * t32; *s33 = t33;
    float tmp[3] = { qV[0] * sh, qV[1] * sh, qV[2] * sh };
    sh *= qV[3];
    qV[0] *= ch; qV[1] *= ch; qV[2] *= ch; qV[3] *= ch;
    qV[z] += sh; qV[3] -= tmp[z]; qV[x] += tmp[y]; qV[y] -= tmp[x];
    t11 = *s22; t21 = *s32; t22 = *s33; t31 = *s21; t32 = *s31; t33 = *s11;      /* re-arrange for the next rotation */
    *s11 = t11; *s21 = t21; *s22 = t22; *s31 = t31; *s32 = t32; *s33 = t33;
}
static void qr_givens(float a1, float a2, float* ch, float* sh) {               /* QRGivensQuaternion, :265-281 */
    const float eps = 1e-6f, q = a1 * a1 + a2 * a2;
    const float rho = q * rsqrt_(q);                                              /* accurateSqrt */
    *sh = rho > eps ? a2 : 0.0f;
    *ch = fabsf(a1) + fmaxf(rho, eps);
    cond_swap(a1 < 0.0f, sh, ch);
    const float w = rsqrt_(*ch * *ch + *sh * *sh);
    *ch *= w; *sh *= w;
}
/* svd(A) -> U, S (upper triangular R of the QR, its diagonal = singular values up to sign), V; all row-major 3x3.  cuda_svd3.h:347-393 */
static void svd3_fast(const float A[9], float U[9], float S[9], float V[9]) {
    const float a11 = A[0], a12 = A[1], a13 = A[2], a21 = A[3], a22 = A[4], a23 = A[5], a31 = A[6], a32 = A[7], a33 = A[8];
    /* A^T A (multAtB) */
    float s11 = a11 * a11 + a21 * a21 + a31 * a31;
    float s21 = a12 * a11 + a22 * a21 + a32 * a31, s22 = a12 * a12 + a22 * a22 + a32 * a32;
    float s31 = a13 * a11 + a23 * a21 + a33 * a31, s32 = a13 * a12 + a23 * a22 + a33 * a32, s33 = a13 * a13 + a23 * a23 + a33 * a33;
    float qV[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
    for (int i = 0; i < 4; ++i) {                                                 /* jacobiEigenanlysis, :214-229 */
        jacobi_conjugation(0, 1, 2, &s11, &s21, &s22, &s31, &s32, &s33, qV);
        jacobi_conjugation(1, 2, 0, &s11, &s21, &s22, &s31, &s32, &s33, qV);
        jacobi_conjugation(2, 0, 1, &s11, &s21, &s22, &s31, &s32, &s33, qV);
    }
    /* quatToMat3, :102-131 */
    const float w = qV[3], x = qV[0], y = qV[1], z = qV[2];
    const float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
    float v11 = 1 - 2 * (qyy + qzz), v12 = 2 * (qxy - qwz), v13 = 2 * (qxz + qwy);
    float v21 = 2 * (qxy + qwz), v22 = 1 - 2 * (qxx + qzz), v23 = 2 * (qyz - qwx);
    float v31 = 2 * (qxz - qwy), v32 = 2 * (qyz + qwx), v33 = 1 - 2 * (qxx + qyy);
    /* B = A V (multAB) */
    float b11 = a11 * v11 + a12 * v21 + a13 * v31, b12 = a11 * v12 + a12 * v22 + a13 * v32, b13 = a11 * v13 + a12 * v23 + a13 * v33;
    float b21 = a21 * v11 + a22 * v21 + a23 * v31, b22 = a21 * v12 + a22 * v22 + a23 * v32, b23 = a21 * v13 + a22 * v23 + a23 * v33;
    float b31 = a31 * v11 + a32 * v21 + a33 * v31, b32 = a31 * v12 + a32 * v22 + a33 * v32, b33 = a31 * v13 + a32 * v23 + a33 * v33;
    /* sortSingularValues, :231-262 (rho2 over b12, b22, b23 as written there) */
    float rho1 = b11 * b11 + b21 * b21 + b31 * b31, rho2 = b12 * b12 + b22 * b22 + b23 * b23, rho3 = b13 * b13 + b23 * b23 + b33 * b33;
    int c = rho1 < rho2;
    cond_neg_swap(c, &b11, &b12); cond_neg_swap(c, &v11, &v12); cond_neg_swap(c, &b21, &b22); cond_neg_swap(c, &v21, &v22); cond_neg_swap(c, &b31, &b32); cond_neg_swap(c, &v31, &v32);
    cond_swap(c, &rho1, &rho2);
    c = rho1 < rho3;
    cond_neg_swap(c, &b11, &b13); cond_neg_swap(c, &v11, &v13); cond_neg_swap(c, &b21, &b23); cond_neg_swap(c, &v21, &v23); cond_neg_swap(c, &b31, &b33); cond_neg_swap(c, &v31, &v33);
    cond_swap(c, &rho1, &rho3);
    c = rho2 < rho3;
    cond_neg_swap(c, &b12, &b13); cond_neg_swap(c, &v12, &v13); cond_neg_swap(c, &b22, &b23); cond_neg_swap(c, &v22, &v23); cond_neg_swap(c, &b32, &b33); cond_neg_swap(c, &v32, &v33);
    /* QRDecomposition, :283-345 */
    float ch1, sh1, ch2, sh2, ch3, sh3, r11, r12, r13, r21, r22, r23, r31, r32, r33;
    qr_givens(b11, b21, &ch1, &sh1);
    float a = 1 - 2 * sh1 * sh1, b = 2 * ch1 * sh1;
    r11 = a * b11 + b * b21; r12 = a * b12 + b * b22; r13 = a * b13 + b * b23;
    r21 = -b * b11 + a * b21; r22 = -b * b12 + a * b22; r23 = -b * b13 + a * b23;
    r31 = b31; r32 = b32; r33 = b33;
    qr_givens(r11, r31, &ch2, &sh2);
    a = 1 - 2 * sh2 * sh2; b = 2 * ch2 * sh2;
    b11 = a * r11 + b * r31; b12 = a * r12 + b * r32; b13 = a * r13 + b * r33;
    b21 = r21; b22 = r22; b23 = r23;
    b31 = -b * r11 + a * r31; b32 = -b * r12 + a * r32; b33 = -b * r13 + a * r33;
    qr_givens(b22, b32, &ch3, &sh3);
    a = 1 - 2 * sh3 * sh3; b = 2 * ch3 * sh3;
    r11 = b11; r12 = b12; r13 = b13;
    r21 = a * b21 + b * b31; r22 = a * b22 + b * b32; r23 = a * b23 + b * b33;
    r31 = -b * b21 + a * b31; r32 = -b * b22 + a * b32; r33 = -b * b23 + a * b33;
    const float sh12 = sh1 * sh1, sh22 = sh2 * sh2, sh32 = sh3 * sh3;
    U[0] = (-1 + 2 * sh12) * (-1 + 2 * sh22);
    U[1] = 4 * ch2 * ch3 * (-1 + 2 * sh12) * sh2 * sh3 + 2 * ch1 * sh1 * (-1 + 2 * sh32);
    U[2] = 4 * ch1 * ch3 * sh1 * sh3 - 2 * ch2 * (-1 + 2 * sh12) * sh2 * (-1 + 2 * sh32);
    U[3] = 2 * ch1 * sh1 * (1 - 2 * sh22);
    U[4] = -8 * ch1 * ch2 * ch3 * sh1 * sh2 * sh3 + (-1 + 2 * sh12) * (-1 + 2 * sh32);
    U[5] = -2 * ch3 * sh3 + 4 * sh1 * (ch3 * sh1 * sh3 + ch1 * ch2 * sh2 * (-1 + 2 * sh32));
    U[6] = 2 * ch2 * sh2;
    U[7] = 2 * ch3 * (1 - 2 * sh22) * sh3;
    U[8] = (-1 + 2 * sh22) * (-1 + 2 * sh32);
    S[0] = r11; S[1] = r12; S[2] = r13; S[3] = r21; S[4] = r22; S[5] = r23; S[6] = r31; S[7] = r32; S[8] = r33;
    V[0] = v11; V[1] = v12; V[2] = v13; V[3] = v21; V[4] = v22; V[5] = v23; V[6] = v31; V[7] = v32; V[8] = v33;
}
/* svd(m, v, s) + svdAbsEV, cuda_svd3.h:436-482: U, V and the ABSOLUTE diagonal of S (the columns of U follow the sign), unsorted */
static void svd3(const float H[9], float U[9], float s[3], float V[9]) {
    float S[9];
    svd3_fast(H, U, S, V);
    s[0] = S[0]; s[1] = S[4]; s[2] = S[8];
    for (int i = 0; i < 3; ++i) if (s[i] < 0.0f) { s[i] *= -1.0f; for (int j = 0; j < 3; ++j) U[3 * j + i] *= -1.0f; }
}

/* matNxM<3,3>::det, cuda_SimpleMatrixUtil.h:1544-1559 */
static float det3(const float m[9]) { return m[0] * m[4] * m[8] + m[1] * m[5] * m[6] + m[2] * m[3] * m[7] - m[6] * m[4] * m[2] - m[7] * m[5] * m[0] - m[8] * m[3] * m[1]; }

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
    { float t; if (evs[0] < evs[1]) { t = evs[0]; evs[0] = evs[1]; evs[1] = t; } if (evs[1] < evs[2]) { t = evs[1]; evs[1] = evs[2]; evs[2] = t; }
      if (evs[0] < evs[1]) { t = evs[0]; evs[0] = evs[1]; evs[1] = t; } }          /* cuda_kabsch.h:139-141 */
    /* R = V D U^T, D = diag(1, 1, -1) when det(U V^T) < 0 (cuda_kabsch.h:185-190): always the THIRD column, whatever the order of the values */
    float UVt[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) UVt[3 * r + c] = U[3 * r] * V[3 * c] + U[3 * r + 1] * V[3 * c + 1] + U[3 * r + 2] * V[3 * c + 2];
    const float d = (det3(UVt) < 0.0f) ? -1.0f : 1.0f;
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

/* =====================================================================================================================================
 * Row a19, second and third filter.  Restated from FilterMatchesBySurfaceAreaCU_Kernel (FL/SiftGPU/SIFTImageManager.cu:318-389) with its
 * helpers computeKeyPointMatchesCovariance / computeCovariance2d / computeAreaOrientedBoundingBox2 / projectKeysToPlane
 * (FL/SiftGPU/cuda_surfaceArea.h:13-163), MYEIGEN::eigenSystem / jacobi (FL/SiftGPU/cuda_SVD.h:17-214, the cyclic Jacobi eigenvalue
 * routine of Numerical Recipes), the 2x2 closed forms of FL/SiftGPU/cuda_EigenValue.h:71-105, warpReduce{Sum,Min,Max}
 * (FL/SiftGPU/cudaUtil.h:25-43); and from FilterMatchesByDenseVerifyCU_Kernel / computeProjError (FL/SiftGPU/SIFTImageManager.cu:413-585,
 * the CUDACACHE_FLOAT_NORMALS branch -- FL/CUDACacheUtil.h:7-8 defines both macros and the float one is tested first).
 *
 * PARITY STATUS: PINNED against the reference's own kernels executed on the CPU (SIFTImageManager.cu compiled by g++ against the CUDA
 * emulation, oracle/_ref/libref_mgr_emulated.so; golden file tests/golden/manager_reference_emulated.npz,
 * tests/test_manager_reference_emulated.py): the surface-area filter's decisions under thresholds that bracket every area of this file
 * within 1e-4 relative; the dense check's decisions follow from this file's PER-PIXEL residual / weight / count once they are added up the
 * way the reference's kernel adds them.  That is not a plain sum: FilterMatchesByDenseVerifyCU_Kernel launches (width, ceil(height / 32))
 * threads, reduces each warp with `val += __shfl_down(val, offset)` (a lane whose source is past the warp's end adds ITSELF) and lets the
 * threads with threadIdx.x % 32 == 0 add their value to the block total -- for the 80 x 60 cache that is lane 0 of warps 0 - 2 and lane 16
 * of warps 2 - 4: part of the image is counted more than once, part not at all.  orc_sift_filter_dense_verify (dense_block_total) and the
 * CUDA path form the total the same way, so their decisions are the reference's (a plain sum flips one in twenty on the test inputs).
 * Restated literally: the 32-lane shuffle-down summation trees (lane 0's
 * association order), the quirk that eigenSystem hands back ROWS of the Jacobi rotation matrix as "eigenvectors" (cuda_SVD.h:94-99 --
 * an orthonormal frame, but not the eigenframe; which frame depends on the literal sweep order, hence the literal Jacobi), the magnitude
 * sort by row exchange, NaN handling (a diagonal 2-D covariance gives a 0/0 axis; the NaN coordinates lose every fminf / fmaxf against
 * the idle lanes' +-FLT_MAX, the extent is -inf and the area 0), and in the dense check the float->int conversion of the GPU
 * (NaN -> 0, saturating).
 * Not the reference's: normalize() is v * (1 / sqrtf(v.v)) here and in the CUDA path (the reference: v * rsqrtf(v.v), an approximate
 * hardware instruction); the dense check's three sums are taken in a fixed order (256 strided partial sums, shuffle-down tree per 32,
 * then the 8 tree results left to right) where the reference uses shared-memory float atomics in a race-dependent order, and the
 * inverse transform comes from the sub-determinant form of the adjugate (mat4_inverse_subdet below).
 * ===================================================================================================================================== */

static float tree_sum32(const float* v) {                    /* warpReduceSum as lane 0 sees it */
    float a[32]; memcpy(a, v, sizeof a);
    for (int off = 16; off > 0; off /= 2) for (int l = 0; l < off; ++l) a[l] = a[l] + a[l + off];
    return a[0];
}
static float tree_min32(const float* v) { float a[32]; memcpy(a, v, sizeof a); for (int off = 16; off > 0; off /= 2) for (int l = 0; l < off; ++l) a[l] = fminf(a[l], a[l + off]); return a[0]; }
static float tree_max32(const float* v) { float a[32]; memcpy(a, v, sizeof a); for (int off = 16; off > 0; off /= 2) for (int l = 0; l < off; ++l) a[l] = fmaxf(a[l], a[l + off]); return a[0]; }

/* cuda_SVD.h:112-214 (Numerical Recipes jacobi, n = 3), zero-based.  a is destroyed; d = eigenvalues, v = rotation matrix. */
static int jacobi3(float a[3][3], float d[3], float v[3][3]) {
    float b[3], z[3];
    for (int p = 0; p < 3; ++p) { for (int q = 0; q < 3; ++q) v[p][q] = 0.0f; v[p][p] = 1.0f; }
    for (int p = 0; p < 3; ++p) { b[p] = d[p] = a[p][p]; z[p] = 0.0f; }
    for (int sweep = 1; sweep <= 50; ++sweep) {
        float sm = 0.0f;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) sm += fabsf(a[p][q]);
        if (sm == 0.0f) return 1;
        const float tresh = sweep < 4 ? 0.2f * sm / 9.0f : 0.0f;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const float g = 100.0f * fabsf(a[p][q]);
                if (sweep > 4 && fabsf(d[p]) + g == fabsf(d[p]) && fabsf(d[q]) + g == fabsf(d[q])) a[p][q] = 0.0f;
                else if (fabsf(a[p][q]) > tresh) {
                    float h = d[q] - d[p], t;
                    if (fabsf(h) + g == fabsf(h)) t = a[p][q] / h;
                    else {
                        const float theta = 0.5f * h / a[p][q];
                        t = 1.0f / (fabsf(theta) + sqrtf(1.0f + theta * theta));
                        if (theta < 0.0f) t = -t;
                    }
                    const float c = 1.0f / sqrtf(1.0f + t * t), s = t * c, tau = s / (1.0f + c);
                    h = t * a[p][q];
                    z[p] -= h; z[q] += h; d[p] -= h; d[q] += h;
                    a[p][q] = 0.0f;
#define ORC_ROT(m, i, j, k, l) { const float g_ = m[i][j], h_ = m[k][l]; m[i][j] = g_ - s * (h_ + g_ * tau); m[k][l] = h_ + s * (g_ - h_ * tau); }
                    for (int j = 0; j < p; ++j) ORC_ROT(a, j, p, j, q)
                    for (int j = p + 1; j < q; ++j) ORC_ROT(a, p, j, j, q)
                    for (int j = q + 1; j < 3; ++j) ORC_ROT(a, p, j, q, j)
                    for (int j = 0; j < 3; ++j) ORC_ROT(v, j, p, j, q)
#undef ORC_ROT
                }
            }
        for (int p = 0; p < 3; ++p) { b[p] += z[p]; d[p] = b[p]; z[p] = 0.0f; }
    }
    return 0;
}

/* MYEIGEN::eigenSystem (cuda_SVD.h:17-20, 70-110): "eigenvectors" ev[i] = ROW i of the rotation matrix, rows exchanged by |eigenvalue| */
static int eigen_system3(const float m[9], float evs[3], f3 ev[3]) {
    float a[3][3], d[3], v[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = m[i + 3 * j];
    if (!jacobi3(a, d, v)) return 0;
    float e[3][3];
    for (int i = 0; i < 3; ++i) { evs[i] = d[i]; for (int j = 0; j < 3; ++j) e[i][j] = v[i][j]; }
    for (int i = 0; i < 3; ++i) {
        float curMax = 0.0f; int arg = -1;
        for (int j = i; j < 3; ++j) if (fabsf(evs[j]) > curMax) { curMax = fabsf(evs[j]); arg = j; }
        if (arg != i && arg != -1) {
            float t = evs[i]; evs[i] = evs[arg]; evs[arg] = t;
            for (int j = 0; j < 3; ++j) { t = e[i][j]; e[i][j] = e[arg][j]; e[arg][j] = t; }
        }
    }
    for (int i = 0; i < 3; ++i) { ev[i].x = e[i][0]; ev[i].y = e[i][1]; ev[i].z = e[i][2]; }
    return 1;
}

static f3 key_point_3d(const KeyPoint* k, const float* Ki) {
    const float v[3] = { k->depth * k->px, k->depth * k->py, k->depth * 1.0f };
    f3 o = { Ki[0] * v[0] + Ki[1] * v[1] + Ki[2] * v[2] + Ki[3], Ki[4] * v[0] + Ki[5] * v[1] + Ki[6] * v[2] + Ki[7], Ki[8] * v[0] + Ki[9] * v[1] + Ki[10] * v[2] + Ki[11] };
    return o;
}
static float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

/* one image of one pair: area of the 2-D oriented bounding box of the key points projected into the plane the "eigenframe" defines */
static float surface_area_one(const KeyPoint* kp, const uint32_t* idx, unsigned n, const float* Ki, unsigned which) {
    f3 pt[32]; float lane[32];
    for (unsigned i = 0; i < 32; ++i) { f3 zero = { 0.0f, 0.0f, 0.0f }; pt[i] = i < n ? key_point_3d(&kp[idx[2 * i + which]], Ki) : zero; }
    /* computeKeyPointMatchesCovariance, cuda_surfaceArea.h:13-53 */
    float mean[3], V[9];
    for (int c = 0; c < 3; ++c) { for (int i = 0; i < 32; ++i) lane[i] = ((const float*)&pt[i])[c]; mean[c] = tree_sum32(lane) / (float)n; }
    for (int j = 0; j < 9; ++j) {
        for (unsigned i = 0; i < 32; ++i) {
            const float* p = (const float*)&pt[i];
            lane[i] = i < n ? (p[j / 3] - mean[j / 3]) * (p[j % 3] - mean[j % 3]) : 0.0f;
        }
        V[j] = tree_sum32(lane) / (float)n;
    }
    float evs[3]; f3 ev[3];
    if (!eigen_system3(V, evs, ev)) return 0.0f;
    /* projectKeysToPlane, cuda_surfaceArea.h:138-161 */
    const f3 mu = { mean[0], mean[1], mean[2] };
    float px[32], py[32];
    for (unsigned i = 0; i < 32; ++i) {
        px[i] = py[i] = 0.0f;
        if (i < n) {
            const f3 dm = { pt[i].x - mu.x, pt[i].y - mu.y, pt[i].z - mu.z };
            const float k = dot3(ev[2], dm);
            const f3 s = { (pt[i].x - k * ev[2].x) - mu.x, (pt[i].y - k * ev[2].y) - mu.y, (pt[i].z - k * ev[2].z) - mu.z };
            px[i] = dot3(s, ev[0]); py[i] = dot3(s, ev[1]);
        }
    }
    /* computeCovariance2d, cuda_surfaceArea.h:59-90 */
    float m2[2], c2[4];
    for (unsigned i = 0; i < 32; ++i) lane[i] = i < n ? px[i] : 0.0f;
    m2[0] = tree_sum32(lane) / (float)n;
    for (unsigned i = 0; i < 32; ++i) lane[i] = i < n ? py[i] : 0.0f;
    m2[1] = tree_sum32(lane) / (float)n;
    for (int j = 0; j < 4; ++j) {
        for (unsigned i = 0; i < 32; ++i) { const float q[2] = { px[i] - m2[0], py[i] - m2[1] }; lane[i] = i < n ? q[j / 2] * q[j % 2] : 0.0f; }
        c2[j] = tree_sum32(lane) / (float)n;
    }
    /* computeAreaOrientedBoundingBox2, cuda_surfaceArea.h:93-136; cuda_EigenValue.h:71-105 */
    const float dd = c2[0] - c2[3];
    const float disc = 0.5f * sqrtf(dd * dd + (4.0f * c2[1]) * c2[1]);
    const float l1 = (c2[0] + c2[3]) / 2.0f + disc, l2 = (c2[0] + c2[3]) / 2.0f - disc;
    float ax[2][2];
    const float ls[2] = { l1, l2 };
    for (int k = 0; k < 2; ++k) {
        float vx = -c2[1], vy = c2[0] - ls[k];
        const float mag = sqrtf(vx * vx + vy * vy);
        vx /= mag; vy /= mag;
        const float inv = 1.0f / sqrtf(vx * vx + vy * vy);            /* normalize(): see the deviation note above */
        ax[k][0] = vx * inv; ax[k][1] = vy * inv;
    }
    float mnx[32], mny[32], mxx[32], mxy[32];
    for (unsigned i = 0; i < 32; ++i) {
        mnx[i] = mny[i] = 3.402823466e+38f; mxx[i] = mxy[i] = -3.402823466e+38f;
        if (i < n) { const float cx = ax[0][0] * px[i] + ax[0][1] * py[i], cy = ax[1][0] * px[i] + ax[1][1] * py[i]; mnx[i] = mxx[i] = cx; mny[i] = mxy[i] = cy; }
    }
    const float ex = tree_max32(mxx) - tree_min32(mnx), ey = tree_max32(mxy) - tree_min32(mny);
    if (ex < 0.00001f || ey < 0.00001f) return 0.0f;
    return ex * ey;
}

/* FilterMatchesBySurfaceAreaCU over pairs [startFrame, numFrames) \\ {curFrame}: zeroes the filtered-match count of a pair whose key points
 * cover less than areaThresh in BOTH images.  areas (optional): [numFrames][2], written for the pairs visited. */
ORC_API void orc_sift_filter_surface_area(unsigned curFrame, unsigned startFrame, unsigned numFrames, const KeyPoint* kp, int32_t* numFiltered,
                                          const uint32_t* fIdxs, const float* colorIntrinsicsInv, float areaThresh, float* areas) {
    for (unsigned p = startFrame; p < numFrames; ++p) {
        if (p == curFrame) continue;
        const int32_t c = numFiltered[p];
        if (c <= 0) continue;
        const unsigned n = (unsigned)(c < MAX_FILTERED ? c : MAX_FILTERED);
        const uint32_t* idx = fIdxs + 2 * (size_t)p * MAX_FILTERED;
        const float a0 = surface_area_one(kp, idx, n, colorIntrinsicsInv, 0), a1 = surface_area_one(kp, idx, n, colorIntrinsicsInv, 1);
        if (areas) { areas[2 * p] = a0; areas[2 * p + 1] = a1; }
        if (a0 < areaThresh && a1 < areaThresh) numFiltered[p] = 0;
    }
}

/* ---- dense verification ---- */
typedef struct { const float* depth; const float* campos; const float* intensity; const float* intensityDerivs; const uint8_t* normalsU4; const float* normals; } CachedFrame;   /* CUDACachedFrame, FL/CUDACacheUtil.h */

static int f2i_gpu(float x) {              /* (int) of the GPU: NaN -> 0, saturating */
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x;
}

/* computeProjError, SIFTImageManager.cu:413-486.  out = (residual, weight, 1) or zeros */
static void proj_error(unsigned idx, unsigned W, unsigned H, float distThresh, float normalThresh, const float* T, const float* K,
                       const CachedFrame* in, const CachedFrame* model, float dMin, float dMax, float out[3]) {
    out[0] = out[1] = out[2] = 0.0f;
    const float* p = in->campos + 4 * (size_t)idx; const float* nI = in->normals + 4 * (size_t)idx;
    const float d = in->depth[idx];
    if (!(p[0] != -INFINITY && nI[0] != -INFINITY && d >= dMin && d <= dMax)) return;
    float pt[4], nt[4];
    for (int r = 0; r < 4; ++r) {
        pt[r] = T[4 * r] * p[0] + T[4 * r + 1] * p[1] + T[4 * r + 2] * p[2] + T[4 * r + 3] * p[3];
        nt[r] = T[4 * r] * nI[0] + T[4 * r + 1] * nI[1] + T[4 * r + 2] * nI[2] + T[4 * r + 3] * 0.0f;
    }
    const float tx = K[0] * pt[0] + K[1] * pt[1] + K[2] * pt[2] + K[3], ty = K[4] * pt[0] + K[5] * pt[1] + K[6] * pt[2] + K[7], tz = K[8] * pt[0] + K[9] * pt[1] + K[10] * pt[2] + K[11];
    const int sx = f2i_gpu(roundf(tx / tz)), sy = f2i_gpu(roundf(ty / tz));
    if (!(sx >= 0 && sy >= 0 && sx < (int)W && sy < (int)H)) return;
    const size_t m = (size_t)sy * W + sx;
    const float* q = model->campos + 4 * m; const float* nT = model->normals + 4 * m;
    if (!(q[0] != -INFINITY && nT[0] != -INFINITY)) return;
    const float e0 = pt[0] - q[0], e1 = pt[1] - q[1], e2 = pt[2] - q[2], e3 = pt[3] - q[3];
    const float dist = sqrtf(e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3);
    const float dN = nt[0] * nT[0] + nt[1] * nT[1] + nt[2] * nT[2];
    const float projDepth = pt[2], tgtDepth = model->depth[m];
    if (!(tgtDepth >= dMin && tgtDepth <= dMax)) return;
    const int bad = (tgtDepth != -INFINITY && projDepth < tgtDepth) && dist > distThresh;
    if ((dN >= normalThresh && dist <= distThresh) || bad) {
        const float zN = (pt[2] - dMin) / (dMax - dMin);
        const float w = fmaxf(0.0f, 0.5f * ((1.0f - dist / distThresh) + (1.0f - zN)));
        out[0] = dist; out[1] = w; out[2] = 1.0f;
    }
}

/* The inverse the CUDA path takes (bundlefusion_b200/csrc/mat4.cuh mat4_inverse_hd): adjugate through 2x2 sub-determinants.  The reference's
 * float4x4::getInverse (cuda_SimpleMatrixUtil.h:975-1100, restated as orc_mat4_inverse) expands every cofactor into six triple products;
 * the two agree to a few ulp, which moves the sums below by ~1e-7 relative -- decisions differ only for a pair sitting on a threshold. */
static void mat4_inverse_subdet(const float* m, float* out) {
    const float a00 = m[0], a01 = m[1], a02 = m[2], a03 = m[3], a10 = m[4], a11 = m[5], a12 = m[6], a13 = m[7];
    const float a20 = m[8], a21 = m[9], a22 = m[10], a23 = m[11], a30 = m[12], a31 = m[13], a32 = m[14], a33 = m[15];
    const float s0 = a00 * a11 - a10 * a01, s1 = a00 * a12 - a10 * a02, s2 = a00 * a13 - a10 * a03;
    const float s3 = a01 * a12 - a11 * a02, s4 = a01 * a13 - a11 * a03, s5 = a02 * a13 - a12 * a03;
    const float c5 = a22 * a33 - a32 * a23, c4 = a21 * a33 - a31 * a23, c3 = a21 * a32 - a31 * a22;
    const float c2 = a20 * a33 - a30 * a23, c1 = a20 * a32 - a30 * a22, c0 = a20 * a31 - a30 * a21;
    const float det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    const float r = 1.0f / det;
    out[0] = (a11 * c5 - a12 * c4 + a13 * c3) * r;   out[1] = (-a01 * c5 + a02 * c4 - a03 * c3) * r;
    out[2] = (a31 * s5 - a32 * s4 + a33 * s3) * r;   out[3] = (-a21 * s5 + a22 * s4 - a23 * s3) * r;
    out[4] = (-a10 * c5 + a12 * c2 - a13 * c1) * r;  out[5] = (a00 * c5 - a02 * c2 + a03 * c1) * r;
    out[6] = (-a30 * s5 + a32 * s2 - a33 * s1) * r;  out[7] = (a20 * s5 - a22 * s2 + a23 * s1) * r;
    out[8] = (a10 * c4 - a11 * c2 + a13 * c0) * r;   out[9] = (-a00 * c4 + a01 * c2 - a03 * c0) * r;
    out[10] = (a30 * s4 - a31 * s2 + a33 * s0) * r;  out[11] = (-a20 * s4 + a21 * s2 - a23 * s0) * r;
    out[12] = (-a10 * c3 + a11 * c1 - a12 * c0) * r; out[13] = (a00 * c3 - a01 * c1 + a02 * c0) * r;
    out[14] = (-a30 * s3 + a31 * s1 - a32 * s0) * r; out[15] = (a20 * s3 - a21 * s1 + a22 * s0) * r;
}

/* The block total as FilterMatchesByDenseVerifyCU_Kernel forms it (SIFTImageManager.cu:520-565), restated literally because it is NOT the plain sum
 * of the per-pixel terms: the block has (W, ceil(H / 32)) threads, thread (x, ty) adds its rows ty * 32 .. ty * 32 + 31 in order; warps are cut
 * from the linear thread id; `val += __shfl_down(val, offset)` -- a lane whose source lies past the warp's end adds ITSELF --; and the lanes that
 * contribute to the total are those with threadIdx.x % 32 == 0 (lane 0 of its warp only in the first thread row).  The reference adds those
 * contributions with shared-memory atomics (order = scheduling); here, and in the CUDA path, in ascending (row, x) order.
 * Needs W * ceil(H / 32) to be a multiple of 32 and at most 1024 (a partial warp in a full-mask shuffle is undefined on the GPU); returns 0 otherwise. */
static int dense_block_total(const float* pix /*[W*H][3]*/, unsigned W, unsigned H, float tot[3]) {
    const unsigned by = (H + 31) / 32, nt = W * by;
    if (nt % 32 != 0 || nt > 1024) return 0;
    static float local[1024][3];
    for (unsigned ty = 0; ty < by; ++ty)
        for (unsigned x = 0; x < W; ++x) {
            float acc[3] = { 0.0f, 0.0f, 0.0f };
            for (unsigned i = 0; i < 32; ++i) { const unsigned y = ty * 32 + i; if (y < H) for (int k = 0; k < 3; ++k) acc[k] += pix[3 * ((size_t)y * W + x) + k]; }
            for (int k = 0; k < 3; ++k) local[ty * W + x][k] = acc[k];
        }
    for (unsigned w0 = 0; w0 < nt; w0 += 32)
        for (int off = 16; off > 0; off /= 2) {
            float nxt[32][3];
            for (int l = 0; l < 32; ++l) { const int src = l + off < 32 ? l + off : l; for (int k = 0; k < 3; ++k) nxt[l][k] = local[w0 + l][k] + local[w0 + src][k]; }
            for (int l = 0; l < 32; ++l) for (int k = 0; k < 3; ++k) local[w0 + l][k] = nxt[l][k];
        }
    tot[0] = tot[1] = tot[2] = 0.0f;
    for (unsigned ty = 0; ty < by; ++ty) for (unsigned x = 0; x < W; x += 32) for (int k = 0; k < 3; ++k) tot[k] += local[ty * W + x][k];
    return 1;
}

/* FilterMatchesByDenseVerifyCU over pairs [startFrame, numFrames) \\ {curFrame}; frames: HOST array of cached-frame pointer records.
 * stats (optional): [numFrames][2] = (err, corr) for the pairs visited. */
ORC_API void orc_sift_filter_dense_verify(unsigned curFrame, unsigned startFrame, unsigned numFrames, unsigned W, unsigned H, const float* intrinsics,
                                          int32_t* numFiltered, const float* fT, const CachedFrame* frames, float distThresh, float normalThresh,
                                          float colorThresh, float errThresh, float corrThresh, float dMin, float dMax, float* stats) {
    (void)colorThresh;
    float* pix = (float*)malloc(sizeof(float) * 3 * (size_t)W * H);
    for (unsigned p = startFrame; p < numFrames; ++p) {
        if (p == curFrame) continue;
        if (numFiltered[p] == 0) continue;
        const float* T = fT + 16 * (size_t)p;
        float Tinv[16];
        mat4_inverse_subdet(T, Tinv);
        for (unsigned idx = 0; idx < W * H; ++idx) {
            float a[3], b[3];
            proj_error(idx, W, H, distThresh, normalThresh, T, intrinsics, &frames[p], &frames[curFrame], dMin, dMax, a);
            proj_error(idx, W, H, distThresh, normalThresh, Tinv, intrinsics, &frames[curFrame], &frames[p], dMin, dMax, b);
            for (int k = 0; k < 3; ++k) pix[3 * (size_t)idx + k] = a[k] + b[k];
        }
        float tot[3];
        if (!dense_block_total(pix, W, H, tot)) continue;
        const float err = tot[0] / tot[1], corr = 0.5f * tot[2] / (float)(W * H);
        if (stats) { stats[2 * p] = err; stats[2 * p + 1] = corr; }
        if (corr < corrThresh || err > errThresh || err != err) numFiltered[p] = 0;
    }
    free(pix);
}

/* SIFTImageManager::VerifyTrajectoryCU (FL/SiftGPU/SIFTImageManager.cu:1036-1150): every launched block b < N (N - 1) / 2 decodes
 * (img0, img1) = (b / N, b % N) -- so only the pairs whose row-major index is below N (N - 1) / 2 are looked at, kept --,
 * skips img0 >= img1 and invalid images, warps frame img0 into img1 with trajectory[img1]^-1 * trajectory[img0] and back, forms the
 * block total as the dense-verify filter does and returns 0 when any visited pair has corr < corrThresh, err > errThresh or err NaN.
 * trajectory: [N][16] row-major.  stats (optional): [N*N][2] = (err, corr) of the visited pairs at b, else untouched. */
ORC_API int orc_sift_verify_trajectory(unsigned numImages, const int32_t* validImages, const float* trajectory, unsigned W, unsigned H, const float* intrinsics,
                                       const CachedFrame* frames, float distThresh, float normalThresh, float colorThresh, float errThresh, float corrThresh,
                                       float dMin, float dMax, float* stats) {
    (void)colorThresh;
    if (numImages < 2) return 0;
    int valid = 1;
    float* pix = (float*)malloc(sizeof(float) * 3 * (size_t)W * H);
    const unsigned numPairs = numImages * (numImages - 1) / 2;
    for (unsigned b = 0; b < numPairs; ++b) {
        const unsigned img0 = b / numImages, img1 = b % numImages;
        if (img0 >= img1) continue;
        if (validImages[img0] == 0 || validImages[img1] == 0) continue;
        float inv1[16], T[16], Tinv[16];
        mat4_inverse_subdet(trajectory + 16 * (size_t)img1, inv1);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
            const float* A = inv1; const float* Bm = trajectory + 16 * (size_t)img0;
            T[4 * r + c] = A[4 * r] * Bm[c] + A[4 * r + 1] * Bm[4 + c] + A[4 * r + 2] * Bm[8 + c] + A[4 * r + 3] * Bm[12 + c];
        }
        mat4_inverse_subdet(T, Tinv);
        for (unsigned idx = 0; idx < W * H; ++idx) {
            float a[3], c[3];
            proj_error(idx, W, H, distThresh, normalThresh, T, intrinsics, &frames[img0], &frames[img1], dMin, dMax, a);
            proj_error(idx, W, H, distThresh, normalThresh, Tinv, intrinsics, &frames[img1], &frames[img0], dMin, dMax, c);
            for (int k = 0; k < 3; ++k) pix[3 * (size_t)idx + k] = a[k] + c[k];
        }
        float tot[3];
        if (!dense_block_total(pix, W, H, tot)) continue;
        const float err = tot[0] / tot[1], corr = 0.5f * tot[2] / (float)(W * H);
        if (stats) { stats[2 * b] = err; stats[2 * b + 1] = corr; }
        if (corr < corrThresh || err > errThresh || err != err) valid = 0;
    }
    free(pix);
    return valid;
}

/* ---- invalidation after a solve: InvalidateImageToImageCU_Kernel, CheckForInvalidFramesSimpleCU_Kernel / CheckForInvalidFramesCU_Kernel
 * (FL/SiftGPU/SIFTImageManager.cu:692-790).  entries: EntryJ records of 32 bytes.  The comprehensive variant is restated by its intent
 * (every still-valid correspondence touching an image with an empty table row), not by the reference's partial grid coverage. ---- */
ORC_API void orc_sift_invalidate_image_to_image(uint8_t* entries, unsigned numResiduals, unsigned imgI, unsigned imgJ) {
    for (unsigned k = 0; k < numResiduals; ++k) {
        uint32_t ij[2]; memcpy(ij, entries + 32 * (size_t)k, 8);
        if (ij[0] == imgI && ij[1] == imgJ) { ij[0] = ij[1] = 0xFFFFFFFFu; memcpy(entries + 32 * (size_t)k, ij, 8); }
    }
}
ORC_API void orc_sift_check_invalid_frames(const int32_t* numEntriesPerRow, int32_t* validImages, unsigned numVars, uint8_t* entries, unsigned numResiduals, int comprehensive) {
    if (comprehensive)
        for (unsigned k = 0; k < numResiduals; ++k) {
            uint32_t ij[2]; memcpy(ij, entries + 32 * (size_t)k, 8);
            if (ij[0] != 0xFFFFFFFFu && ((ij[0] < numVars && numEntriesPerRow[ij[0]] == 0) || (ij[1] < numVars && numEntriesPerRow[ij[1]] == 0))) {
                ij[0] = ij[1] = 0xFFFFFFFFu; memcpy(entries + 32 * (size_t)k, ij, 8);
            }
        }
    for (unsigned v = 0; v < numVars; ++v) if (numEntriesPerRow[v] == 0) validImages[v] = 0;
}

/* SIFTImageManager::filterFrames, FL/SiftGPU/SIFTImageManager.cpp:551-575: returns the last matched frame or -1, sets validImages[curFrame] */
ORC_API int orc_sift_filter_frames(unsigned curFrame, unsigned startFrame, unsigned numFrames, const int32_t* numFiltered, int32_t* validImages) {
    if (numFrames == 0) return -1;
    int connected = 0, last = -1;
    for (int i = (int)numFrames - 1; i >= (int)startFrame; --i)
        if (validImages[i] != 0 && numFiltered[i] > 0 && i != (int)curFrame) { connected = 1; last = i; break; }
    validImages[curFrame] = connected;
    return last;
}

/* small entry points for pinning the pieces against the reference's host-callable code (oracle/ref_kabsch_host.cpp) */
ORC_API int orc_eigen_system3(const float* m9, float* evs, float* ev9) {
    f3 ev[3];
    const int ok = eigen_system3(m9, evs, ev);
    if (ok) for (int i = 0; i < 3; ++i) { ev9[3 * i] = ev[i].x; ev9[3 * i + 1] = ev[i].y; ev9[3 * i + 2] = ev[i].z; }
    return ok;
}
ORC_API void orc_sym_eigenvalues3(const float* m9, float* evs) { sym_eigenvalues(m9, evs); }
ORC_API void orc_kabsch(const float* src, const float* tgt, unsigned n, float* T, float* evs) { kabsch((const f3*)src, (const f3*)tgt, n, T, evs); }
ORC_API unsigned orc_filter_pair(const float* keyPoints, uint32_t* idx, float* dist, unsigned numRaw, float* T, const float* Ki, unsigned minNum, float maxRes2) {
    return filter_pair((const KeyPoint*)keyPoints, idx, dist, numRaw, T, Ki, minNum, maxRes2);
}
ORC_API float orc_rsqrt_host_estimate(float x) { float out; _mm_store_ss(&out, _mm_rsqrt_ss(_mm_load_ss(&x))); return out; }

/* per-pixel contributions of the dense verification, for pinning against the reference's kernel, whose own reduction is NOT a plain sum (see
 * tests/test_manager_reference_emulated.py): out [W*H][3] = (residual, weight, count) of pixel idx, input->model plus model->input */
ORC_API void orc_sift_dense_verify_pixels(unsigned p, unsigned curFrame, unsigned W, unsigned H, const float* intrinsics, const float* fT, const CachedFrame* frames,
                                          float distThresh, float normalThresh, float dMin, float dMax, float* out) {
    const float* T = fT + 16 * (size_t)p;
    float Tinv[16];
    mat4_inverse_subdet(T, Tinv);
    for (unsigned idx = 0; idx < W * H; ++idx) {
        float a[3], b[3];
        proj_error(idx, W, H, distThresh, normalThresh, T, intrinsics, &frames[p], &frames[curFrame], dMin, dMax, a);
        proj_error(idx, W, H, distThresh, normalThresh, Tinv, intrinsics, &frames[curFrame], &frames[p], dMin, dMax, b);
        for (int k = 0; k < 3; ++k) out[3 * (size_t)idx + k] = a[k] + b[k];
    }
}
