/*
 * solver_oracle.c -- CPU restatement of the reference's sparse bundle-adjustment solver (GN + Jacobi-PCG in
 * Lie space), float32, single thread.
 *
 * TEST INFRASTRUCTURE ONLY (see tsdf_oracle.c header for the rules).  "Parity unpinned": the reference has
 * no tests or golden vectors for this path; this file is pinned by the known-answer tests in
 * tests/test_solver_oracle.py (SE(3) exp/log identities, recovery of known poses from exact correspondences,
 * agreement with an independent float64 dense Gauss-Newton written in numpy).
 *
 * Restates, function by function (FL/ = FriedLiver/Source/):
 *   exp / log of SE(3)          FL/Solver/LieDerivUtil.h:19-207
 *   generators, Lie update      FL/Solver/LieDerivUtil.h:231-242, 301-307
 *   -J^T f + preconditioner     FL/Solver/SolverBundlingEquationsLie.h:63-148
 *   J p, J^T (J p)              FL/Solver/SolverBundlingEquationsLie.h:154-228
 *   PCG init / iteration        FL/Solver/SolverBundling.cu:756-794, 894-1108
 *   GN loop + early-outs        FL/Solver/SolverBundling.cu:1137-1220 (5e-7 on p.Ap, 0.005 on max|delta|)
 *   variable -> corr table      FL/Solver/SolverBundling.cu:1226-1248 (slots in ascending correspondence order)
 *   max residual                FL/Solver/SolverBundlingEquationsLie.h:27-40, SolverBundling.cu:511-550
 * Sums run in the order a sequential reading of those kernels gives (ascending variable / correspondence).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/bf_solver.h"

#define ORC_API __attribute__((visibility("default")))
#define FLOAT_EPSILON 0.000001f       /* FL/SolverUtil.h:9 */

typedef struct { float x, y, z; } v3;
static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length(v3 a) { return sqrtf(dot(a, a)); }
static inline v3 ld3(const float* p, unsigned i) { return V(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
static inline void st3(float* p, unsigned i, v3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }
/* affine transform of a point, implicit w = 1 (cuda_SimpleMatrixUtil.h:937-944) */
static inline v3 xf(const float* m, v3 v) {
    return V(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * 1.0f, m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * 1.0f,
             m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * 1.0f);
}

/* LieDerivUtil.h:19-47 */
static void rodrigues(v3 w, float A, float B, float R[9]) {
    const float wx2 = w.x * w.x, wy2 = w.y * w.y, wz2 = w.z * w.z;
    R[0] = 1.0f - B * (wy2 + wz2); R[4] = 1.0f - B * (wx2 + wz2); R[8] = 1.0f - B * (wx2 + wy2);
    float a = A * w.z, b = B * (w.x * w.y); R[1] = b - a; R[3] = b + a;
    a = A * w.y; b = B * (w.x * w.z); R[2] = b + a; R[6] = b - a;
    a = A * w.x; b = B * (w.y * w.z); R[5] = b - a; R[7] = b + a;
}
/* LieDerivUtil.h:50-76 */
static void exp_rotation(v3 w, float R[9]) {
    const float theta_sq = dot(w, w), theta = sqrtf(theta_sq);
    float A, B;
    if (theta_sq < 1e-8) { A = 1.0f - 0.16666667f * theta_sq; B = 0.5f; }
    else if (theta_sq < 1e-6) { B = 0.5f - 0.25f * 0.16666667f * theta_sq; A = 1.0f - theta_sq * 0.16666667f * (1.0f - 0.05f * theta_sq); }
    else { const float inv = 1.0f / theta; A = sinf(theta) * inv; B = (1 - cosf(theta)) * (inv * inv); }
    rodrigues(w, A, B, R);
}
/* LieDerivUtil.h:79-133 */
static v3 ln_rotation(const float* M /* 4x4 row-major, uses the 3x3 part */) {
#define Rm(r, c) M[(r) * 4 + (c)]
    const float cos_angle = (Rm(0, 0) + Rm(1, 1) + Rm(2, 2) - 1.0f) * 0.5f;
    v3 result = V((Rm(2, 1) - Rm(1, 2)) * 0.5f, (Rm(0, 2) - Rm(2, 0)) * 0.5f, (Rm(1, 0) - Rm(0, 1)) * 0.5f);
    const float sin_angle_abs = length(result);
    if (cos_angle > 0.70710678118654752440f) {
        if (sin_angle_abs > 0) result = mul(result, asinf(sin_angle_abs) / sin_angle_abs);
    } else if (cos_angle > -0.70710678118654752440f) {
        const float angle = acosf(cos_angle);
        result = mul(result, angle / sin_angle_abs);
    } else {
        const float angle = 3.14159265358979323846f - asinf(sin_angle_abs);
        const float d0 = Rm(0, 0) - cos_angle, d1 = Rm(1, 1) - cos_angle, d2 = Rm(2, 2) - cos_angle;
        v3 r2;
        if (fabsf(d0) > fabsf(d1) && fabsf(d0) > fabsf(d2)) r2 = V(d0, (Rm(1, 0) + Rm(0, 1)) * 0.5f, (Rm(0, 2) + Rm(2, 0)) * 0.5f);
        else if (fabsf(d1) > fabsf(d2)) r2 = V((Rm(1, 0) + Rm(0, 1)) * 0.5f, d1, (Rm(2, 1) + Rm(1, 2)) * 0.5f);
        else r2 = V((Rm(0, 2) + Rm(2, 0)) * 0.5f, (Rm(2, 1) + Rm(1, 2)) * 0.5f, d2);
        if (dot(r2, result) < 0) r2 = mul(r2, -1.0f);
        result = mul(r2, angle / length(r2));
    }
#undef Rm
    return result;
}
/* LieDerivUtil.h:160-207 */
ORC_API void orc_pose_to_matrix(const float rot3[3], const float trans3[3], float M[16]) {
    const v3 rot = V(rot3[0], rot3[1], rot3[2]), trans = V(trans3[0], trans3[1], trans3[2]);
    const float theta_sq = dot(rot, rot), theta = sqrtf(theta_sq);
    float A, B;
    v3 translation;
    const v3 cr = cross(rot, trans);
    if (theta_sq < 1e-8) {
        A = 1.0f - 0.16666667f * theta_sq; B = 0.5f;
        translation = add(trans, mul(cr, 0.5f));
    } else {
        float C;
        if (theta_sq < 1e-6) { C = 0.16666667f * (1.0f - 0.05f * theta_sq); A = 1.0f - theta_sq * C; B = 0.5f - 0.25f * 0.16666667f * theta_sq; }
        else { const float inv = 1.0f / theta; A = sinf(theta) * inv; B = (1 - cosf(theta)) * (inv * inv); C = (1 - A) * (inv * inv); }
        const v3 wc = cross(rot, cr);
        translation = add(add(trans, mul(cr, B)), mul(wc, C));
    }
    float R[9];
    rodrigues(rot, A, B, R);
    M[0] = R[0]; M[1] = R[1]; M[2] = R[2]; M[3] = translation.x;
    M[4] = R[3]; M[5] = R[4]; M[6] = R[5]; M[7] = translation.y;
    M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = translation.z;
    M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
}
/* LieDerivUtil.h:135-158 */
ORC_API void orc_matrix_to_pose(const float M[16], float rot3[3], float trans3[3]) {
    const v3 t = V(M[3], M[7], M[11]);
    v3 rot = ln_rotation(M);
    const float theta = length(rot);
    float shtot = 0.5f;
    if (theta > 0.00001f) shtot = sinf(theta * 0.5f) / theta;
    float H[9];
    exp_rotation(mul(rot, -0.5f), H);
    v3 trans = V(H[0] * t.x + H[1] * t.y + H[2] * t.z, H[3] * t.x + H[4] * t.y + H[5] * t.z, H[6] * t.x + H[7] * t.y + H[8] * t.z);
    if (theta > 0.001f) trans = sub(trans, mul(rot, dot(t, rot) * (1 - 2 * shtot) / dot(rot, rot)));
    else trans = sub(trans, mul(rot, dot(t, rot) / 24));
    trans = mul(trans, 1.0f / (2 * shtot));
    rot3[0] = rot.x; rot3[1] = rot.y; rot3[2] = rot.z;
    trans3[0] = trans.x; trans3[1] = trans.y; trans3[2] = trans.z;
}
static void mat4_mul(const float* a, const float* b, float* o) {
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c)
        o[r * 4 + c] = a[r * 4] * b[c] + a[r * 4 + 1] * b[4 + c] + a[r * 4 + 2] * b[8 + c] + a[r * 4 + 3] * b[12 + c];
}
/* general fp32 4x4 inverse (cuda_SimpleMatrixUtil.h:980-1100), adjugate / determinant */
ORC_API void orc_mat4_inverse(const float* m, float* out) {
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    const float r = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * r;
}

/* host-side view of the solver buffers (all HOST pointers here) */
typedef struct OrcSolver {
    unsigned N, C, maxCorrPerImage;
    BFEntryJ* corr;                 /* [C] (may be invalidated by the table build) */
    int* varToCorr;                 /* [N * maxCorrPerImage] */
    int* numEntriesPerRow;          /* [N] */
    float *xRot, *xTrans;           /* [N][3] unknowns */
    float *deltaRot, *deltaTrans, *rRot, *rTrans, *zRot, *zTrans, *pRot, *pTrans, *ApRot, *ApTrans, *precRot, *precTrans;
    float* Jp;                      /* [C][3] */
    float* T;                       /* [N][16] */
    float* Tinv;                    /* [N][16] */
    float* rDotzOld;                /* [N] */
} OrcSolver;

static inline int corr_valid(const BFEntryJ* c) { return c->imgIdx_i != 0xFFFFFFFFu; }

/* SolverBundling.cu:1226-1248 */
ORC_API void orc_solver_build_table(BFEntryJ* corr, unsigned C, unsigned maxCorrPerImage, int* varToCorr, int* numEntriesPerRow, unsigned N) {
    memset(numEntriesPerRow, 0, sizeof(int) * N);
    /* When only ONE of the two rows overflows, the reference leaves the other row's slot unwritten (it then holds whatever
     * the buffer held: -1 from the constructor's memset, i.e. an out-of-bounds read).  The oracle marks such slots -1 and
     * skips them -- the defined behaviour closest to the intent ("invalidate"). */
    memset(varToCorr, 0xff, sizeof(int) * (size_t)N * maxCorrPerImage);
    for (unsigned x = 0; x < C; ++x) {
        BFEntryJ* c = &corr[x];
        if (!corr_valid(c)) continue;
        int o0 = numEntriesPerRow[c->imgIdx_i]++;
        int o1 = numEntriesPerRow[c->imgIdx_j]++;
        if ((unsigned)o0 < maxCorrPerImage && (unsigned)o1 < maxCorrPerImage) {
            varToCorr[c->imgIdx_i * maxCorrPerImage + o0] = (int)x;
            varToCorr[c->imgIdx_j * maxCorrPerImage + o1] = (int)x;
        } else {
            c->imgIdx_i = 0xFFFFFFFFu; c->imgIdx_j = 0xFFFFFFFFu;
        }
    }
}

/* SolverBundlingEquationsLie.h:63-148 (sparse part) */
static void eval_minus_jtf(OrcSolver* s, unsigned v, float wSparse, v3* resRot, v3* resTrans) {
    v3 rRot = V(0, 0, 0), rTrans = V(0, 0, 0), pRot = V(0, 0, 0), pTrans = V(0, 0, 0);
    st3(s->deltaRot, v, V(0, 0, 0)); st3(s->deltaTrans, v, V(0, 0, 0));
    int n = s->numEntriesPerRow[v];
    if ((unsigned)n > s->maxCorrPerImage) n = (int)s->maxCorrPerImage;
    for (int k = 0; k < n; ++k) {
        const int ci0 = s->varToCorr[v * s->maxCorrPerImage + k];
        if (ci0 < 0) continue;
        const BFEntryJ* c = &s->corr[ci0];
        if (!corr_valid(c)) continue;
        const float* TI = &s->T[16 * c->imgIdx_i];
        const float* TJ = &s->T[16 * c->imgIdx_j];
        const v3 pi = V(c->pos_i[0], c->pos_i[1], c->pos_i[2]), pj = V(c->pos_j[0], c->pos_j[1], c->pos_j[2]);
        v3 wp; float sign = 1;
        if (v != c->imgIdx_i) { sign = -1; wp = xf(TJ, pj); } else wp = xf(TI, pi);
        const v3 da = V(0.0f, -wp.z, wp.y), db = V(wp.z, 0.0f, -wp.x), dc = V(-wp.y, wp.x, 0.0f);
        const v3 r = sub(xf(TI, pi), xf(TJ, pj));
        rRot = add(rRot, mul(V(dot(da, r), dot(db, r), dot(dc, r)), sign));
        rTrans = add(rTrans, mul(r, sign));
        pRot = add(pRot, V(dot(da, da), dot(db, db), dot(dc, dc)));
        pTrans = add(pTrans, V(1.0f, 1.0f, 1.0f));
    }
    *resRot = mul(rRot, -wSparse);
    *resTrans = mul(rTrans, -wSparse);
    st3(s->precRot, v, V(pRot.x > FLOAT_EPSILON ? 1.0f / pRot.x : 1.0f, pRot.y > FLOAT_EPSILON ? 1.0f / pRot.y : 1.0f, pRot.z > FLOAT_EPSILON ? 1.0f / pRot.z : 1.0f));
    st3(s->precTrans, v, V(pTrans.x > FLOAT_EPSILON ? 1.0f / pTrans.x : 1.0f, pTrans.y > FLOAT_EPSILON ? 1.0f / pTrans.y : 1.0f, pTrans.z > FLOAT_EPSILON ? 1.0f / pTrans.z : 1.0f));
}

/* SolverBundlingEquationsLie.h:195-228 */
static v3 apply_j(const OrcSolver* s, unsigned ci, float wSparse) {
    v3 b = V(0, 0, 0);
    const BFEntryJ* c = &s->corr[ci];
    if (!corr_valid(c)) return b;
    if (c->imgIdx_i > 0) {
        const v3 wp = xf(&s->T[16 * c->imgIdx_i], V(c->pos_i[0], c->pos_i[1], c->pos_i[2]));
        const v3 da = V(0.0f, -wp.z, wp.y), db = V(wp.z, 0.0f, -wp.x), dc = V(-wp.y, wp.x, 0.0f);
        const v3 pp = ld3(s->pRot, c->imgIdx_i);
        b = add(b, add(add(add(mul(da, pp.x), mul(db, pp.y)), mul(dc, pp.z)), ld3(s->pTrans, c->imgIdx_i)));
    }
    if (c->imgIdx_j > 0) {
        const v3 wp = xf(&s->T[16 * c->imgIdx_j], V(c->pos_j[0], c->pos_j[1], c->pos_j[2]));
        const v3 da = V(0.0f, -wp.z, wp.y), db = V(wp.z, 0.0f, -wp.x), dc = V(-wp.y, wp.x, 0.0f);
        const v3 pp = ld3(s->pRot, c->imgIdx_j);
        b = sub(b, add(add(add(mul(da, pp.x), mul(db, pp.y)), mul(dc, pp.z)), ld3(s->pTrans, c->imgIdx_j)));
    }
    return mul(b, wSparse);
}

/* SolverBundlingEquationsLie.h:154-193 */
static void apply_jt(const OrcSolver* s, unsigned v, v3* outRot, v3* outTrans) {
    v3 oR = V(0, 0, 0), oT = V(0, 0, 0);
    int n = s->numEntriesPerRow[v];
    if ((unsigned)n > s->maxCorrPerImage) n = (int)s->maxCorrPerImage;
    for (int k = 0; k < n; ++k) {
        const int ci = s->varToCorr[v * s->maxCorrPerImage + k];
        if (ci < 0) continue;
        const BFEntryJ* c = &s->corr[ci];
        if (!corr_valid(c)) continue;
        v3 wp; float sign = 1;
        if (v != c->imgIdx_i) { sign = -1; wp = xf(&s->T[16 * c->imgIdx_j], V(c->pos_j[0], c->pos_j[1], c->pos_j[2])); }
        else wp = xf(&s->T[16 * c->imgIdx_i], V(c->pos_i[0], c->pos_i[1], c->pos_i[2]));
        const v3 da = V(0.0f, -wp.z, wp.y), db = V(wp.z, 0.0f, -wp.x), dc = V(-wp.y, wp.x, 0.0f);
        const v3 jp = ld3(s->Jp, (unsigned)ci);
        oR = add(oR, mul(V(dot(da, jp), dot(db, jp), dot(dc, jp)), sign));
        oT = add(oT, mul(jp, sign));
    }
    *outRot = oR; *outTrans = oT;
}

/* Runs the whole solveBundlingStub for the SPARSE term (weightsDense* must be 0).
 * xRot / xTrans: in/out unknowns [N][3].  stats[0]=GN iterations, stats[1]=total PCG iterations.
 * Returns 0, or -1 on allocation failure. */
ORC_API int orc_solver_solve_sparse(BFEntryJ* corr, unsigned C, unsigned N, unsigned maxCorrPerImage,
                                    float* xRot, float* xTrans, unsigned nNonLinear, unsigned nLinear,
                                    const float* weightsSparse, int rebuildTable, int* varToCorr, int* numEntriesPerRow,
                                    unsigned stats[4]) {
    OrcSolver s;
    memset(&s, 0, sizeof s);
    s.N = N; s.C = C; s.maxCorrPerImage = maxCorrPerImage; s.corr = corr; s.varToCorr = varToCorr; s.numEntriesPerRow = numEntriesPerRow;
    s.xRot = xRot; s.xTrans = xTrans;
    float* buf = (float*)calloc((size_t)N * (3 * 12 + 32 + 1) + (size_t)C * 3 + 16, sizeof(float));
    if (!buf) return -1;
    float* p = buf;
    s.deltaRot = p; p += 3 * N; s.deltaTrans = p; p += 3 * N; s.rRot = p; p += 3 * N; s.rTrans = p; p += 3 * N;
    s.zRot = p; p += 3 * N; s.zTrans = p; p += 3 * N; s.pRot = p; p += 3 * N; s.pTrans = p; p += 3 * N;
    s.ApRot = p; p += 3 * N; s.ApTrans = p; p += 3 * N; s.precRot = p; p += 3 * N; s.precTrans = p; p += 3 * N;
    s.T = p; p += 16 * N; s.Tinv = p; p += 16 * N; s.rDotzOld = p; p += N; s.Jp = p;
    if (rebuildTable) orc_solver_build_table(corr, C, maxCorrPerImage, varToCorr, numEntriesPerRow, N);
    unsigned totalPcg = 0, gnRun = 0;
    for (unsigned nIter = 0; nIter < nNonLinear; ++nIter) {
        const float wS = weightsSparse[nIter];
        ++gnRun;
        for (unsigned k = 0; k < N; ++k) { orc_pose_to_matrix(&xRot[3 * k], &xTrans[3 * k], &s.T[16 * k]); orc_mat4_inverse(&s.T[16 * k], &s.Tinv[16 * k]); }
        /* Initialization (:756-794) */
        float scanAlpha0 = 0.0f, scanAlpha1 = 0.0f;
        for (unsigned x = 1; x < N; ++x) {
            v3 resRot, resTrans;
            eval_minus_jtf(&s, x, wS, &resRot, &resTrans);
            st3(s.rRot, x, resRot); st3(s.rTrans, x, resTrans);
            const v3 pR = mulv(ld3(s.precRot, x), resRot), pT = mulv(ld3(s.precTrans, x), resTrans);
            st3(s.pRot, x, pR); st3(s.pTrans, x, pT);
            scanAlpha0 += dot(resRot, pR) + dot(resTrans, pT);
            st3(s.ApRot, x, V(0, 0, 0)); st3(s.ApTrans, x, V(0, 0, 0));
        }
        for (unsigned x = 1; x < N; ++x) s.rDotzOld[x] = scanAlpha0;
        /* PCG (:1024-1108) */
        for (unsigned lin = 0; lin < nLinear; ++lin) {
            int last = (lin == nLinear - 1);
            ++totalPcg;
            scanAlpha0 = 0.0f; scanAlpha1 = 0.0f;
            for (unsigned c = 0; c < C; ++c) st3(s.Jp, c, apply_j(&s, c, wS));
            for (unsigned x = 1; x < N; ++x) {
                v3 r, t; apply_jt(&s, x, &r, &t);
                st3(s.ApRot, x, add(ld3(s.ApRot, x), r)); st3(s.ApTrans, x, add(ld3(s.ApTrans, x), t));
            }
            for (unsigned x = 1; x < N; ++x) scanAlpha0 += dot(ld3(s.pRot, x), ld3(s.ApRot, x)) + dot(ld3(s.pTrans, x), ld3(s.ApTrans, x));
            const float dotProduct = scanAlpha0;
            for (unsigned x = 1; x < N; ++x) {
                float alpha = 0.0f;
                if (dotProduct > FLOAT_EPSILON) alpha = s.rDotzOld[x] / dotProduct;
                st3(s.deltaRot, x, add(ld3(s.deltaRot, x), mul(ld3(s.pRot, x), alpha)));
                st3(s.deltaTrans, x, add(ld3(s.deltaTrans, x), mul(ld3(s.pTrans, x), alpha)));
                const v3 rR = sub(ld3(s.rRot, x), mul(ld3(s.ApRot, x), alpha)), rT = sub(ld3(s.rTrans, x), mul(ld3(s.ApTrans, x), alpha));
                st3(s.rRot, x, rR); st3(s.rTrans, x, rT);
                const v3 zR = mulv(ld3(s.precRot, x), rR), zT = mulv(ld3(s.precTrans, x), rT);
                st3(s.zRot, x, zR); st3(s.zTrans, x, zT);
                scanAlpha1 += dot(zR, rR) + dot(zT, rT);
            }
            if (getenv("ORC_DEBUG")) fprintf(stderr, "gn %u pcg %u pAp %.6e rz_old %.6e rz_new %.6e\n", nIter, lin, scanAlpha0, s.rDotzOld[1], scanAlpha1);
            if (fabsf(scanAlpha0) < 5e-7f) last = 1;
            for (unsigned x = 1; x < N; ++x) {
                const float rDotzNew = scanAlpha1, rDotzOld = s.rDotzOld[x];
                float beta = 0.0f;
                if (rDotzOld > FLOAT_EPSILON) beta = rDotzNew / rDotzOld;
                s.rDotzOld[x] = rDotzNew;
                st3(s.pRot, x, add(ld3(s.zRot, x), mul(ld3(s.pRot, x), beta)));
                st3(s.pTrans, x, add(ld3(s.zTrans, x), mul(ld3(s.pTrans, x), beta)));
                st3(s.ApRot, x, V(0, 0, 0)); st3(s.ApTrans, x, V(0, 0, 0));
                if (last) {   /* computeLieUpdate, LieDerivUtil.h:301-307 */
                    float U[16], Cm[16], P[16];
                    orc_pose_to_matrix(&s.deltaRot[3 * x], &s.deltaTrans[3 * x], U);
                    orc_pose_to_matrix(&xRot[3 * x], &xTrans[3 * x], Cm);
                    mat4_mul(U, Cm, P);
                    orc_matrix_to_pose(P, &xRot[3 * x], &xTrans[3 * x]);
                }
            }
            if (last) break;
        }
        /* EvalGNConvergence (:694-749): max |delta| over variables > 0 (all images treated as valid here) */
        if (nIter < nNonLinear - 1) {
            float m = 0.0f;
            for (unsigned x = 1; x < N; ++x) for (int k = 0; k < 3; ++k) {
                m = fmaxf(m, fabsf(s.deltaRot[3 * x + k])); m = fmaxf(m, fabsf(s.deltaTrans[3 * x + k]));
            }
            if (m < 0.005f) break;
        }
    }
    if (stats) { stats[0] = gnRun; stats[1] = totalPcg; }
    free(buf);
    return 0;
}

/* evalAbsMaxResidualDevice over all correspondences (SolverBundlingEquationsLie.h:27-40; SolverBundling.cu:511-550 +
 * CUDASolverBundling.cpp:313-329): returns the maximum and writes its correspondence index */
ORC_API float orc_solver_max_residual(const BFEntryJ* corr, unsigned C, const float* xRot, const float* xTrans, float wSparse, int* outIndex) {
    float best = 0.0f; int bi = 0;
    for (unsigned x = 0; x < C; ++x) {
        const BFEntryJ* c = &corr[x];
        if (!corr_valid(c)) continue;
        float TI[16], TJ[16];
        orc_pose_to_matrix(&xRot[3 * c->imgIdx_i], &xTrans[3 * c->imgIdx_i], TI);
        orc_pose_to_matrix(&xRot[3 * c->imgIdx_j], &xTrans[3 * c->imgIdx_j], TJ);
        const v3 d = sub(xf(TI, V(c->pos_i[0], c->pos_i[1], c->pos_i[2])), xf(TJ, V(c->pos_j[0], c->pos_j[1], c->pos_j[2])));
        const float r = fmaxf(wSparse * fabsf(d.z), fmaxf(wSparse * fabsf(d.x), wSparse * fabsf(d.y)));
        if (best < r) { best = r; bi = (int)x; }
    }
    if (outIndex) *outIndex = bi;
    return best;
}

/* sum of squared sparse residuals (evalFDevice, SolverBundlingEquationsLie.h:42-57) -- the GN energy */
ORC_API double orc_solver_energy(const BFEntryJ* corr, unsigned C, const float* xRot, const float* xTrans, float wSparse) {
    double e = 0.0;
    for (unsigned x = 0; x < C; ++x) {
        const BFEntryJ* c = &corr[x];
        if (!corr_valid(c)) continue;
        float TI[16], TJ[16];
        orc_pose_to_matrix(&xRot[3 * c->imgIdx_i], &xTrans[3 * c->imgIdx_i], TI);
        orc_pose_to_matrix(&xRot[3 * c->imgIdx_j], &xTrans[3 * c->imgIdx_j], TJ);
        const v3 d = sub(xf(TI, V(c->pos_i[0], c->pos_i[1], c->pos_i[2])), xf(TJ, V(c->pos_j[0], c->pos_j[1], c->pos_j[2])));
        e += (double)(wSparse * dot(d, d));
    }
    return e;
}
