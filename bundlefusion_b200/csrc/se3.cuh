// se3.cuh -- small 3-vector helpers and the SE(3) exp / log maps of the reference's Lie parametrisation (FL/Solver/LieDerivUtil.h:19-207),
// shared by the solver (solver.cu) and the trajectory glue (trajectory.cu).
#pragma once
#include <cuda_runtime.h>

namespace bf {

struct V3 { float x, y, z; };
__host__ __device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r = { x, y, z }; return r; }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ V3 mulv(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__host__ __device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__host__ __device__ __forceinline__ V3 ld3(const float* p, unsigned i) { return mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__host__ __device__ __forceinline__ void st3(float* p, unsigned i, V3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }
__host__ __device__ __forceinline__ V3 xf(const float* m, V3 v) {
    return mk(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3], m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7], m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11]);
}

// ---- SE(3) exp / log (LieDerivUtil.h:19-207) ------------------------------------------------------------
__host__ __device__ __forceinline__ void rodrigues(V3 w, float A, float B, float* R /*9*/) {
    const float wx2 = w.x * w.x, wy2 = w.y * w.y, wz2 = w.z * w.z;
    R[0] = 1.0f - B * (wy2 + wz2); R[4] = 1.0f - B * (wx2 + wz2); R[8] = 1.0f - B * (wx2 + wy2);
    float a = A * w.z, b = B * (w.x * w.y); R[1] = b - a; R[3] = b + a;
    a = A * w.y; b = B * (w.x * w.z); R[2] = b + a; R[6] = b - a;
    a = A * w.x; b = B * (w.y * w.z); R[5] = b - a; R[7] = b + a;
}
static __host__ __device__ void exp_rotation(V3 w, float* R) {
    const float theta_sq = dot(w, w), theta = sqrtf(theta_sq);
    float A, B;
    if (theta_sq < 1e-8) { A = 1.0f - 0.16666667f * theta_sq; B = 0.5f; }
    else if (theta_sq < 1e-6) { B = 0.5f - 0.25f * 0.16666667f * theta_sq; A = 1.0f - theta_sq * 0.16666667f * (1.0f - 0.05f * theta_sq); }
    else { const float inv = 1.0f / theta; A = sinf(theta) * inv; B = (1 - cosf(theta)) * (inv * inv); }
    rodrigues(w, A, B, R);
}
static __host__ __device__ V3 ln_rotation(const float* M) {
#define Rm(r, c) M[(r) * 4 + (c)]
    const float cos_angle = (Rm(0, 0) + Rm(1, 1) + Rm(2, 2) - 1.0f) * 0.5f;
    V3 result = mk((Rm(2, 1) - Rm(1, 2)) * 0.5f, (Rm(0, 2) - Rm(2, 0)) * 0.5f, (Rm(1, 0) - Rm(0, 1)) * 0.5f);
    const float sin_angle_abs = length(result);
    if (cos_angle > 0.70710678118654752440f) {
        if (sin_angle_abs > 0) result = result * (asinf(sin_angle_abs) / sin_angle_abs);
    } else if (cos_angle > -0.70710678118654752440f) {
        const float angle = acosf(cos_angle);
        result = result * (angle / sin_angle_abs);
    } else {
        const float angle = 3.14159265358979323846f - asinf(sin_angle_abs);
        const float d0 = Rm(0, 0) - cos_angle, d1 = Rm(1, 1) - cos_angle, d2 = Rm(2, 2) - cos_angle;
        V3 r2;
        if (fabsf(d0) > fabsf(d1) && fabsf(d0) > fabsf(d2)) r2 = mk(d0, (Rm(1, 0) + Rm(0, 1)) * 0.5f, (Rm(0, 2) + Rm(2, 0)) * 0.5f);
        else if (fabsf(d1) > fabsf(d2)) r2 = mk((Rm(1, 0) + Rm(0, 1)) * 0.5f, d1, (Rm(2, 1) + Rm(1, 2)) * 0.5f);
        else r2 = mk((Rm(0, 2) + Rm(2, 0)) * 0.5f, (Rm(2, 1) + Rm(1, 2)) * 0.5f, d2);
        if (dot(r2, result) < 0) r2 = r2 * -1.0f;
        result = r2 * (angle / length(r2));
    }
#undef Rm
    return result;
}
static __host__ __device__ void pose_to_matrix(V3 rot, V3 trans, float* M /*16*/) {
    const float theta_sq = dot(rot, rot), theta = sqrtf(theta_sq);
    float A, B;
    V3 translation;
    const V3 cr = cross(rot, trans);
    if (theta_sq < 1e-8) {
        A = 1.0f - 0.16666667f * theta_sq; B = 0.5f;
        translation = trans + cr * 0.5f;
    } else {
        float C;
        if (theta_sq < 1e-6) { C = 0.16666667f * (1.0f - 0.05f * theta_sq); A = 1.0f - theta_sq * C; B = 0.5f - 0.25f * 0.16666667f * theta_sq; }
        else { const float inv = 1.0f / theta; A = sinf(theta) * inv; B = (1 - cosf(theta)) * (inv * inv); C = (1 - A) * (inv * inv); }
        const V3 wc = cross(rot, cr);
        translation = trans + cr * B + wc * C;
    }
    float R[9];
    rodrigues(rot, A, B, R);
    M[0] = R[0]; M[1] = R[1]; M[2] = R[2]; M[3] = translation.x;
    M[4] = R[3]; M[5] = R[4]; M[6] = R[5]; M[7] = translation.y;
    M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = translation.z;
    M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
}
static __host__ __device__ void matrix_to_pose(const float* M, V3& rot, V3& trans) {
    const V3 t = mk(M[3], M[7], M[11]);
    rot = ln_rotation(M);
    const float theta = length(rot);
    float shtot = 0.5f;
    if (theta > 0.00001f) shtot = sinf(theta * 0.5f) / theta;
    float H[9];
    exp_rotation(rot * -0.5f, H);
    trans = mk(H[0] * t.x + H[1] * t.y + H[2] * t.z, H[3] * t.x + H[4] * t.y + H[5] * t.z, H[6] * t.x + H[7] * t.y + H[8] * t.z);
    if (theta > 0.001f) trans = trans - rot * (dot(t, rot) * (1 - 2 * shtot) / dot(rot, rot));
    else trans = trans - rot * (dot(t, rot) / 24);
    trans = trans * (1.0f / (2 * shtot));
}

}  // namespace bf
