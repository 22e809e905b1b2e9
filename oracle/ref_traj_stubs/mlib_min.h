// mlib_min.h -- TEST INFRASTRUCTURE ONLY.  The handful of mLib types FL/TrajectoryManager.{h,cpp} and the Lie part of FL/PoseHelper.h use,
// written from the published semantics of niessner/mLib (core-math: point3d / point6d / matrix3x3 / matrix4x4; mLib itself is an un-vendored
// submodule and does not compile under g++): row-major matrices, `|` = dot product (products summed left to right), `^` = cross product,
// length = sqrt of the dot.  Only what those two files need; no reference code.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>
#include <list>
#include <mutex>
#include <vector>

namespace ml {
namespace math { static const float PIf = 3.14159265358979323846f; }
struct vec3f {
    float array[3];
    vec3f() { array[0] = array[1] = array[2] = 0.0f; }
    vec3f(float x, float y, float z) { array[0] = x; array[1] = y; array[2] = z; }
    float& operator[](unsigned i) { return array[i]; }
    const float& operator[](unsigned i) const { return array[i]; }
    float lengthSq() const { return array[0] * array[0] + array[1] * array[1] + array[2] * array[2]; }
    float length() const { return std::sqrt(lengthSq()); }
    vec3f& operator*=(float s) { array[0] *= s; array[1] *= s; array[2] *= s; return *this; }
    vec3f& operator-=(const vec3f& o) { array[0] -= o.array[0]; array[1] -= o.array[1]; array[2] -= o.array[2]; return *this; }
    vec3f operator*(float s) const { return vec3f(array[0] * s, array[1] * s, array[2] * s); }
    float operator|(const vec3f& o) const { return array[0] * o.array[0] + array[1] * o.array[1] + array[2] * o.array[2]; }
    vec3f operator^(const vec3f& o) const { return vec3f(array[1] * o.array[2] - array[2] * o.array[1], array[2] * o.array[0] - array[0] * o.array[2], array[0] * o.array[1] - array[1] * o.array[0]); }
};
struct vec6f {
    float array[6];
    vec6f() { for (float& v : array) v = 0.0f; }
    float& operator[](unsigned i) { return array[i]; }
    const float& operator[](unsigned i) const { return array[i]; }
    vec6f operator-(const vec6f& o) const { vec6f r; for (int i = 0; i < 6; ++i) r.array[i] = array[i] - o.array[i]; return r; }
    float operator|(const vec6f& o) const { float s = array[0] * o.array[0]; for (int i = 1; i < 6; ++i) s += array[i] * o.array[i]; return s; }
    vec3f getVec3() const { return vec3f(array[0], array[1], array[2]); }
};
struct mat3f {
    float m[3][3];
    mat3f() { std::memset(m, 0, sizeof m); }
    float& operator()(unsigned r, unsigned c) { return m[r][c]; }
    const float& operator()(unsigned r, unsigned c) const { return m[r][c]; }
    float trace() const { return m[0][0] + m[1][1] + m[2][2]; }
    vec3f operator*(const vec3f& v) const { return vec3f(m[0][0] * v[0] + m[0][1] * v[1] + m[0][2] * v[2], m[1][0] * v[0] + m[1][1] * v[1] + m[1][2] * v[2], m[2][0] * v[0] + m[2][1] * v[1] + m[2][2] * v[2]); }
};
struct mat4f {
    float matrix[16];
    mat4f() { std::memset(matrix, 0, sizeof matrix); }
    static mat4f zero(float v = 0.0f) { mat4f r; for (float& e : r.matrix) e = v; return r; }
    static mat4f identity() { mat4f r; r.matrix[0] = r.matrix[5] = r.matrix[10] = r.matrix[15] = 1.0f; return r; }
    float& operator[](unsigned i) { return matrix[i]; }
    const float& operator[](unsigned i) const { return matrix[i]; }
    float& operator()(unsigned r, unsigned c) { return matrix[4 * r + c]; }
    const float& operator()(unsigned r, unsigned c) const { return matrix[4 * r + c]; }
    mat3f getRotation() const { mat3f r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = matrix[4 * i + j]; return r; }
    vec3f getTranslation() const { return vec3f(matrix[3], matrix[7], matrix[11]); }
    void setRotationMatrix(const mat3f& r) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) matrix[4 * i + j] = r(i, j); }
    void setTranslationVector(const vec3f& t) { matrix[3] = t[0]; matrix[7] = t[1]; matrix[11] = t[2]; }
};
}  // namespace ml
using namespace ml;

struct float4x4 { float entries[16]; };
enum { cudaMemcpyDeviceToHost = 2 };
static inline int cudaMemcpy(void* d, const void* s, size_t n, int) { std::memcpy(d, s, n); return 0; }
#define MLIB_CUDA_SAFE_CALL(x) (x)
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
