// mat4.cuh -- 4x4 row-major float matrices on host and device with ONE operation order (translation units that include this are
// built -fmad=false, so host and device agree bit for bit).
#pragma once
#include <cuda_runtime.h>

namespace bf {

// The inverse as the reference's HOST forms it: float4x4::getInverse (FL/SiftGPU/cuda_SimpleMatrixUtil.h:980-1100) and mLib's mat4f::getInverse
// (external/mLib/include/core-math/matrix4x4.h:587-710) are the same expansion -- every adjugate entry a sum of six triple products, each product evaluated
// left to right, the determinant from the first row, one reciprocal -- compiled without contraction by the reference's host compiler.  The pose inverses the
// host hands to the kernels (setLastRigidTransform, FL/DepthSensing/CUDASceneRepHashSDF.h:128-134; the ray cast's view matrix, CUDARayCastSDF.cpp:92) are
// formed with this one, so that they equal the reference's bit for bit (tests/test_mat4_inverse_reference.py: against both reference implementations, g++).
// The order of the 96 products is the formula's (the one MESA's gluInvertMatrix made common); entry (r, c) of the adjugate below is row-major index 4r + c.
__host__ __device__ inline void mat4_inverse_ref(const float* m, float* out) {
    float a[16];
    a[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    a[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    a[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    a[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    a[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    a[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    a[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    a[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    a[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    a[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    a[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    a[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    a[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    a[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    a[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    a[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const float det = m[0] * a[0] + m[1] * a[4] + m[2] * a[8] + m[3] * a[12];
    const float r = 1.0f / det;
    for (int i = 0; i < 16; ++i) out[i] = a[i] * r;
}

// The inverse the DEVICE-side filters take (sift_filter.cu, sift_verify.cu, trajectory.cu): the adjugate through 2x2 sub-determinants -- 40 products instead of
// 96.  The reference's kernels call float4x4::getInverse there, which nvcc contracts into FMAs in an order of its own choosing; neither form reproduces that, the
// two agree to a few ulp (oracle/filter_oracle.c: mat4_inverse_subdet).
__host__ __device__ inline void mat4_inverse_hd(const float* m, float* out) {
    // 2x2 sub-determinants of the lower two rows (s*) and upper two rows (c*)
    const float a00 = m[0], a01 = m[1], a02 = m[2], a03 = m[3];
    const float a10 = m[4], a11 = m[5], a12 = m[6], a13 = m[7];
    const float a20 = m[8], a21 = m[9], a22 = m[10], a23 = m[11];
    const float a30 = m[12], a31 = m[13], a32 = m[14], a33 = m[15];
    const float s0 = a00 * a11 - a10 * a01, s1 = a00 * a12 - a10 * a02, s2 = a00 * a13 - a10 * a03;
    const float s3 = a01 * a12 - a11 * a02, s4 = a01 * a13 - a11 * a03, s5 = a02 * a13 - a12 * a03;
    const float c5 = a22 * a33 - a32 * a23, c4 = a21 * a33 - a31 * a23, c3 = a21 * a32 - a31 * a22;
    const float c2 = a20 * a33 - a30 * a23, c1 = a20 * a32 - a30 * a22, c0 = a20 * a31 - a30 * a21;
    const float det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    const float r = 1.0f / det;
    out[0] = (a11 * c5 - a12 * c4 + a13 * c3) * r;
    out[1] = (-a01 * c5 + a02 * c4 - a03 * c3) * r;
    out[2] = (a31 * s5 - a32 * s4 + a33 * s3) * r;
    out[3] = (-a21 * s5 + a22 * s4 - a23 * s3) * r;
    out[4] = (-a10 * c5 + a12 * c2 - a13 * c1) * r;
    out[5] = (a00 * c5 - a02 * c2 + a03 * c1) * r;
    out[6] = (-a30 * s5 + a32 * s2 - a33 * s1) * r;
    out[7] = (a20 * s5 - a22 * s2 + a23 * s1) * r;
    out[8] = (a10 * c4 - a11 * c2 + a13 * c0) * r;
    out[9] = (-a00 * c4 + a01 * c2 - a03 * c0) * r;
    out[10] = (a30 * s4 - a31 * s2 + a33 * s0) * r;
    out[11] = (-a20 * s4 + a21 * s2 - a23 * s0) * r;
    out[12] = (-a10 * c3 + a11 * c1 - a12 * c0) * r;
    out[13] = (a00 * c3 - a01 * c1 + a02 * c0) * r;
    out[14] = (-a30 * s3 + a31 * s1 - a32 * s0) * r;
    out[15] = (a20 * s3 - a21 * s1 + a22 * s0) * r;
}

// float4x4::operator* (FL/SiftGPU/cuda_SimpleMatrixUtil.h:1164-1187): r_ij = a_i1 b_1j + a_i2 b_2j + a_i3 b_3j + a_i4 b_4j, fused the way
// nvcc fuses that expression (the pattern read off the reference's SASS for the 3-product case, oracle/tsdf_oracle.c header)
__host__ __device__ inline void mat4_mul_hd(const float* a, const float* b, float* r) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r[4 * i + j] = fmaf(a[4 * i + 3], b[12 + j], fmaf(a[4 * i + 2], b[8 + j], fmaf(a[4 * i], b[j], a[4 * i + 1] * b[4 + j])));
}

}  // namespace bf
