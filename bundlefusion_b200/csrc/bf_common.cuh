// bf_common.cuh -- shared device/host helpers for the bundlefusion_b200 CUDA library.
//
// Arithmetic contract of the TSDF path (see oracle/tsdf_oracle.c header): every float
// operation is individually rounded (this TU is compiled with -fmad=false, IEEE div/sqrt),
// float->int is cvt.rzi with saturation (CUDA's native behaviour), so results can be
// compared bit-for-bit with the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define BF_API extern "C" __attribute__((visibility("default")))

namespace bf {

// ---- per-thread last error + global stream --------------------------------------------
void set_last_error(const char* where, cudaError_t e);
cudaStream_t stream();

// returns the error code from the enclosing function (bf* extension API)
#define BF_CHECK(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) { ::bf::set_last_error(#expr, _e); return (int)_e; } \
    } while (0)

// cutilSafeCall behaviour for the reference-named stubs: print and exit(-1)
// (FriedLiver/Include/cutil/inc/cutil_inline_runtime.h:277-285, minus the getchar()).
#define BF_SAFE(expr)                                                         \
    do {                                                                      \
        int _rc = (expr);                                                     \
        if (_rc != 0) {                                                       \
            fprintf(stderr, "%s(%i) : bundlefusion_b200 runtime error %d: %s\n", __FILE__, __LINE__, _rc, \
                    cudaGetErrorString((cudaError_t)_rc));                    \
            exit(-1);                                                         \
        }                                                                     \
    } while (0)

static inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace bf
