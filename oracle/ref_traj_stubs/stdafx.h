// TEST INFRASTRUCTURE ONLY (oracle/build_ref.py: build_trajectory_host): what the reference's stdafx.h gives TrajectoryManager.cpp -- mLib's math types (the real ones,
// from /root/reference/external/mLib/include) -- plus the CUDA-side POD and the one runtime call the class makes, on host memory.
#pragma once
#include <cmath>
#include <cstring>
using std::isnan;
#include "mLibCore.h"
using namespace ml;
struct float4x4 { float entries[16]; };                        // cuda_SimpleMatrixUtil.h's layout; updateOptimizedTransform only copies it
enum { cudaMemcpyDeviceToHost = 2 };
static inline int cudaMemcpy(void* d, const void* s, size_t n, int) { std::memcpy(d, s, n); return 0; }      // "device" memory is host memory here
#define MLIB_CUDA_SAFE_CALL(x) (x)
