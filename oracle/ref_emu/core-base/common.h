// TEST INFRASTRUCTURE ONLY: the three macros of mLib's core-base/common.h that the reference's CUDA-side headers use.
#pragma once
#include <stdexcept>
#include <string>
typedef unsigned char uchar;
#define MLIB_EXCEPTION(s) std::runtime_error(std::string(s))
#define MLIB_ASSERT(x)
#define SAFE_DELETE_ARRAY(p) { if (p) { delete[] (p); (p) = NULL; } }
