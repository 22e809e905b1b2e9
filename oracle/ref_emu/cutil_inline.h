// TEST INFRASTRUCTURE ONLY: stands in for the SDK's <cutil_inline.h> in the CPU emulation of the reference (see ref_emu_cuda.h).
#pragma once
#include "ref_emu_cuda.h"
#define cutilSafeCall(x) (x)
#define cutilCheckMsg(x)
