// TEST INFRASTRUCTURE ONLY: stands in for <cuda_runtime_api.h> in the CPU emulation of the reference (see ref_emu_cuda.h).
#pragma once
#include "ref_emu_cuda.h"
