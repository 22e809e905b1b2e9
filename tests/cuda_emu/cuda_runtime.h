// TEST INFRASTRUCTURE ONLY: stands in for <cuda_runtime.h> when a csrc/*.cu file is compiled for the CPU emulation (see cuda_emu.h).
#pragma once
#include "cuda_emu.h"
