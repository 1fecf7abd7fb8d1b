// TEST INFRASTRUCTURE ONLY -- stands in for <cuda_runtime.h> when the product's KERNEL SOURCE (faster_b200/csrc/fq_kernels_t.cuh) is
// compiled for the host under a lock-step warp emulation (tests/cpp/kernel_emu.cpp).  Not a CPU path of the product: nothing under
// faster_b200/ or include/ uses it, and the library still refuses to work without a GPU.  See simt_emu.h.
#pragma once
#include "simt_emu.h"
