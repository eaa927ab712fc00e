// wgrad_tc.h — tcgen05 weight-gradient path (wgrad_tc.cu) behind unpp_wgrad / unpp_wgrad_grid.
#pragma once
#include <cuda_runtime.h>
#include "../../include/unpp.h"

namespace unpp {

// 3x3, dense dZ, every source and the output 16 or 32 channels wide, accumulators fit TMEM.
bool wgrad_tc_eligible(const UnppWgradArgs* a);
int wgrad_tc_grid(const UnppWgradArgs* a);                       // CTAs = partials written
int wgrad_tc_launch(const UnppWgradArgs* a, cudaStream_t stream);  // same partial layout as the mma.sync kernel

}  // namespace unpp
