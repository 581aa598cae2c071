// common.cuh -- device-side helpers shared by the sm_100a GPUAR kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "coder_math.h"

namespace gpuar {

constexpr uint32_t kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// launch bookkeeping (host; api.cu)
void count_launch(int n = 1);

}  // namespace gpuar
