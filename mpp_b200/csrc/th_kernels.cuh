// th_kernels.cuh -- placeholder state (filled in by the TH milestone)
#pragma once
#include <cuda_runtime.h>
#include "vsfm_kernels.cuh"
namespace mpp {
struct THState { cudaStream_t stream = nullptr; SnesOpts so; };
}
