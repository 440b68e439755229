// thermal_kernels.cuh -- placeholder state (filled in by the thermal milestone)
#pragma once
#include <cuda_runtime.h>
namespace mpp {
struct ThermalState { cudaStream_t stream = nullptr; double cnfac = 0.5; };
}
