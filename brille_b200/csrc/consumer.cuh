// consumer.cuh -- device-resident consumers of the interpolation output (consumer.cu)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "device_tables.cuh"

namespace b200 {

cudaError_t launch_structure_factor(const SFDev& c, const double* dQ, const double* dvecs, size_t n, uint32_t M, double* dsf, int sm_count,
                                    cudaStream_t stream, const uint32_t* order = nullptr, const uint32_t* segment = nullptr, uint32_t cap = 0);

}  // namespace b200
