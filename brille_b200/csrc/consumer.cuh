// consumer.cuh -- device-resident consumers of the interpolation output (consumer.cu)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "device_tables.cuh"

namespace b200 {

cudaError_t launch_structure_factor(const SFDev& c, const double* dQ, const double* dvecs, size_t n, uint32_t M, double* dsf, int sm_count,
                                    cudaStream_t stream, const uint32_t* order = nullptr, const uint32_t* segment = nullptr, uint32_t cap = 0);

// powder average (consumer.cu): device copy of b200_powder_config_t with the derived constants
struct PowderDev {
  uint32_t n_qbins, n_wbins;
  double q_lo, dq, inv_dq, w_lo, inv_dw;
  int weight;
  double B[9], Binv[9];  // x = B q (Cartesian 1/angstrom) and its inverse
};
cudaError_t launch_powder_q(double* dQ, size_t first, size_t n, uint64_t n_dir_local, uint64_t dir_lo, uint64_t n_dir, uint64_t seed,
                            const PowderDev& c, int sm_count, cudaStream_t stream);
cudaError_t launch_powder_bin(const double* dQ, const double* dvals, const double* dsf, size_t n, uint32_t M, uint32_t vspan, const PowderDev& c,
                              double* hist, double* counts, int sm_count, cudaStream_t stream);

}  // namespace b200
