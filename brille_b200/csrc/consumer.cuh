// consumer.cuh -- device-resident consumers of the interpolation output (consumer.cu)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

// one-phonon structure factor configuration (device copy of b200_sf_config_t)
struct SFDev {
  uint32_t n_atoms;
  const double* coef;  // (n_atoms,2) complex coefficient per atom
  const double* pos;   // (n_atoms,3) fractional positions, or null: no exp(2 pi i Q.r) factor
  const double* dw;    // (n_atoms,9) Debye-Waller matrices in the basis of qv, or null
  double T[9];         // qv = T Q (row-major)
  int conjugate;       // 1: qv . conj(eps)
};

cudaError_t launch_structure_factor(const SFDev& c, const double* dQ, const double* dvecs, size_t n, uint32_t M, double* dsf, int sm_count,
                                    cudaStream_t stream);

}  // namespace b200
