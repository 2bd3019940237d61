// consumer.cu -- device-resident consumer of the interpolated eigenvectors.
//
// brille's callers (Euphonic's BrilleInterpolator, brilleu's s_qw; brille's own validation/profiling.md:30-67 times
// exactly that loop) reduce the (nQ, modes, atoms, 3) complex eigenvectors that ir_interpolate_at returns to one number
// per (Q, mode) straight away: the one-phonon structure factor
//
//     F(Q, nu) = sum_k  c_k  exp(-qv^T W_k qv)  exp(2 pi i Q.r_k)  ( qv . eps*_{nu,k}(Q) ),     S(Q, nu) = |F(Q, nu)|^2
//
// (Euphonic 1.x, QpointPhononModes.calculate_structure_factor: c_k = b_k / sqrt(m_k), qv = Cartesian Q, W_k the
// Debye-Waller matrix of atom k; the conjugate is Euphonic's convention and optional here).  The commented-out
// `ir_interpolate_at_dw` of the reference (wrap/_common_grid.hpp:343-405) fuses the same kind of per-atom Debye-Waller
// reduction behind the interpolation.  With the reduction on the device the 16*modes*3*atoms bytes per Q of eigenvectors never
// cross PCIe: 8*modes bytes per Q do.
//
// k_structure_factor: a CTA takes QB points at a time; every (point, atom) factor c_k e^{-W} e^{2 pi i Q.r_k} is evaluated
// once into shared memory, then one thread per (point, mode) streams its row of 3*atoms complex numbers (read once:
// ld.global.cs) and writes |F|^2.  HBM-read bound: 16*3*atoms*modes bytes per Q.
#include "consumer.cuh"

namespace b200 {

constexpr int SF_QB = 64;  // points per CTA round

__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 v;
  asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(256) k_structure_factor(const double* __restrict__ Q, const double2* __restrict__ vecs, size_t n,
                                                          uint32_t M, uint32_t NAT, const __grid_constant__ SFDev c,
                                                          double* __restrict__ sf) {
  extern __shared__ __align__(16) unsigned char sf_smem[];
  double2* const PH = reinterpret_cast<double2*>(sf_smem);              // [SF_QB][NAT]
  double* const QV = reinterpret_cast<double*>(PH + (size_t)SF_QB * NAT);  // [SF_QB][3]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const size_t n_blocks = (n + SF_QB - 1) / SF_QB;
  const size_t S = 3 * (size_t)NAT;
  for (size_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const size_t q0 = blk * SF_QB;
    const uint32_t nq = (uint32_t)min((size_t)SF_QB, n - q0);
    // ---- per (point, atom) factor -------------------------------------------------------------------------------
    for (uint32_t u = tid; u < nq * NAT; u += nthr) {
      const uint32_t t = u / NAT, k = u - t * NAT;
      const double* qr = Q + 3 * (q0 + t);
      const double q[3] = {qr[0], qr[1], qr[2]};
      double v[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] = c.T[3 * i] * q[0] + c.T[3 * i + 1] * q[1] + c.T[3 * i + 2] * q[2];
      if (k == 0) {
        QV[3 * t] = v[0];
        QV[3 * t + 1] = v[1];
        QV[3 * t + 2] = v[2];
      }
      double2 f = make_double2(c.coef[2 * k], c.coef[2 * k + 1]);
      if (c.pos) {
        const double* r = c.pos + 3 * k;
        const double dot = q[0] * r[0] + q[1] * r[1] + q[2] * r[2];
        double sn, cs;
        sincos(6.283185307179586476925286766559 * dot, &sn, &cs);
        f = make_double2(f.x * cs - f.y * sn, f.x * sn + f.y * cs);
      }
      if (c.dw) {
        const double* W = c.dw + 9 * k;
        double w = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) w += v[i] * (W[3 * i] * v[0] + W[3 * i + 1] * v[1] + W[3 * i + 2] * v[2]);
        const double e = exp(-w);
        f.x *= e;
        f.y *= e;
      }
      PH[u] = f;
    }
    __syncthreads();
    // ---- per (point, mode) reduction over the atoms ----------------------------------------------------------------
    const double sgn = c.conjugate ? -1.0 : 1.0;
    for (uint32_t p = tid; p < nq * M; p += nthr) {
      const uint32_t t = p / M;
      const double2* row = vecs + (q0 * M + p) * S;
      const double v0 = QV[3 * t], v1 = QV[3 * t + 1], v2 = QV[3 * t + 2];
      const double2* ph = PH + (size_t)t * NAT;
      double Fr = 0.0, Fi = 0.0;
      for (uint32_t k = 0; k < NAT; ++k) {
        const double2 e0 = ld_stream(row + 3 * k), e1 = ld_stream(row + 3 * k + 1), e2 = ld_stream(row + 3 * k + 2);
        const double dr = v0 * e0.x + v1 * e1.x + v2 * e2.x;
        const double di = sgn * (v0 * e0.y + v1 * e1.y + v2 * e2.y);
        const double2 f = ph[k];
        Fr += f.x * dr - f.y * di;
        Fi += f.x * di + f.y * dr;
      }
      sf[q0 * M + p] = Fr * Fr + Fi * Fi;
    }
    __syncthreads();
  }
}

cudaError_t launch_structure_factor(const SFDev& c, const double* dQ, const double* dvecs, size_t n, uint32_t M, double* dsf, int sm_count,
                                    cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const size_t smem = (size_t)SF_QB * c.n_atoms * 16 + (size_t)SF_QB * 24;
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_structure_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const size_t blocks = (n + SF_QB - 1) / SF_QB, cap = (size_t)sm_count * 8;
  k_structure_factor<<<(unsigned)(blocks < cap ? blocks : cap), 256, smem, stream>>>(dQ, reinterpret_cast<const double2*>(dvecs), n, M,
                                                                                     c.n_atoms, c, dsf);
  return cudaGetLastError();
}

}  // namespace b200
