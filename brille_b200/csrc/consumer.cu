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
// k_structure_factor: a CTA takes QB points at a time.  (1) every (point, atom) factor c_k e^{-W} e^{2 pi i Q.r_k} is evaluated
// once into shared memory; (2) one thread per (point, mode, atom) loads its complex 3-vector -- consecutive lanes own
// consecutive 48-byte pieces, fetched as one aligned 32-byte + one 16-byte streaming load, so a warp reads 1536 contiguous bytes
// with two instructions -- and leaves factor * (qv . eps) in shared memory; (3) one thread per (point, mode) sums the atoms and
// writes |F|^2 (coalesced).  HBM-read bound: 16*3*atoms*modes bytes per Q.
#include "consumer.cuh"

namespace b200 {

__device__ __forceinline__ void load48_stream(const double2* p, double2& a, double2& b, double2& c) {
  if ((reinterpret_cast<uintptr_t>(p) & 31u) == 0) {
    asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(c.x), "=d"(c.y) : "l"(p + 2));
  } else {
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(p));
    asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(b.x), "=d"(b.y), "=d"(c.x), "=d"(c.y) : "l"(p + 1));
  }
}

__global__ void __launch_bounds__(256) k_structure_factor(const double* __restrict__ Q, const double2* __restrict__ vecs, size_t n,
                                                          uint32_t M, uint32_t NAT, uint32_t QB, const __grid_constant__ SFDev c,
                                                          double* __restrict__ sf, const uint32_t* __restrict__ order,
                                                          const uint32_t* __restrict__ segment, uint32_t cap) {
  // list mode (order != NULL; the points of the fused path that the cell kernel does not take): row j of vecs belongs to the
  // point order[segment[1] + j], j < min(segment[2], cap) -- the compact rows the general interpolation kernel wrote
  if (order) n = segment[2] < cap ? segment[2] : cap;
  if (n == 0) return;
  const uint32_t* const list = order ? order + segment[1] : nullptr;
  extern __shared__ __align__(16) unsigned char sf_smem[];
  double2* const PH = reinterpret_cast<double2*>(sf_smem);  // [QB][NAT] per (point, atom) factor
  double2* const FP = PH + (size_t)QB * NAT;                // [QB][M][NAT] per (point, mode, atom) term
  double* const QV = reinterpret_cast<double*>(FP + (size_t)QB * M * NAT);  // [QB][3]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const size_t n_blocks = (n + QB - 1) / QB;
  const uint32_t MN = M * NAT;
  const uint32_t nat_magic = 0xffffffffu / NAT + 1u, mn_magic = 0xffffffffu / MN + 1u;  // exact for u * d < 2^32
  const uint32_t m_magic = 0xffffffffu / M + 1u;
  const double sgn = c.conjugate ? -1.0 : 1.0;
  for (size_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const size_t q0 = blk * QB;
    const uint32_t nq = (uint32_t)min((size_t)QB, n - q0);
    // ---- (1) per (point, atom) factor ------------------------------------------------------------------------------
    for (uint32_t u = tid; u < nq * NAT; u += nthr) {
      const uint32_t t = NAT == 1u ? u : __umulhi(u, nat_magic), k = u - t * NAT;
      const double* qr = Q + 3 * (list ? (size_t)list[q0 + t] : q0 + t);
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
    // ---- (2) per (point, mode, atom) term ------------------------------------------------------------------------------
    const double2* base = vecs + q0 * MN * 3;
    for (uint32_t u = tid; u < nq * MN; u += nthr) {
      const uint32_t t = __umulhi(u, mn_magic), k = u - (NAT == 1u ? u : __umulhi(u, nat_magic)) * NAT;
      double2 e0, e1, e2;
      load48_stream(base + 3 * (size_t)u, e0, e1, e2);
      const double v0 = QV[3 * t], v1 = QV[3 * t + 1], v2 = QV[3 * t + 2];
      const double dr = v0 * e0.x + v1 * e1.x + v2 * e2.x;
      const double di = sgn * (v0 * e0.y + v1 * e1.y + v2 * e2.y);
      const double2 f = PH[t * NAT + k];
      FP[u] = make_double2(f.x * dr - f.y * di, f.x * di + f.y * dr);
    }
    __syncthreads();
    // ---- (3) per (point, mode): sum over the atoms ----------------------------------------------------------------------
    for (uint32_t p = tid; p < nq * M; p += nthr) {
      const double2* fp = FP + (size_t)p * NAT;
      double Fr = 0.0, Fi = 0.0;
      for (uint32_t k = 0; k < NAT; ++k) {
        Fr += fp[k].x;
        Fi += fp[k].y;
      }
      const double v = Fr * Fr + Fi * Fi;
      if (list) {
        const uint32_t t = M == 1u ? p : __umulhi(p, m_magic);
        sf[(size_t)list[q0 + t] * M + (p - t * M)] = v;
      } else {
        sf[q0 * M + p] = v;
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_structure_factor(const SFDev& c, const double* dQ, const double* dvecs, size_t n, uint32_t M, double* dsf, int sm_count,
                                    cudaStream_t stream, const uint32_t* order, const uint32_t* segment, uint32_t cap) {
  if (n == 0) return cudaSuccess;
  // points per CTA round: about 2048 (point, mode, atom) terms (32 KB of shared memory), at least one point
  const uint32_t MN = M * c.n_atoms;
  uint32_t QB = 2048u / (MN ? MN : 1u);
  QB = QB < 1u ? 1u : (QB > 64u ? 64u : QB);
  const size_t smem = (size_t)QB * c.n_atoms * 16 + (size_t)QB * MN * 16 + (size_t)QB * 24;
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (configured < 48 * 1024) configured = 48 * 1024;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_structure_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const size_t blocks = (n + QB - 1) / QB, grid_cap = (size_t)sm_count * (order ? 2 : 6);  // list mode: normally (almost) empty
  k_structure_factor<<<(unsigned)(blocks < grid_cap ? blocks : grid_cap), 256, smem, stream>>>(dQ, reinterpret_cast<const double2*>(dvecs), n, M,
                                                                                               c.n_atoms, QB, c, dsf, order, segment, cap);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Powder average ("powder binning of |F(Q)|^2 per mode"; the loop validation/profiling.md:30-67 times):
// the sphere of directions is sampled at every |Q| and the one-phonon intensity |F(Q, nu)|^2 is binned on (|Q|, omega_nu(Q)).
// With the points generated and the histogram accumulated on the device NOTHING per Q crosses PCIe: a sweep returns
// n_qbins x n_wbins doubles.
//
// k_powder_q     point g of the sweep = (|Q| bin i, direction j): |Q| at the centre of bin i, direction from a counter-based
//                generator (splitmix64 of seed and g: reproducible on the host, independent of the launch shape), Q = B^-1 (|Q| d)
// k_powder_bin   one thread per (point, mode): |B Q| -> |Q| bin, eigenvalue -> energy bin, atomicAdd of the weighted intensity
//                (FP64 reductions at L2; the histogram is L2 resident), one count per point for the normalisation
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_powder_q(double* __restrict__ Q, size_t first, size_t n, uint64_t n_dir_local, uint64_t dir_lo,
                                                  uint64_t n_dir, uint64_t seed, PowderDev c) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t g = first + t;
    const uint64_t i = g / n_dir_local, j = dir_lo + (g - i * n_dir_local);
    const uint64_t z1 = splitmix64(seed ^ splitmix64(i * n_dir + j)), z2 = splitmix64(z1);
    const double u1 = (double)(z1 >> 11) * 0x1.0p-53, u2 = (double)(z2 >> 11) * 0x1.0p-53;
    const double ct = 1.0 - 2.0 * u1, st = sqrt(fmax(0.0, 1.0 - ct * ct));
    double sp, cp;
    sincospi(2.0 * u2, &sp, &cp);
    const double qn = c.q_lo + ((double)i + 0.5) * c.dq;
    const double x = qn * st * cp, y = qn * st * sp, zc = qn * ct;
    Q[3 * t] = c.Binv[0] * x + c.Binv[1] * y + c.Binv[2] * zc;
    Q[3 * t + 1] = c.Binv[3] * x + c.Binv[4] * y + c.Binv[5] * zc;
    Q[3 * t + 2] = c.Binv[6] * x + c.Binv[7] * y + c.Binv[8] * zc;
  }
}

// A CTA takes 256 consecutive points.  In a sweep they sit in one |Q| bin (the points are |Q|-bin major): the row of that bin is
// accumulated in shared memory and flushed with one global reduction per non-empty energy bin; points of other |Q| bins (any
// order of caller-provided points works) go to the global histogram directly.  The per-bin point counts are aggregated per warp.
__global__ void __launch_bounds__(256) k_powder_bin(const double* __restrict__ Q, const double* __restrict__ vals, const double* __restrict__ sf,
                                                    size_t n, uint32_t M, uint32_t vspan, PowderDev c, double* __restrict__ hist,
                                                    double* __restrict__ counts, uint32_t row_bins) {
  extern __shared__ __align__(16) unsigned char pb_smem[];
  double* const row = reinterpret_cast<double*>(pb_smem);  // [row_bins] (0: no private row)
  __shared__ int s_iq[256];
  __shared__ int s_row;
  const int tid = threadIdx.x;
  for (size_t p0 = (size_t)blockIdx.x * 256; p0 < n; p0 += (size_t)gridDim.x * 256) {
    const uint32_t np = (uint32_t)min((size_t)256, n - p0);
    // ---- per point: |B Q| -> |Q| bin; counts, one reduction per distinct bin of a warp ------------------------------------
    int iq = -1;
    if ((uint32_t)tid < np) {
      const double* q = Q + 3 * (p0 + tid);
      const double q0 = q[0], q1 = q[1], q2 = q[2];
      const double x = c.B[0] * q0 + c.B[1] * q1 + c.B[2] * q2, y = c.B[3] * q0 + c.B[4] * q1 + c.B[5] * q2, z = c.B[6] * q0 + c.B[7] * q1 + c.B[8] * q2;
      const double fq = (sqrt(x * x + y * y + z * z) - c.q_lo) * c.inv_dq;
      if (fq >= 0.0 && fq < (double)c.n_qbins) iq = (int)fq;
    }
    s_iq[tid] = iq;
    {
      const unsigned same = __match_any_sync(0xffffffffu, iq);
      if (iq >= 0 && (int)(tid & 31) == __ffs(same) - 1) atomicAdd(counts + iq, (double)__popc(same));
    }
    if (tid == 0) s_row = iq;  // the block's private row: the |Q| bin of its first point
    for (uint32_t k = tid; k < row_bins; k += 256) row[k] = 0.0;
    __syncthreads();
    const int my_row = row_bins ? s_row : -1;
    // ---- per (point, mode): energy bin, weighted intensity ---------------------------------------------------------------
    for (uint32_t g = tid; g < np * M; g += 256) {
      const uint32_t t = g / M;
      const int jq = s_iq[t];
      if (jq < 0) continue;
      const size_t pm = (p0 + t) * M + (g - t * M);
      const double w = vals[pm * vspan];
      const double fw = (w - c.w_lo) * c.inv_dw;
      if (!(fw >= 0.0) || fw >= (double)c.n_wbins) continue;
      double v = sf[pm];
      if (c.weight == 1) {
        if (!(w > 0.0)) continue;
        v /= w;
      }
      if (jq == my_row) atomicAdd(row + (uint32_t)fw, v);
      else atomicAdd(hist + (size_t)jq * c.n_wbins + (uint32_t)fw, v);
    }
    __syncthreads();
    if (my_row >= 0)
      for (uint32_t k = tid; k < row_bins; k += 256) {
        const double v = row[k];
        if (v != 0.0) atomicAdd(hist + (size_t)my_row * c.n_wbins + k, v);
      }
    __syncthreads();
  }
}

cudaError_t launch_powder_q(double* dQ, size_t first, size_t n, uint64_t n_dir_local, uint64_t dir_lo, uint64_t n_dir, uint64_t seed,
                            const PowderDev& c, int sm_count, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const size_t want = (n + 255) / 256, cap = (size_t)sm_count * 16;
  k_powder_q<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(dQ, first, n, n_dir_local, dir_lo, n_dir, seed, c);
  return cudaGetLastError();
}
cudaError_t launch_powder_bin(const double* dQ, const double* dvals, const double* dsf, size_t n, uint32_t M, uint32_t vspan, const PowderDev& c,
                              double* hist, double* counts, int sm_count, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const size_t want = (n + 255) / 256, cap = (size_t)sm_count * 8;
  const uint32_t row_bins = c.n_wbins <= 4096u ? c.n_wbins : 0u;  // (a private row of up to 32 KB)
  k_powder_bin<<<(unsigned)(want < cap ? want : cap), 256, (size_t)row_bins * sizeof(double), stream>>>(dQ, dvals, dsf, n, M, vspan, c, hist, counts,
                                                                                                          row_bins);
  return cudaGetLastError();
}

}  // namespace b200
