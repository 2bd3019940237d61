// interp.cu -- stage 2 of the path: permuted, phase-aligned linear interpolation fused with the rotation
// back to Q.
//   DualInterpolator::interpolate_at / get_permutations   interpolatordual.hpp:149-155,374-382
//   Interpolator::interpolate_at_mix                      interpolator_at.tpp:91-127
//   utils::antiphase                                      utilities.tpp:567-579
//   Interpolator::rotate_in_place                         interpolator.hpp:386-428
//   rip_gamma_complex / rip_real / rip_recip / rip_axial  interpolator_gamma.tpp:49-139, interpolator_real.tpp:18-61, ...
//
// General kernel: a group of LANES lanes owns one (Q, mode) unit.  Lanes stride over the scalar elements,
// the 3-vectors (one per atom for eigenvectors) and the 3x3 matrices of the mode; the Hermitian product
// needed for the phase alignment is reduced over the group with warp shuffles.  Interpolated 3-vectors /
// matrices are rotated in registers and written once -- there is no intermediate buffer and the output is
// never read back.
#include "device_tables.cuh"
#include "brille_b200.h"

namespace b200 {

struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }


// sum over the LANES lanes of one group; `mask` names exactly those lanes (groups of one warp may sit
// in different loop iterations, so a full-warp mask would be wrong)
template <int LANES>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// rotation description for one unit
struct Rot {
  int kind;           // -1 none, 0 real, 1 recip, 2 axial, 3/4 gamma
  const double* R;    // matrix applied to vectors (already transposed for recip, R^-1 for axial/gamma)
  const double* Rm;   // second matrix for the matrix sandwich
  double det;
  int ridx, invridx;
};

// x <- M x for a complex (or real, im=0) 3-vector held as 3 cplx
__device__ __forceinline__ void rot3(const double* M, const cplx* v, cplx* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o[i].re = (M[3 * i] * v[0].re + M[3 * i + 1] * v[1].re) + M[3 * i + 2] * v[2].re;
    o[i].im = (M[3 * i] * v[0].im + M[3 * i + 1] * v[1].im) + M[3 * i + 2] * v[2].im;
  }
}

template <int LANES, bool CPLX>
__device__ __forceinline__ void interp_unit(const InterpDev& id, const DataDev& dd, bool phase, int nv, const uint32_t* vtx,
                                            const double* wgt, const uint32_t* pb, uint32_t b, int lane, unsigned gmask,
                                            const Rot& rot, const double* q_ir, double* out_row) {
  constexpr int W = CPLX ? 2 : 1;
  const uint32_t S = id.span;
  const size_t row = (size_t)id.branches * S;
  const double* base[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) base[i] = id.data + ((size_t)(i < nv ? vtx[i] : vtx[0]) * row + (size_t)pb[i] * S) * W;
  // weights times phase factors (unused slots carry weight 0 and alias vertex 0)
  cplx f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = {i < nv ? wgt[i] : 0.0, 0.0};
  if (CPLX && phase) {
    const double* d0 = id.data + ((size_t)vtx[0] * row + (size_t)b * S) * W;  // pivot keeps its own branch b
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      if (i >= nv) break;
      double re = 0.0, im = 0.0;
      for (uint32_t s = lane; s < S; s += LANES) {
        double2 a = reinterpret_cast<const double2*>(d0)[s];
        double2 x = reinterpret_cast<const double2*>(base[i])[s];
        re += a.x * x.x + a.y * x.y;
        im += a.x * x.y - a.y * x.x;
      }
      re = group_sum<LANES>(re, gmask);
      im = group_sum<LANES>(im, gmask);
      double th = -atan2(im, re);
      double sn, cs;
      sincos(th, &sn, &cs);
      f[i] = {wgt[i] * cs, wgt[i] * sn};
    }
  }
  double* out = out_row + (size_t)b * S * W;
  // scalars
  for (uint32_t s = lane; s < id.no0; s += LANES) {
    cplx acc = {0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i >= nv) break;
      if (CPLX) {
        double2 x = reinterpret_cast<const double2*>(base[i])[s];
        cplx t = cmul(f[i], {x.x, x.y});
        acc.re += t.re;
        acc.im += t.im;
      } else {
        acc.re += f[i].re * base[i][s];
      }
    }
    if (CPLX) reinterpret_cast<double2*>(out)[s] = make_double2(acc.re, acc.im);
    else out[s] = acc.re;
  }
  // 3-vectors
  const uint32_t G = dd.n_ops;
  for (uint32_t k = lane; k < id.no1; k += LANES) {
    const uint32_t off = id.no0 + 3 * k;
    cplx acc[3] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i >= nv) break;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (CPLX) {
          double2 x = reinterpret_cast<const double2*>(base[i])[off + c];
          cplx t = cmul(f[i], {x.x, x.y});
          acc[c].re += t.re;
          acc[c].im += t.im;
        } else {
          acc[c].re += f[i].re * base[i][off + c];
        }
      }
    }
    cplx o[3] = {acc[0], acc[1], acc[2]};
    uint32_t dest = k;
    if (rot.kind >= 3) {
      rot3(rot.R, acc, o);  // R_invR v  (interpolator_gamma.tpp:104)
      dest = dd.gamma_F0[(size_t)k * G + rot.invridx];
      const double* gv = dd.gamma_vectors + 3 * (size_t)dd.gamma_vidx[(size_t)k * G + rot.invridx];
      double dot = ((0.0 + q_ir[0] * gv[0]) + q_ir[1] * gv[1]) + q_ir[2] * gv[2];
      double sn, cs;
      sincos(6.283185307179586476925286766559 * dot, &sn, &cs);
      cplx ph = {cs, sn};
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = cmul(ph, o[c]);
    } else if (rot.kind >= 0) {
      rot3(rot.R, acc, o);
      if (rot.kind == 2) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { o[c].re *= rot.det; o[c].im *= rot.det; }
      }
    }
    const uint32_t doff = id.no0 + 3 * dest;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (CPLX) reinterpret_cast<double2*>(out)[doff + c] = make_double2(o[c].re, o[c].im);
      else out[doff + c] = o[c].re;
    }
  }
  // 3x3 matrices
  if (id.no2) {
    const uint32_t Nmat = (uint32_t)(sqrt((double)id.no2)) / 3u;
    // Gamma: the reference rotates only the first Nmat x Nmat of its 9 Nmat^2 matrices (interpolator_gamma.tpp:116-132) and copies
    // the whole mode back from a work array it shares with the vector pass (:100-114,134): the elements behind the rotated matrices
    // come out as what that array holds there -- element e of the rotated VECTORS of the mode while e < 3 no1, zero beyond.
    if (rot.kind >= 3) __syncwarp(gmask);  // (the rotated vectors of this mode were written by the other lanes of the group)
    for (uint32_t mm = lane; mm < id.no2; mm += LANES) {
      const uint32_t off = id.no0 + 3 * id.no1 + 9 * mm;
      if (rot.kind >= 3 && mm >= Nmat * Nmat) {
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          const uint32_t e = 9 * mm + c;
          double2 v = make_double2(0.0, 0.0);
          if (e < 3 * id.no1) v = reinterpret_cast<const double2*>(out)[id.no0 + e];
          reinterpret_cast<double2*>(out)[off + c] = v;
        }
        continue;
      }
      cplx acc[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) acc[c] = {0.0, 0.0};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i >= nv) break;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          if (CPLX) {
            double2 x = reinterpret_cast<const double2*>(base[i])[off + c];
            cplx t = cmul(f[i], {x.x, x.y});
            acc[c].re += t.re;
            acc[c].im += t.im;
          } else {
            acc[c].re += f[i].re * base[i][off + c];
          }
        }
      }
      uint32_t dmm = mm;
      if (rot.kind >= 0) {
        // first T = M * Rm, then O = R * T  (interpolator_real.tpp:51-57; gamma: interpolator_gamma.tpp:116-134)
        cplx t[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            t[3 * i + j].re = (acc[3 * i].re * rot.Rm[j] + acc[3 * i + 1].re * rot.Rm[3 + j]) + acc[3 * i + 2].re * rot.Rm[6 + j];
            t[3 * i + j].im = (acc[3 * i].im * rot.Rm[j] + acc[3 * i + 1].im * rot.Rm[3 + j]) + acc[3 * i + 2].im * rot.Rm[6 + j];
          }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            acc[3 * i + j].re = (rot.R[3 * i] * t[j].re + rot.R[3 * i + 1] * t[3 + j].re) + rot.R[3 * i + 2] * t[6 + j].re;
            acc[3 * i + j].im = (rot.R[3 * i] * t[j].im + rot.R[3 * i + 1] * t[3 + j].im) + rot.R[3 * i + 2] * t[6 + j].im;
          }
        if (rot.kind >= 3) {
          const uint32_t nn = mm / Nmat, m2 = mm % Nmat;
          const double* g1 = dd.gamma_vectors + 3 * (size_t)dd.gamma_vidx[(size_t)nn * G + rot.ridx];
          const double* g2 = dd.gamma_vectors + 3 * (size_t)dd.gamma_vidx[(size_t)m2 * G + rot.invridx];
          double d1 = ((0.0 + q_ir[0] * g1[0]) + q_ir[1] * g1[1]) + q_ir[2] * g1[2];
          double d2 = ((0.0 + q_ir[0] * g2[0]) + q_ir[1] * g2[1]) + q_ir[2] * g2[2];
          double s1, c1, s2, c2;
          sincos(6.283185307179586476925286766559 * d1, &s1, &c1);
          sincos(6.283185307179586476925286766559 * d2, &s2, &c2);
          cplx pp = cmul({c1, s1}, {c2, s2});
#pragma unroll
          for (int c = 0; c < 9; ++c) acc[c] = cmul(pp, acc[c]);
          const uint32_t v = dd.gamma_F0[(size_t)nn * G + rot.ridx];
          const uint32_t kk = dd.gamma_F0[(size_t)m2 * G + rot.invridx];
          dmm = v * Nmat + kk;
        }
      }
      const uint32_t doff = id.no0 + 3 * id.no1 + 9 * dmm;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        if (CPLX) reinterpret_cast<double2*>(out)[doff + c] = make_double2(acc[c].re, acc[c].im);
        else out[doff + c] = acc[c].re;
      }
    }
  }
}

__device__ __forceinline__ Rot make_rot(const InterpDev& id, const DataDev& dd, bool ir, int ridx, int invridx) {
  Rot r;
  r.kind = ir ? id.rot_kind : -1;
  r.ridx = ridx;
  r.invridx = invridx;
  r.det = 1.0;
  r.R = r.Rm = nullptr;
  if (r.kind < 0) return r;
  if (r.kind <= 2 && dd.rot_is_identity[ridx]) {  // interpolator_real.tpp:41
    r.kind = -1;
    return r;
  }
  switch (r.kind) {
    case 0:  // v <- R v ; M <- R M R^-1
      r.R = dd.rot_int + 9 * ridx;
      r.Rm = dd.rot_int + 9 * invridx;
      break;
    case 1:  // v <- R^T v ; M <- R^T M (R^-1)^T : transposed copies live G*9 further on
      r.R = dd.rot_int + 9 * (dd.n_ops + ridx);
      r.Rm = dd.rot_int + 9 * (dd.n_ops + invridx);
      break;
    case 2:  // v <- det(R) R^-1 v ; M <- R^-1 M R
      r.R = dd.rot_int + 9 * invridx;
      r.Rm = dd.rot_int + 9 * ridx;
      r.det = dd.rot_det[ridx];
      break;
    case 3:  // gamma, lattice units: v <- R^-1 v ; M <- R^-1 M R
      r.R = dd.rot_int + 9 * invridx;
      r.Rm = dd.rot_int + 9 * ridx;
      break;
    default:  // gamma, cartesian
      r.R = dd.rot_cart + 9 * invridx;
      r.Rm = dd.rot_cart + 9 * ridx;
      break;
  }
  return r;
}

template <int LANES>
__global__ void __launch_bounds__(256)
k_interp(DataDev dd, LocateIn in, size_t n, int ir, double* __restrict__ vals_out, double* __restrict__ vecs_out,
         const uint32_t* __restrict__ order, const uint32_t* __restrict__ segment, uint32_t compact_cap, unsigned long long* overflow) {
  // list mode (order != NULL): only the points order[segment[1] .. segment[1]+segment[2]) are processed -- the last
  // bucket of the counting sort in cellinterp.cu, whose population is only known on the device
  // compact mode (compact_cap != 0, list mode only; fused structure factor): the vectors row of the j-th listed point goes to row
  // j of vecs_out, a scratch of compact_cap rows; points beyond the scratch are counted in *overflow and skipped
  if (order) n = segment[2];
  if (n == 0) return;
  if (compact_cap && n > compact_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *overflow = n - compact_cap;
    n = compact_cap;
  }
  const uint32_t B = dd.values.branches;
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = LANES >= 32 ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
  const size_t group = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const size_t n_groups = ((size_t)gridDim.x * blockDim.x) / LANES;
  const size_t units = n * (size_t)B;
  const size_t vrow = (size_t)B * dd.values.span * (dd.values.is_complex ? 2 : 1);
  const size_t wrow = (size_t)B * dd.vectors.span * (dd.vectors.is_complex ? 2 : 1);
  // all lanes of a warp must run the same number of iterations (shuffles inside)
  const size_t iters = (units + n_groups - 1) / n_groups;
  for (size_t it = 0; it < iters; ++it) {
    size_t u = group + it * n_groups;
    const bool live = u < units;
    if (!live) u = units - 1;
    const size_t q = order ? (size_t)order[segment[1] + u / B] : u / B;
    const uint32_t b = (uint32_t)(u % B);
    const uint32_t st = in.status[q];
    const bool failed = (st & (B200_ST_OUTSIDE_BZ | B200_ST_OUTSIDE_WEDGE | B200_ST_NOT_FOUND)) != 0;
    int nv = failed ? 0 : in.n_vert[q];
    uint32_t vtx[8];
    double wgt[8];
    uint32_t pb[8];
    {
      const uint4* vp = reinterpret_cast<const uint4*>(in.vertex + 8 * q);
      uint4 a = vp[0], c = vp[1];
      vtx[0] = a.x; vtx[1] = a.y; vtx[2] = a.z; vtx[3] = a.w; vtx[4] = c.x; vtx[5] = c.y; vtx[6] = c.z; vtx[7] = c.w;
      const double2* wp = reinterpret_cast<const double2*>(in.weight + REC_DOUBLES * q);
#pragma unroll
      for (int j = 0; j < 4; ++j) { double2 w2 = wp[j]; wgt[2 * j] = w2.x; wgt[2 * j + 1] = w2.y; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) pb[i] = b;
    if (dd.n_perm_rows > 1 && nv > 0) {
      // DualInterpolator::get_permutations (interpolatordual.hpp:374-382): table(pivot, vertex_i)
      const uint64_t slots = in.slots[q];
      const uint32_t cell = in.cell[q];
      const int tet = in.tet[q];
      const uint32_t s0 = (uint32_t)(slots & 0xff);
      for (int i = 0; i < nv; ++i) {
        const uint32_t si = (uint32_t)((slots >> (8 * i)) & 0xff);
        uint32_t rowi;
        if (tet >= 0) rowi = dd.tet_perm[(size_t)tet * 16 + s0 * 4 + si];
        else rowi = dd.cube_perm[(size_t)in.node_index[cell] * 64 + s0 * 8 + si];
        pb[i] = dd.perm_rows[(size_t)rowi * B + b];
      }
    }
    const int ridx = in.ridx[q], invridx = in.invridx[q];
    double qir[3] = {in.q_ir[3 * q], in.q_ir[3 * q + 1], in.q_ir[3 * q + 2]};
    if (!live) nv = 0;  // keep the warp converged for the shuffles, write nothing useful
    double* vrow_p = vals_out + q * vrow;
    double* wrow_p = vecs_out + (compact_cap ? u / B : q) * wrow;
    if (nv == 0) {
      if (live) {  // failed point: zero row (the reference's outputs are zero-initialised)
        const uint32_t sv = dd.values.span * (dd.values.is_complex ? 2 : 1), sw = dd.vectors.span * (dd.vectors.is_complex ? 2 : 1);
        for (uint32_t s = lane; s < sv; s += LANES) vrow_p[(size_t)b * sv + s] = 0.0;
        for (uint32_t s = lane; s < sw; s += LANES) wrow_p[(size_t)b * sw + s] = 0.0;
      }
      // participate in the shuffles of the other groups of this warp: none are needed because every
      // shuffle below is confined to the LANES lanes of one group, which all share nv
      continue;
    }
    Rot rv = make_rot(dd.values, dd, ir != 0, ridx, invridx);
    Rot rw = make_rot(dd.vectors, dd, ir != 0, ridx, invridx);
    if (dd.values.is_complex) interp_unit<LANES, true>(dd.values, dd, false, nv, vtx, wgt, pb, b, lane, gmask, rv, qir, vrow_p);
    else interp_unit<LANES, false>(dd.values, dd, false, nv, vtx, wgt, pb, b, lane, gmask, rv, qir, vrow_p);
    if (dd.vectors.is_complex) interp_unit<LANES, true>(dd.vectors, dd, true, nv, vtx, wgt, pb, b, lane, gmask, rw, qir, wrow_p);
    else interp_unit<LANES, false>(dd.vectors, dd, true, nv, vtx, wgt, pb, b, lane, gmask, rw, qir, wrow_p);
  }
}

template <int LANES>
static cudaError_t launch_lanes(const DataDev& dd, const LocateIn& in, size_t n, int ir, double* vals, double* vecs,
                                int sm_count, cudaStream_t stream, const uint32_t* order, const uint32_t* segment, uint32_t compact_cap,
                                unsigned long long* overflow) {
  const int threads = 256;
  const size_t units = n * (size_t)dd.values.branches;
  size_t want = (units * LANES + threads - 1) / threads;
  size_t cap = (size_t)sm_count * (order ? 2 : 32);  // list mode: the segment is normally (almost) empty
  int blocks = (int)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  k_interp<LANES><<<blocks, threads, 0, stream>>>(dd, in, n, ir, vals, vecs, order, segment, compact_cap, overflow);
  return cudaGetLastError();
}

cudaError_t launch_interp(const DataDev& dd, const LocateIn& in, size_t n, int ir, double* vals, double* vecs, int sm_count,
                          cudaStream_t stream, const uint32_t* order, const uint32_t* segment, uint32_t compact_cap,
                          unsigned long long* overflow) {
  if (n == 0) return cudaSuccess;
  // lanes per (Q, mode) unit: enough to cover the items of the widest segment of the vectors' mode
  uint32_t items = dd.vectors.no1 > dd.vectors.no2 ? dd.vectors.no1 : dd.vectors.no2;
  if (dd.vectors.no0 > items) items = dd.vectors.no0;
  if (items <= 1) return launch_lanes<1>(dd, in, n, ir, vals, vecs, sm_count, stream, order, segment, compact_cap, overflow);
  if (items <= 2) return launch_lanes<2>(dd, in, n, ir, vals, vecs, sm_count, stream, order, segment, compact_cap, overflow);
  if (items <= 4) return launch_lanes<4>(dd, in, n, ir, vals, vecs, sm_count, stream, order, segment, compact_cap, overflow);
  if (items <= 8) return launch_lanes<8>(dd, in, n, ir, vals, vecs, sm_count, stream, order, segment, compact_cap, overflow);
  if (items <= 16) return launch_lanes<16>(dd, in, n, ir, vals, vecs, sm_count, stream, order, segment, compact_cap, overflow);
  return launch_lanes<32>(dd, in, n, ir, vals, vecs, sm_count, stream, order, segment, compact_cap, overflow);
}

}  // namespace b200
