// cell_common.cuh -- pieces shared by the two cell-batched interpolation kernels (cellinterp.cu, cellinterp_ws.cu)
#pragma once
#include "device_tables.cuh"
#include "brille_b200.h"

namespace b200 {

// Store 48 contiguous bytes at a 16-byte aligned address as one aligned 32-byte store (sm_100 STG.256) plus one 16-byte
// store.  Consecutive lanes hold consecutive 48-byte pieces, so within each of the two instructions the lanes cover whole
// 32-byte sectors: L2 receives one full-sector write per sector instead of three partial ones from three 16-byte stores.
__device__ __forceinline__ void store32(double2* p, const double2 lo, const double2 hi) {  // p 32-byte aligned
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(lo.x), "d"(lo.y), "d"(hi.x), "d"(hi.y) : "memory");
}
__device__ __forceinline__ void store16(double2* p, const double2 v) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void store48(double2* out, const double2 o0, const double2 o1, const double2 o2) {
  const bool even = (reinterpret_cast<uintptr_t>(out) & 31u) == 0;
  store32(even ? out : out + 1, even ? o0 : o1, even ? o1 : o2);
  store16(even ? out + 2 : out, even ? o2 : o0);
}

// u = ph * (R * a) for one complex 3-vector, written with explicit fused/unfused operations so that every code path
// that finishes a point produces the same bits (the result must not depend on which path a point happens to take)
__device__ __forceinline__ void rotate_phase(const double* R, const double2 a0, const double2 a1, const double2 a2, const double2 ph,
                                             bool use_phase, double2& u0, double2& u1, double2& u2) {
  u0.x = __fma_rn(R[2], a2.x, __fma_rn(R[1], a1.x, __dmul_rn(R[0], a0.x)));
  u0.y = __fma_rn(R[2], a2.y, __fma_rn(R[1], a1.y, __dmul_rn(R[0], a0.y)));
  u1.x = __fma_rn(R[5], a2.x, __fma_rn(R[4], a1.x, __dmul_rn(R[3], a0.x)));
  u1.y = __fma_rn(R[5], a2.y, __fma_rn(R[4], a1.y, __dmul_rn(R[3], a0.y)));
  u2.x = __fma_rn(R[8], a2.x, __fma_rn(R[7], a1.x, __dmul_rn(R[6], a0.x)));
  u2.y = __fma_rn(R[8], a2.y, __fma_rn(R[7], a1.y, __dmul_rn(R[6], a0.y)));
  if (use_phase) {
    u0 = make_double2(__fma_rn(-ph.y, u0.y, __dmul_rn(ph.x, u0.x)), __fma_rn(ph.y, u0.x, __dmul_rn(ph.x, u0.y)));
    u1 = make_double2(__fma_rn(-ph.y, u1.y, __dmul_rn(ph.x, u1.x)), __fma_rn(ph.y, u1.x, __dmul_rn(ph.x, u1.y)));
    u2 = make_double2(__fma_rn(-ph.y, u2.y, __dmul_rn(ph.x, u2.x)), __fma_rn(ph.y, u2.x, __dmul_rn(ph.x, u2.y)));
  }
}
__device__ __forceinline__ void rotate_phase_store(const double* R, const double2 a0, const double2 a1, const double2 a2,
                                                   const double2 ph, bool use_phase, double2* out) {
  double2 u0, u1, u2;
  rotate_phase(R, a0, a1, a2, ph, use_phase, u0, u1, u2);
  store48(out, u0, u1, u2);
}

// Phase that aligns a vertex' eigenvector to the pivot's (utilities.tpp:567-579): z = <d_pivot|d_v>, factor
// e^{-i arg z} = conj(z)/|z| (the reference evaluates polar(1, -atan2(Im z, Re z)), the same number up to rounding;
// z == 0 gives 1 in both).  Explicit fused operations: the on-the-fly kernel and the cell-table builder must agree bitwise.
__device__ __forceinline__ void align_accumulate(const double2 p0, const double2 x, double& re, double& im) {
  re = __fma_rn(p0.y, x.y, __fma_rn(p0.x, x.x, re));
  im = __fma_rn(-p0.y, x.x, __fma_rn(p0.x, x.y, im));
}
__device__ __forceinline__ double2 align_factor(double re, double im) {
  const double m = fmax(fabs(re), fabs(im));
  double2 f = make_double2(1.0, 0.0);
  if (m > 0.0) {
    const double r = re / m, q = im / m;
    const double n = 1.0 / sqrt(__fma_rn(q, q, __dmul_rn(r, r)));
    f = make_double2(__dmul_rn(r, n), -__dmul_rn(q, n));
  }
  return f;
}
__device__ __forceinline__ double2 align_apply(const double2 f, const double2 x) {
  return make_double2(__fma_rn(-f.y, x.y, __dmul_rn(f.x, x.x)), __fma_rn(f.y, x.x, __dmul_rn(f.x, x.y)));
}

// dynamic shared memory carve-up (all offsets 16-byte aligned)
struct SmemPlan {
  size_t D, V, W, PH, RS, F0, QI, RI, PHI, total;
};
__host__ __device__ inline SmemPlan plan_smem(uint32_t nvmax, uint32_t mpp, uint32_t S, uint32_t no0v, uint32_t chunk, uint32_t n_at, uint32_t G, bool gamma) {
  SmemPlan p;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) / 16 * 16; return at; };
  p.D = take((size_t)nvmax * mpp * S * 16);  // nvmax = 8 when the grid has cube cells, else 4
  p.V = take((size_t)nvmax * mpp * no0v * 8);
  p.W = take((size_t)chunk * 8 * 8);
  p.PH = take(gamma ? (size_t)chunk * n_at * 16 : 0);
  p.RS = take((size_t)G * 9 * 8);
  p.F0 = take(gamma ? (size_t)n_at * G * 4 : 0);
  p.QI = take((size_t)chunk * 4);
  p.RI = take((size_t)chunk * 4);
  p.PHI = take((size_t)(nvmax - 1) * mpp * 16);
  p.total = o;
  return p;
}


// Everything one pass (a block of `mb` modes starting at b0) of one work item needs; all pointers are shared memory
// except the outputs.  `tid`/`nthr` are the index and count of the threads that execute the pass together.
struct CellPass {
  const double2* D;   // [NV][mpp][S] permuted, phase-aligned vertex rows
  const double* V;    // [NV][mpp][no0v] permuted eigenvalue rows
  const double* W;    // [NV][CH] weights, transposed
  const double2* PH;  // [CH][NAT] Gamma phases
  const double* RS;   // [G][9] rotation matrices applied to the vectors
  const uint32_t* F0; // [NAT][G] atom permutation
  const uint32_t* QI; // [CH] point index
  const uint32_t* RI; // [CH] matrix index | Ridx << 16
  uint32_t CH, mpp, mb, b0, len, M, S, NAT, no0v, G;
  int NV, kind;
  bool gamma;
  const double* rot_det;
  double* vals_out;
  double* vecs_out;
  uint32_t* task_ctr;  // shared counter (zero at the start of the pass) for the dynamic deal of tasks, or nullptr
  // fused structure-factor finish (cell_sf_pass): PH then holds the combined per-(point, SOURCE atom) factor
  //   coef_l e^{-W_l} e^{2 pi i Q.r_l} * [conj](Gamma phase)   with l = F0(k, R) the destination atom,
  const double* QV;   // [CH][3] g = (T Q)^T R of every point (R the point's rotation): the finish is g . a
  double* sf_out;     // (n, M) |F|^2
  int conjugate;
};

// TQ consecutive doubles / 32-bit words of a shared-memory array (TQ = 2 or 4, 8- resp. 16-byte aligned)
template <int TQ>
__device__ __forceinline__ void load_tile(const double* p, double* w) {
  if (TQ == 4) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
  } else {
    const double2 a = *reinterpret_cast<const double2*>(p);
    w[0] = a.x; w[1] = a.y;
  }
}
template <int TQ>
__device__ __forceinline__ void load_tile(const uint32_t* p, uint32_t* w) {
  if (TQ == 4) {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
  } else {
    const uint2 a = *reinterpret_cast<const uint2*>(p);
    w[0] = a.x; w[1] = a.y;
  }
}

// TQ = points per register tile: 4 (48 accumulator registers, 2 CTAs/SM) or 2 (24, 3 CTAs/SM).  The arithmetic per point
// does not depend on TQ: both give the same bits.
template <int TQ>
__device__ __forceinline__ void cell_compute_pass(const CellPass& c, int tid, int nthr) {
  const double2* D = c.D;
  const double* V = c.V;
  const double* W = c.W;
  const double2* PH = c.PH;
  const double* RS = c.RS;
  const uint32_t* F0 = c.F0;
  const uint32_t* QI = c.QI;
  const uint32_t* RI = c.RI;
  const uint32_t CH = c.CH, mpp = c.mpp, mb = c.mb, b0 = c.b0, M = c.M, S = c.S, NAT = c.NAT, no0v = c.no0v, G = c.G;
  const int NV = c.NV, kind = c.kind;
  const bool gamma = c.gamma;
  const size_t vrow = (size_t)M * no0v, wrow = (size_t)M * S;
  struct { uint32_t len; } item = {c.len};
  struct { const double* rot_det; } dd_ = {c.rot_det};
  struct { double* vals_out; double* vecs_out; decltype(dd_) dd; } a = {c.vals_out, c.vecs_out, dd_};
    // ---- eigenvalues: plain weighted sum; a task is one value column for TQ consecutive points ------------------------
    const uint32_t ntile = (item.len + TQ - 1) / TQ;
    {
      const uint32_t per_v = mb * no0v;
      for (uint32_t task = tid; task < ntile * per_v; task += nthr) {
        const uint32_t tile = task / per_v, r = task - tile * per_v, t0 = tile * TQ;
        double acc[TQ];
#pragma unroll
        for (int t = 0; t < TQ; ++t) acc[t] = 0.0;
        for (int i = 0; i < NV; ++i) {
          const double v = V[(size_t)i * mpp * no0v + r];
          double w[TQ];
          load_tile<TQ>(W + (size_t)i * CH + t0, w);
#pragma unroll
          for (int t = 0; t < TQ; ++t) acc[t] = __fma_rn(w[t], v, acc[t]);
        }
        const uint32_t nt = min((uint32_t)TQ, item.len - t0);
        uint32_t qis[TQ];
        load_tile<TQ>(QI + t0, qis);
#pragma unroll
        for (int t = 0; t < TQ; ++t)
          if ((uint32_t)t < nt) a.vals_out[(size_t)qis[t] * vrow + (size_t)b0 * no0v + r] = acc[t];
      }
    }
    // ---- eigenvectors: weighted sum of pre-phased rows, rotation, atom permutation, Gamma phase ----------------------
    // A task is one 3-vector (mode b, atom k) for TQ consecutive points: the three complex numbers of every corner are
    // read from shared memory once and reused for the TQ points (register tile), which makes the loop FP64-bound
    // instead of shared-memory-bound.
    const uint32_t per_q = mb * NAT;
    // task -> (tile, r = b * NAT + k) is advanced incrementally: no integer division in the loop
    const uint32_t step_tile = (uint32_t)nthr / per_q, step_r = (uint32_t)nthr - step_tile * per_q;
    const uint32_t nat_magic = 0xffffffffu / NAT + 1u;  // floor(r / NAT) == umulhi(r, magic) for r * NAT < 2^32
    const uint32_t pq_magic = 0xffffffffu / per_q + 1u;
    const uint32_t n_task = ntile * per_q;
    uint32_t tile = (uint32_t)tid / per_q, r = (uint32_t)tid - tile * per_q;
    // Static deal (task = tid, tid + nthr, ...) or, when the caller provides a shared counter, warps draw rounds of 32
    // consecutive tasks from it: a warp that was held up draws fewer rounds, so all warps reach the barrier that ends the
    // pass within one round of each other.
    for (uint32_t task = tid;; task += nthr, tile += step_tile, r += step_r) {
      if (c.task_ctr) {
        uint32_t base = 0;
        if ((tid & 31) == 0) base = atomicAdd(c.task_ctr, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_task) break;
        task = base + (uint32_t)(tid & 31);
        if (task >= n_task) continue;  // (only the last round of a pass is partial)
        tile = per_q == 1u ? task : __umulhi(task, pq_magic);
        r = task - tile * per_q;
      } else {
        if (task >= n_task) break;
        if (r >= per_q) { r -= per_q; ++tile; }
      }
      const uint32_t b = NAT == 1u ? r : __umulhi(r, nat_magic), k = r - b * NAT;
      const uint32_t t0 = tile * TQ;
      const double2* src = D + (size_t)b * S + 3 * k;
      double2 acc[TQ][3];
#pragma unroll
      for (int t = 0; t < TQ; ++t) acc[t][0] = acc[t][1] = acc[t][2] = make_double2(0.0, 0.0);
      for (int i = 0; i < NV; ++i) {
        const double2* x = src + (size_t)i * mpp * S;
        const double2 x0 = x[0], x1 = x[1], x2 = x[2];
        double w[TQ];
        load_tile<TQ>(W + (size_t)i * CH + t0, w);
#pragma unroll
        for (int t = 0; t < TQ; ++t) {
          acc[t][0].x += w[t] * x0.x; acc[t][0].y += w[t] * x0.y;
          acc[t][1].x += w[t] * x1.x; acc[t][1].y += w[t] * x1.y;
          acc[t][2].x += w[t] * x2.x; acc[t][2].y += w[t] * x2.y;
        }
      }
      // ---- finish: rotation, atom permutation, Gamma phase, store ------------------------------------------------------
      const uint32_t nt = min((uint32_t)TQ, item.len - t0);
      uint32_t rrs[TQ], qis[TQ];
      load_tile<TQ>(RI + t0, rrs);
      load_tile<TQ>(QI + t0, qis);
      double2* const out_base = reinterpret_cast<double2*>(a.vecs_out) + (size_t)(b0 + b) * S;
      if (gamma && nt == TQ && (rrs[0] & 0xffffu) == (rrs[TQ - 1] & 0xffffu) && ((wrow & 1) == 0)) {
        // The four points share the rotation (the sort is by cell and operation): one matrix, one destination atom, no
        // branches.  The 48 output bytes of a point go out as one 32-byte-aligned 32-byte store plus one 16-byte store
        // (see store48); which of the three components forms the aligned pair depends only on the parity of the
        // destination (rows are a multiple of 32 bytes here), so the ROWS of the matrix are loaded in store order
        // (pair, pair, single) and no data has to be shuffled afterwards.
        const uint32_t ri = rrs[0] & 0xffffu;
        const uint32_t dest = F0[k * G + ri];
        double2* const out0 = out_base + 3 * dest;
        const bool even = (reinterpret_cast<uintptr_t>(out0) & 31u) == 0;  // row starts are 32-byte aligned
        const double* Rs = RS + 9 * ri;
        const double* r0 = Rs + (even ? 0 : 3);  // rows of the aligned pair ...
        const double* r1 = Rs + (even ? 3 : 6);
        const double* r2 = Rs + (even ? 6 : 0);  // ... and of the single component
        const double R[9] = {r0[0], r0[1], r0[2], r1[0], r1[1], r1[2], r2[0], r2[1], r2[2]};
        const uint32_t off32 = even ? 0u : 1u, off16 = even ? 2u : 0u;
        const double2* php = PH + (size_t)t0 * NAT + k;
#pragma unroll
        for (int t = 0; t < TQ; ++t) {
          double2 u0, u1, u2;
          rotate_phase(R, acc[t][0], acc[t][1], acc[t][2], php[(size_t)t * NAT], true, u0, u1, u2);
          double2* const o = out0 + (size_t)qis[t] * wrow;
          store32(o + off32, u0, u1);
          store16(o + off16, u2);
        }
        continue;
      }
#pragma unroll
      for (int t = 0; t < TQ; ++t) {
        if ((uint32_t)t >= nt) break;
        const uint32_t qi = t0 + t;
        uint32_t dest = k;
        double2* out = out_base + (size_t)qis[t] * wrow;
        if (kind >= 0) {
          const uint32_t rr = rrs[t];
          const uint32_t ri = rr & 0xffffu;
          double2 ph = make_double2(1.0, 0.0);
          if (gamma) {
            dest = F0[k * G + ri];
            ph = PH[(size_t)qi * NAT + k];
          }
          out += 3 * dest;
          rotate_phase_store(RS + 9 * ri, acc[t][0], acc[t][1], acc[t][2], ph, gamma, out);
          if (kind == 2) {  // axial: det(R) R^-1 v
            const double det = a.dd.rot_det[rr >> 16];
#pragma unroll
            for (int c = 0; c < 3; ++c) out[c] = make_double2(out[c].x * det, out[c].y * det);  // (rare path: read back)
          }
        } else {
          out += 3 * dest;
          store48(out, acc[t][0], acc[t][1], acc[t][2]);
        }
      }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Fused structure-factor finish: the same weighted sum and rotation as cell_compute_pass, but instead of
// storing the rotated 3-vector u_l = ph * R a_k of (mode b, atom k -> l) its term of
//     F(Q, b) = sum_l c_l e^{-W_l} e^{2 pi i Q.r_l} (qv . u_l^[*])
// is formed in registers and summed over the atoms with shuffles inside the NAT consecutive lanes that hold the atoms of one
// mode (NAT <= 32, padded to a power of two NATP with idle lanes: lane groups are aligned because the tasks of a tile are dealt
// r = b * NATP + k fastest and both the thread count and the tasks per tile are multiples of NATP).  The scalar Gamma phase commutes with the dot product, so it is
// folded into the per-(point, atom) factor PH: one complex multiply per term instead of three.  One code path for every
// point (the rotation is looked up per point), so the result does not depend on the composition of a tile or an item: it is
// bit-reproducible across chunkings.  The eigenvectors are never written: 8 bytes per (Q, mode) leave the SM.
// ---------------------------------------------------------------------------------------------------------------------
template <int TQ>
__device__ __forceinline__ void cell_sf_pass(const CellPass& c, int tid, int nthr) {
  const double2* D = c.D;
  const double* V = c.V;
  const double* W = c.W;
  const double2* PH = c.PH;
  const double* QV = c.QV;
  const uint32_t* QI = c.QI;
  const uint32_t CH = c.CH, mpp = c.mpp, mb = c.mb, b0 = c.b0, M = c.M, S = c.S, NAT = c.NAT, no0v = c.no0v;
  const int NV = c.NV;
  const size_t vrow = (size_t)M * no0v;
  const uint32_t len = c.len;
  const uint32_t ntile = (len + TQ - 1) / TQ;
  {  // eigenvalues: as in cell_compute_pass
    const uint32_t per_v = mb * no0v;
    for (uint32_t task = tid; task < ntile * per_v; task += nthr) {
      const uint32_t tile = task / per_v, r = task - tile * per_v, t0 = tile * TQ;
      double acc[TQ];
#pragma unroll
      for (int t = 0; t < TQ; ++t) acc[t] = 0.0;
      for (int i = 0; i < NV; ++i) {
        const double v = V[(size_t)i * mpp * no0v + r];
        double w[TQ];
        load_tile<TQ>(W + (size_t)i * CH + t0, w);
#pragma unroll
        for (int t = 0; t < TQ; ++t) acc[t] = __fma_rn(w[t], v, acc[t]);
      }
      const uint32_t nt = min((uint32_t)TQ, len - t0);
      uint32_t qis[TQ];
      load_tile<TQ>(QI + t0, qis);
#pragma unroll
      for (int t = 0; t < TQ; ++t)
        if ((uint32_t)t < nt) c.vals_out[(size_t)qis[t] * vrow + (size_t)b0 * no0v + r] = acc[t];
    }
  }
  // atoms per mode padded to a power of two (NATP lanes per mode; the lanes k >= NAT idle): the lane groups stay aligned for
  // any number of atoms <= 32
  uint32_t NATP = 1;
  while (NATP < NAT) NATP <<= 1;
  const uint32_t natp_shift = 31u - (uint32_t)__clz(NATP);
  const uint32_t per_q = mb * NATP;
  const uint32_t step_tile = (uint32_t)nthr / per_q, step_r = (uint32_t)nthr - step_tile * per_q;
  const uint32_t n_task = ntile * per_q;
  const uint32_t lane = (uint32_t)tid & 31u;
  const double sgn = c.conjugate ? -1.0 : 1.0;
  const uint32_t dstep = mpp * S;  // elements between the rows of two corners
  uint32_t tile = (uint32_t)tid / per_q, r = (uint32_t)tid - tile * per_q;
  // The loop is WARP-uniform (it runs while the first lane of the warp has a task; lanes past the end redo task 0 and store
  // nothing) and so is everything inside it (the points past the end of the last tile carry zero weights and are only kept from
  // storing): the shuffles can name the full warp, which costs one instruction each instead of a guarded sequence.
  for (uint32_t task = tid; task - lane < n_task; task += nthr, tile += step_tile, r += step_r) {
    if (r >= per_q) { r -= per_q; ++tile; }
    const bool live = task < n_task;
    const uint32_t tl = live ? tile : 0u, rl = live ? r : 0u;
    const uint32_t b = rl >> natp_shift, kp = rl & (NATP - 1u);
    const bool atom = kp < NAT;  // (a padding lane: contributes zero)
    const uint32_t k = atom ? kp : 0u;
    const uint32_t t0 = tl * TQ;
    const double2* src = D + (size_t)b * S + 3 * k;
    double2 acc[TQ][3];
    const double2* x = src;
    const double* wp = W + t0;
    {  // first corner: w * x (== fma(w, x, 0) up to the sign of a zero) instead of zeroing 6 * TQ accumulators first
      const double2 x0 = x[0], x1 = x[1], x2 = x[2];
      double w[TQ];
      load_tile<TQ>(wp, w);
#pragma unroll
      for (int t = 0; t < TQ; ++t) {
        acc[t][0] = make_double2(__dmul_rn(w[t], x0.x), __dmul_rn(w[t], x0.y));
        acc[t][1] = make_double2(__dmul_rn(w[t], x1.x), __dmul_rn(w[t], x1.y));
        acc[t][2] = make_double2(__dmul_rn(w[t], x2.x), __dmul_rn(w[t], x2.y));
      }
    }
    for (int i = 1; i < NV; ++i) {  // (running pointers: no multiplications in the loop)
      x += dstep;
      wp += CH;
      const double2 x0 = x[0], x1 = x[1], x2 = x[2];
      double w[TQ];
      load_tile<TQ>(wp, w);
#pragma unroll
      for (int t = 0; t < TQ; ++t) {
        acc[t][0].x = __fma_rn(w[t], x0.x, acc[t][0].x); acc[t][0].y = __fma_rn(w[t], x0.y, acc[t][0].y);
        acc[t][1].x = __fma_rn(w[t], x1.x, acc[t][1].x); acc[t][1].y = __fma_rn(w[t], x1.y, acc[t][1].y);
        acc[t][2].x = __fma_rn(w[t], x2.x, acc[t][2].x); acc[t][2].y = __fma_rn(w[t], x2.y, acc[t][2].y);
      }
    }
    const uint32_t nt = live ? min((uint32_t)TQ, len - t0) : 0u;
    uint32_t qis[TQ];
    load_tile<TQ>(QI + t0, qis);
#pragma unroll
    for (int t = 0; t < TQ; ++t) {
      // qv . (R a) = (qv^T R) . a: the row vector g = qv^T R is per point and comes ready from the item's tables
      const double* gp = QV + 3 * (size_t)(t0 + t);
      const double g0 = gp[0], g1 = gp[1], g2 = gp[2];
      const double dr = __fma_rn(g2, acc[t][2].x, __fma_rn(g1, acc[t][1].x, __dmul_rn(g0, acc[t][0].x)));
      const double di = sgn * __fma_rn(g2, acc[t][2].y, __fma_rn(g1, acc[t][1].y, __dmul_rn(g0, acc[t][0].y)));
      const double2 f = PH[(size_t)(t0 + t) * NAT + k];
      double Fr = atom ? __fma_rn(-f.y, di, __dmul_rn(f.x, dr)) : 0.0;
      double Fi = atom ? __fma_rn(f.y, dr, __dmul_rn(f.x, di)) : 0.0;
      // butterfly over the atoms of the mode (every lane ends with the same sum); the stages are spelled out, under uniform
      // predicates, so that no loop counter lives next to the accumulators
#define B200_SF_STAGE(o_) if (NATP > (o_)) { Fr += __shfl_xor_sync(0xffffffffu, Fr, (o_)); Fi += __shfl_xor_sync(0xffffffffu, Fi, (o_)); }
      B200_SF_STAGE(1) B200_SF_STAGE(2) B200_SF_STAGE(4) B200_SF_STAGE(8) B200_SF_STAGE(16)
#undef B200_SF_STAGE
      if (kp == 0 && (uint32_t)t < nt) c.sf_out[(size_t)qis[t] * M + b0 + b] = __fma_rn(Fi, Fi, __dmul_rn(Fr, Fr));
    }
  }
}

}  // namespace b200
