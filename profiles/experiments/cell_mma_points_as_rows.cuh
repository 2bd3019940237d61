// cell_mma.cuh -- the compute pass of the pipelined cell kernel (cellinterp_tma.cu) on the FP64 tensor-core instruction.
//
// The weighted sum of interpolator_at.tpp:91-127 for the points of one cell,
//     out[point p, slot v, component j] = sum_corner  w[p][corner] * D[corner][v][j]          (complex D, real w)
// is a small GEMM with K = the 4 or 8 corners of the cell: exactly the shape of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  A warp
// takes 8 points (the M rows) and 4 "slots" -- 3-vectors (mode b, atom k) -- at a time: for each of the 3 components the 8
// columns of the B operand are (re, im) of that component of the 4 slots, so that after the three (or, for a cube, six chained)
// DMMAs lane (g, t) = (lane / 4, lane % 4) holds the complete complex 3-vector of slot t for point g and finishes it (rotation,
// atom permutation, Gamma phase, 48-byte store) without exchanging anything with another lane.
//
// Why: measured on the B200 (profiles/microbench/dmma_probe.cu) the instruction has the throughput of the FP64 pipe (37 TFLOP/s,
// no more than DFMA) and its result is BIT-IDENTICAL to the chain fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c)))) -- the corner
// loop the kernel ran before (cell_compute_pass) and the other interpolation kernels still run.  What changes is everything
// around the multiplications: one 8-byte shared-memory load per lane feeds 256 multiply-adds (was: five 16-byte loads per 24),
// the warp instruction count per (point, slot) drops about threefold, the rotation matrix, the point's indices and its weights are
// loop invariants of a lane (its point is fixed while it walks the slots), and a lane needs 6 accumulator registers instead of 48.
//
// The cell records carry the B operand ready-made (k_build_cell_table):
//     D'[group q][k-step s][component j][corner t of the step][slot a of the group][re, im]      (32 doubles per (q, s, j))
// slot v = (local mode) * SPM + atom, group q = v / 4, a = v % 4; SPM = slots per mode = n_atoms (1, 2 or a multiple of 4) or n_atoms
// rounded up to a multiple of 4 (zero columns), so that the atoms of a mode never straddle the lanes of two modes irregularly.
#pragma once
#include "cell_common.cuh"

namespace b200 {

__host__ __device__ inline uint32_t mma_slots_per_mode(uint32_t nat) { return (nat <= 2u || (nat & 3u) == 0u) ? nat : ((nat + 3u) & ~3u); }
__host__ __device__ inline uint32_t mma_groups(uint32_t modes, uint32_t nat) { return (modes * mma_slots_per_mode(nat) + 3u) / 4u; }
// bytes of the D' part of one tile (nv corners, mpp modes per pass)
__host__ __device__ inline size_t mma_d_bytes(uint32_t nv, uint32_t mpp, uint32_t nat) { return (size_t)mma_groups(mpp, nat) * (nv / 4u) * 768u; }
// position (in complex numbers) of (corner i, local mode bl, atom k, component j) inside D'
__host__ __device__ inline size_t mma_d_index(uint32_t nv, uint32_t nat, uint32_t i, uint32_t bl, uint32_t k, uint32_t j) {
  const uint32_t v = bl * mma_slots_per_mode(nat) + k, q = v >> 2, a = v & 3u, s = i >> 2, t = i & 3u, KS = nv >> 2;
  return ((((size_t)q * KS + s) * 3u + j) * 4u + t) * 4u + a;
}

// D(8x8) += A(8x4) * B(4x8): lane (g, t) supplies A[g][t], B[t][g] and owns D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// eigenvalues: plain weighted sum; a task is one value column for TQ consecutive points (as in cell_compute_pass)
template <int TQ>
__device__ __forceinline__ void cell_values_pass(const CellPass& c, int tid, int nthr) {
  const double* V = c.V;
  const double* W = c.W;
  const uint32_t CH = c.CH, mpp = c.mpp, no0v = c.no0v, len = c.len;
  const size_t vrow = (size_t)c.M * no0v;
  const uint32_t ntile = (len + TQ - 1) / TQ, per_v = c.mb * no0v;
  for (uint32_t task = tid; task < ntile * per_v; task += nthr) {
    const uint32_t tile = task / per_v, r = task - tile * per_v, t0 = tile * TQ;
    double acc[TQ];
#pragma unroll
    for (int t = 0; t < TQ; ++t) acc[t] = 0.0;
    for (int i = 0; i < c.NV; ++i) {
      const double v = V[(size_t)i * mpp * no0v + r];
      double w[TQ];
      load_tile<TQ>(W + (size_t)i * CH + t0, w);
#pragma unroll
      for (int t = 0; t < TQ; ++t) acc[t] = __fma_rn(w[t], v, acc[t]);
    }
    const uint32_t nt = min((uint32_t)TQ, len - t0);
    uint32_t qis[TQ];
    load_tile<TQ>(c.QI + t0, qis);
#pragma unroll
    for (int t = 0; t < TQ; ++t)
      if ((uint32_t)t < nt) c.vals_out[(size_t)qis[t] * vrow + (size_t)c.b0 * no0v + r] = acc[t];
  }
}

// The pass of one work item (<= CH points of one cell) over the modes [b0, b0 + mb).  All 32 lanes of a warp run every DMMA
// (the points past the end of the item carry zero weights: make_tables pads to a multiple of 8); only the finish is predicated.
// A unit of work = (8 points, a block of slot groups that holds whole modes); the warps draw units from the shared counter.
// SF: fused structure-factor finish (see cell_sf_pass in cell_common.cuh for the formula): |F|^2 per (point, mode) instead of
// the eigenvectors.
template <int TQ, bool SF>
__device__ __forceinline__ void cell_mma_pass(const CellPass& c, int tid, int nthr) {
  cell_values_pass<TQ>(c, tid, nthr);
  const double* const Dd = reinterpret_cast<const double*>(c.D);
  const double* const W = c.W;
  const double2* const PH = c.PH;
  const uint32_t CH = c.CH, mb = c.mb, b0 = c.b0, M = c.M, S = c.S, NAT = c.NAT, G = c.G, len = c.len;
  const int kind = c.kind;
  const bool gamma = c.gamma;
  const uint32_t lane = (uint32_t)tid & 31u, g = lane >> 2, t = lane & 3u;
  const uint32_t KS = (uint32_t)c.NV >> 2;                 // k-steps: 1 tetrahedron, 2 cube
  const uint32_t SPM = mma_slots_per_mode(NAT);
  const uint32_t n_groups = (mb * SPM + 3u) >> 2;          // slot groups that hold the modes of this pass
  const uint32_t gpb = SPM >= 4u ? (SPM >> 2) : 1u;        // groups per block: a block never splits a mode
  const uint32_t n_blocks = (n_groups + gpb - 1u) / gpb;
  const uint32_t n_mt = (len + 7u) >> 3;                   // tiles of 8 points
  const uint32_t nwarp = (uint32_t)nthr >> 5;
  // about four units per warp when the item is large enough for that: short items are cut finer along the slots
  uint32_t cuts = (4u * nwarp + n_mt - 1u) / n_mt;
  cuts = cuts < 1u ? 1u : (cuts > n_blocks ? n_blocks : cuts);
  const uint32_t bpu = (n_blocks + cuts - 1u) / cuts;      // blocks per unit
  cuts = (n_blocks + bpu - 1u) / bpu;
  const uint32_t n_units = n_mt * cuts;
  const size_t wrow = (size_t)M * S;                       // complex numbers per output row
  const uint32_t spm_magic = 0xffffffffu / SPM + 1u;       // floor(v / SPM) == umulhi(v, magic) for v * SPM < 2^32
  const bool spm_pow2 = (SPM & (SPM - 1u)) == 0u;
  const uint32_t spm_shift = 31u - (uint32_t)__clz(SPM);
  const double sgn = SF ? (c.conjugate ? -1.0 : 1.0) : 1.0;
  for (;;) {
    uint32_t u = 0;
    if (lane == 0) u = atomicAdd(c.task_ctr, 1u);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= n_units) break;
    const uint32_t mt = u / cuts, cut = u - mt * cuts;
    const uint32_t p = mt * 8u + g;                        // this lane's point (position in the item)
    const bool pvalid = p < len;
    const double a0 = W[(size_t)t * CH + p];
    const double a1 = KS == 2u ? W[(size_t)(4u + t) * CH + p] : 0.0;
    const uint32_t rr = c.RI[p], qi = c.QI[p], ri = rr & 0xffffu;
    double R[9];
    if (!SF && kind >= 0) {
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = c.RS[9u * ri + i];
    }
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    if (SF) {
      const double* gp = c.QV + 3 * (size_t)p;
      g0 = gp[0]; g1 = gp[1]; g2 = gp[2];
    }
    double Fr_sum = 0.0, Fi_sum = 0.0;                     // (SF, SPM > 4) the groups of a mode are added up in order
    const uint32_t q_lo = cut * bpu * gpb, q_hi = min(n_groups, (cut + 1u) * bpu * gpb);
    for (uint32_t q = q_lo; q < q_hi; ++q) {
      const uint32_t v = 4u * q + t;
      const uint32_t b = spm_pow2 ? (v >> spm_shift) : __umulhi(v, spm_magic), k = v - b * SPM;
      const double* Bq = Dd + (size_t)q * KS * 96u + t * 8u + g;
      double2 acc0 = make_double2(0.0, 0.0), acc1 = acc0, acc2 = acc0;
      dmma_8x8x4(acc0.x, acc0.y, a0, Bq[0]);
      dmma_8x8x4(acc1.x, acc1.y, a0, Bq[32]);
      dmma_8x8x4(acc2.x, acc2.y, a0, Bq[64]);
      if (KS == 2u) {
        dmma_8x8x4(acc0.x, acc0.y, a1, Bq[96]);
        dmma_8x8x4(acc1.x, acc1.y, a1, Bq[128]);
        dmma_8x8x4(acc2.x, acc2.y, a1, Bq[160]);
      }
      const bool atom = k < NAT && b < mb;
      if (!SF) {
        // ---- finish: rotation, atom permutation, Gamma phase, store (the same operations, in the same order, as the other
        // interpolation kernels: rotate_phase) ------------------------------------------------------------------------------
        if (pvalid && atom) {
          double2* out = reinterpret_cast<double2*>(c.vecs_out) + (size_t)qi * wrow + (size_t)(b0 + b) * S;
          uint32_t dest = k;
          if (kind >= 0) {
            double2 ph = make_double2(1.0, 0.0);
            if (gamma) {
              dest = c.F0[k * G + ri];
              ph = PH[(size_t)p * NAT + k];
            }
            out += 3 * dest;
            rotate_phase_store(R, acc0, acc1, acc2, ph, gamma, out);
            if (kind == 2) {  // axial: det(R) R^-1 v
              const double det = c.rot_det[rr >> 16];
#pragma unroll
              for (int e = 0; e < 3; ++e) out[e] = make_double2(out[e].x * det, out[e].y * det);  // (rare path: read back)
            }
          } else {
            store48(out + 3 * dest, acc0, acc1, acc2);
          }
        }
      } else {
        // ---- fused structure factor: qv . (R a) = (qv^T R) . a with the row vector g of the point; times the per-(point,
        // source atom) factor; summed over the atoms of the mode: the lanes of the quad (butterfly), then the groups of the mode
        const double dr = __fma_rn(g2, acc2.x, __fma_rn(g1, acc1.x, __dmul_rn(g0, acc0.x)));
        const double di = sgn * __fma_rn(g2, acc2.y, __fma_rn(g1, acc1.y, __dmul_rn(g0, acc0.y)));
        double Fr = 0.0, Fi = 0.0;
        if (atom) {
          const double2 f = PH[(size_t)p * NAT + k];
          Fr = __fma_rn(-f.y, di, __dmul_rn(f.x, dr));
          Fi = __fma_rn(f.y, dr, __dmul_rn(f.x, di));
        }
        if (SPM >= 2u) { Fr += __shfl_xor_sync(0xffffffffu, Fr, 1); Fi += __shfl_xor_sync(0xffffffffu, Fi, 1); }
        if (SPM >= 4u) { Fr += __shfl_xor_sync(0xffffffffu, Fr, 2); Fi += __shfl_xor_sync(0xffffffffu, Fi, 2); }
        bool last = true;
        if (SPM > 4u) {
          const uint32_t in_mode = q - b * gpb;   // (SPM multiple of 4: the groups of mode b are [b * gpb, (b + 1) * gpb))
          Fr_sum = in_mode == 0u ? Fr : Fr_sum + Fr;
          Fi_sum = in_mode == 0u ? Fi : Fi_sum + Fi;
          Fr = Fr_sum; Fi = Fi_sum;
          last = in_mode + 1u == gpb;
        }
        if (last && pvalid && k == 0u && b < mb) c.sf_out[(size_t)qi * M + b0 + b] = __fma_rn(Fi, Fi, __dmul_rn(Fr, Fr));
      }
    }
  }
}

}  // namespace b200
