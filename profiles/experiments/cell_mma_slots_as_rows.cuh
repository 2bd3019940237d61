// cell_mma.cuh -- the compute pass of the pipelined cell kernel (cellinterp_tma.cu) on the FP64 tensor-core instruction.
//
// The weighted sum of interpolator_at.tpp:91-127 for the points of one cell,
//     out[slot v, component j][point p] = sum_corner  D[corner][v][j] * w[corner][p]          (complex D, real w)
// is a small GEMM with K = the 4 or 8 corners of the cell: exactly the shape of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  A "slot"
// is one 3-vector (mode b, atom k).  A warp step takes 8 slots (the M rows) and 8 points (the N columns): one DMMA (two chained
// ones for a cube) per scalar c = (component, re / im), six scalars.  In the accumulator layout of the instruction lane
// (g, t) = (lane / 4, lane % 4) then owns ALL six scalars of slot g for the two points 2t, 2t+1 of the tile and finishes both
// (rotation, atom permutation, Gamma phase, 48-byte store) without exchanging anything with another lane.  Within one store
// instruction the 8 lanes with the same t write 8 consecutive 48-byte pieces of one output row: 4 rows x 384 contiguous bytes.
// (The transposed assignment -- points as rows, (slot, re / im) as columns, one point per lane -- was built first: identical
// results, but a store instruction then covers 8 rows x 192 bytes, L2 sees 1.8x the write requests and the kernel is slower than
// the FMA version: profiles/README.md.)
//
// Why: measured on the B200 (profiles/microbench/dmma_probe.cu) the instruction has the throughput of the FP64 pipe (37 TFLOP/s,
// no more than DFMA) and its result is BIT-IDENTICAL to the chain fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c)))) -- the corner
// loop the kernel ran before (cell_compute_pass) and the other interpolation kernels still run.  What changes is everything
// around the multiplications: one 8-byte shared-memory load per lane feeds 256 multiply-adds (was: five 16-byte loads per 24),
// the slot of a lane and everything derived from it is shared by its two points, and a lane needs 12 accumulator registers
// instead of 48.
//
// The cell records carry the A operand ready-made (k_build_cell_table):
//     D'[group Q][k-step s][scalar c][slot g of the group][corner t of the step]                 (32 doubles per (Q, s, c), lane order)
// slot v = (local mode) * SPM + atom, group Q = v / 8, g = v % 8; SPM = slots per mode = n_atoms if that is 1, 2, 4 or a multiple
// of 8, else n_atoms rounded up to 4 or to a multiple of 8 (zero rows): the atoms of a mode never straddle two groups irregularly.
#pragma once
#include "cell_common.cuh"

namespace b200 {

__host__ __device__ inline uint32_t mma_slots_per_mode(uint32_t nat) {
  if (nat <= 2u || nat == 4u || (nat & 7u) == 0u) return nat;
  return nat == 3u ? 4u : ((nat + 7u) & ~7u);
}
__host__ __device__ inline uint32_t mma_groups(uint32_t modes, uint32_t nat) { return (modes * mma_slots_per_mode(nat) + 7u) / 8u; }
// bytes of the D' part of one tile (nv corners, mpp modes per pass)
__host__ __device__ inline size_t mma_d_bytes(uint32_t nv, uint32_t mpp, uint32_t nat) { return (size_t)mma_groups(mpp, nat) * (nv / 4u) * 1536u; }
// position (in doubles) of scalar c = 2 * component + (0 re, 1 im) of (corner i, local mode bl, atom k) inside D'
__host__ __device__ inline size_t mma_d_index(uint32_t nv, uint32_t nat, uint32_t i, uint32_t bl, uint32_t k, uint32_t c) {
  const uint32_t v = bl * mma_slots_per_mode(nat) + k, Q = v >> 3, g = v & 7u, s = i >> 2, t = i & 3u, KS = nv >> 2;
  return ((((size_t)Q * KS + s) * 6u + c) * 8u + g) * 4u + t;
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
  const uint32_t n_groups = (mb * SPM + 7u) >> 3;          // groups of 8 slots that hold the modes of this pass
  const uint32_t gpb = SPM >= 8u ? (SPM >> 3) : 1u;        // groups per block: a block never splits a mode
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
    const uint32_t p0 = mt * 8u;
    // B operand: the weights of point p0 + g at the corners t (and 4 + t)
    const double w0 = W[(size_t)t * CH + p0 + g];
    const double w1 = KS == 2u ? W[(size_t)(4u + t) * CH + p0 + g] : 0.0;
    // this lane's two points (positions in the item): the columns 2t, 2t+1 of the accumulators
    const uint32_t pA = p0 + 2u * t, pB = pA + 1u;
    const bool validA = pA < len, validB = pB < len;
    const uint2 rr2 = *reinterpret_cast<const uint2*>(c.RI + pA), qi2 = *reinterpret_cast<const uint2*>(c.QI + pA);
    const uint32_t riA = rr2.x & 0xffffu, riB = rr2.y & 0xffffu;
    double R[9];
    if (!SF && kind >= 0) {
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = c.RS[9u * riA + i];
    }
    double gA0 = 0.0, gA1 = 0.0, gA2 = 0.0, gB0 = 0.0, gB1 = 0.0, gB2 = 0.0;
    if (SF) {
      const double* gp = c.QV + 3 * (size_t)pA;
      gA0 = gp[0]; gA1 = gp[1]; gA2 = gp[2]; gB0 = gp[3]; gB1 = gp[4]; gB2 = gp[5];
    }
    double sumA_r = 0.0, sumA_i = 0.0, sumB_r = 0.0, sumB_i = 0.0;  // (SF, SPM > 8) the groups of a mode are added up in order
    const uint32_t q_lo = cut * bpu * gpb, q_hi = min(n_groups, (cut + 1u) * bpu * gpb);
    for (uint32_t q = q_lo; q < q_hi; ++q) {
      const uint32_t v = 8u * q + g;
      const uint32_t b = spm_pow2 ? (v >> spm_shift) : __umulhi(v, spm_magic), k = v - b * SPM;
      const double* Aq = Dd + (size_t)q * KS * 192u + lane;
      double acc[6][2];
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        acc[e][0] = acc[e][1] = 0.0;
        dmma_8x8x4(acc[e][0], acc[e][1], Aq[32 * e], w0);
      }
      if (KS == 2u) {
#pragma unroll
        for (int e = 0; e < 6; ++e) dmma_8x8x4(acc[e][0], acc[e][1], Aq[192 + 32 * e], w1);
      }
      const bool atom = k < NAT && b < mb;
      if (!SF) {
        // ---- finish: rotation, atom permutation, Gamma phase, store (the same operations, in the same order, as the other
        // interpolation kernels: rotate_phase) ------------------------------------------------------------------------------
        if (atom) {
          double2* const row = reinterpret_cast<double2*>(c.vecs_out) + (size_t)(b0 + b) * S;
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            if (x == 0 ? !validA : !validB) continue;
            const uint32_t p = x == 0 ? pA : pB, rr = x == 0 ? rr2.x : rr2.y, ri = x == 0 ? riA : riB;
            const double2 a0 = make_double2(acc[0][x], acc[1][x]), a1 = make_double2(acc[2][x], acc[3][x]), a2 = make_double2(acc[4][x], acc[5][x]);
            double2* out = row + (size_t)(x == 0 ? qi2.x : qi2.y) * wrow;
            uint32_t dest = k;
            if (kind >= 0) {
              double2 ph = make_double2(1.0, 0.0);
              if (gamma) {
                dest = c.F0[k * G + ri];
                ph = PH[(size_t)p * NAT + k];
              }
              out += 3 * dest;
              double2 u0, u1, u2;
              if (x == 0 || riB == riA) {
                rotate_phase(R, a0, a1, a2, ph, gamma, u0, u1, u2);
              } else {  // (the two points of a lane are neighbours in the item, which is sorted by operation: rare)
                double R1[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) R1[i] = c.RS[9u * ri + i];
                rotate_phase(R1, a0, a1, a2, ph, gamma, u0, u1, u2);
              }
              if (kind == 2) {  // axial: det(R) R^-1 v
                const double det = c.rot_det[rr >> 16];
                u0 = make_double2(u0.x * det, u0.y * det);
                u1 = make_double2(u1.x * det, u1.y * det);
                u2 = make_double2(u2.x * det, u2.y * det);
              }
              store48(out, u0, u1, u2);  // (selects, not branches: every lane takes part in both store instructions)
            } else {
              out += 3 * dest;
              store48(out, a0, a1, a2);
            }
          }
        }
      } else {
        // ---- fused structure factor: qv . (R a) = (qv^T R) . a with the row vector g of the point; times the per-(point,
        // source atom) factor; summed over the atoms of the mode: the lanes g of the mode (butterfly), then the groups of the mode
        double FrA = 0.0, FiA = 0.0, FrB = 0.0, FiB = 0.0;
        {
          const double drA = __fma_rn(gA2, acc[4][0], __fma_rn(gA1, acc[2][0], __dmul_rn(gA0, acc[0][0])));
          const double diA = sgn * __fma_rn(gA2, acc[5][0], __fma_rn(gA1, acc[3][0], __dmul_rn(gA0, acc[1][0])));
          const double drB = __fma_rn(gB2, acc[4][1], __fma_rn(gB1, acc[2][1], __dmul_rn(gB0, acc[0][1])));
          const double diB = sgn * __fma_rn(gB2, acc[5][1], __fma_rn(gB1, acc[3][1], __dmul_rn(gB0, acc[1][1])));
          if (atom) {
            const double2 fA = PH[(size_t)pA * NAT + k], fB = PH[(size_t)pB * NAT + k];
            FrA = __fma_rn(-fA.y, diA, __dmul_rn(fA.x, drA));
            FiA = __fma_rn(fA.y, drA, __dmul_rn(fA.x, diA));
            FrB = __fma_rn(-fB.y, diB, __dmul_rn(fB.x, drB));
            FiB = __fma_rn(fB.y, drB, __dmul_rn(fB.x, diB));
          }
        }
#define B200_MMA_SF_STAGE(o_)                                                                                     \
  if (SPM >= 2u * ((o_) >> 2)) {                                                                                  \
    FrA += __shfl_xor_sync(0xffffffffu, FrA, (o_)); FiA += __shfl_xor_sync(0xffffffffu, FiA, (o_));               \
    FrB += __shfl_xor_sync(0xffffffffu, FrB, (o_)); FiB += __shfl_xor_sync(0xffffffffu, FiB, (o_));               \
  }
        B200_MMA_SF_STAGE(4) B200_MMA_SF_STAGE(8) B200_MMA_SF_STAGE(16)
#undef B200_MMA_SF_STAGE
        bool last = true;
        if (SPM > 8u) {
          const uint32_t in_mode = q - b * gpb;   // (SPM multiple of 8: the groups of mode b are [b * gpb, (b + 1) * gpb))
          sumA_r = in_mode == 0u ? FrA : sumA_r + FrA; sumA_i = in_mode == 0u ? FiA : sumA_i + FiA;
          sumB_r = in_mode == 0u ? FrB : sumB_r + FrB; sumB_i = in_mode == 0u ? FiB : sumB_i + FiB;
          FrA = sumA_r; FiA = sumA_i; FrB = sumB_r; FiB = sumB_i;
          last = in_mode + 1u == gpb;
        }
        if (last && k == 0u && b < mb) {
          if (validA) c.sf_out[(size_t)qi2.x * M + b0 + b] = __fma_rn(FiA, FiA, __dmul_rn(FrA, FrA));
          if (validB) c.sf_out[(size_t)qi2.y * M + b0 + b] = __fma_rn(FiB, FiB, __dmul_rn(FrB, FrB));
        }
      }
    }
  }
}

}  // namespace b200
