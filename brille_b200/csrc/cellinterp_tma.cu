// cellinterp_tma.cu -- the cell-batched interpolation stage, persistent and software-pipelined (the default fast path).
//
// Same arithmetic as cellinterp.cu (interpolator_at.tpp:91-127 + interpolator_gamma.tpp:49-139).  Two things change:
//
//   1. CELL TABLE.  Everything the on-the-fly kernel derives while staging a cell -- which vertex rows, the permutation
//      of their modes relative to the pivot (interpolatordual.hpp:374-382) and the phase e^{-i arg<d_pivot|d_v>} that
//      aligns every vertex' eigenvector to the pivot's (utilities.tpp:567-579) -- depends only on (cell, fill data).
//      k_build_cell_table evaluates it ONCE per fill() into one contiguous record per cell and per pass of modes:
//          [ D: NV x mpp x S complex | V: NV x mpp x no0v double ]
//      (NV = 8 cube / 4 tetrahedron, emission order of trellis_node.hpp:143-147,285-287).  A record is exactly what
//      the compute pass wants in shared memory, so staging a cell becomes one bulk asynchronous copy (TMA,
//      cp.async.bulk + mbarrier) issued by one thread.
//
//   2. PERSISTENT CTAs, ONE BARRIER PER ITEM.  Each CTA walks its share of the work items (cell, <= chunk points)
//      produced by the counting sort, which orders the points by (cell, point group operation).  While item n is
//      computed, the record of the next cell is already in flight into the other half of a double buffer (TMA); every
//      thread then turns the raw record (96 bytes in a 128-byte line) of "its" point of item n+1 (fetched with a per-point bulk copy issued
//      one item earlier, through the sort order loaded another item earlier) into the transposed weights, Gamma phases
//      and indices of item n+1 in the other half of a second double buffer.  Nothing in that step needs another
//      thread, so warps drift freely between the two steps and the trigonometry of one warp overlaps the stores of
//      another; the single __syncthreads per item publishes the buffers.  Consecutive items of the same cell reuse
//      the staged record.  Rotation tables and the Gamma vectors are loaded once per CTA.
#include "cell_common.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, bulk copy (TMA without tensor map)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// cell table
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t cell_tile_bytes(uint32_t nv, uint32_t mpp, uint32_t S, uint32_t no0v) {
  return (size_t)nv * mpp * ((size_t)S * 16 + (size_t)no0v * 8);  // nv is 4 or 8: always a multiple of 16
}
__host__ __device__ inline size_t cell_record_offset(const CellTableDev& ct, uint32_t key) {
  return key < ct.n_cubes ? (size_t)key * ct.cube_bytes : (size_t)ct.n_cubes * ct.cube_bytes + (size_t)(key - ct.n_cubes) * ct.tet_bytes;
}

// one thread per (cell, emitted vertex, padded mode)
__global__ void __launch_bounds__(256) k_build_cell_table(DataDev dd, const uint32_t* __restrict__ cube_vertices,
                                                          const uint32_t* __restrict__ tet_vertices, CellTableDev ct,
                                                          unsigned char* __restrict__ table) {
  const InterpDev& vals = dd.values;
  const InterpDev& vecs = dd.vectors;
  const uint32_t M = vecs.branches, S = vecs.span, no0v = vals.span, mpp = ct.mpp;
  const uint32_t padded = ct.n_pass * mpp;
  const size_t n_cube_rows = (size_t)ct.n_cubes * 8 * padded, n_rows = n_cube_rows + (size_t)ct.n_tets * 4 * padded;
  const size_t vrow = (size_t)M * no0v, wrow = (size_t)M * S;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_rows; g += (size_t)gridDim.x * blockDim.x) {
    const bool is_cube = g < n_cube_rows;
    const uint32_t NV = is_cube ? 8u : 4u;
    const size_t gl = is_cube ? g : g - n_cube_rows;
    const uint32_t cellidx = (uint32_t)(gl / ((size_t)NV * padded));
    const uint32_t rem = (uint32_t)(gl - (size_t)cellidx * NV * padded);
    const uint32_t i = rem / padded, bb = rem - i * padded, pass = bb / mpp, bl = bb - pass * mpp;
    const uint32_t key = is_cube ? cellidx : ct.n_cubes + cellidx;
    unsigned char* rec = table + cell_record_offset(ct, key) + (size_t)pass * cell_tile_bytes(NV, mpp, S, no0v);
    double2* Drow = reinterpret_cast<double2*>(rec) + ((size_t)i * mpp + bl) * S;
    double* Vrow = reinterpret_cast<double*>(rec + (size_t)NV * mpp * S * 16) + ((size_t)i * mpp + bl) * no0v;
    if (bb >= M) {  // padding of the last pass
      for (uint32_t e = 0; e < S; ++e) Drow[e] = make_double2(0.0, 0.0);
      for (uint32_t e = 0; e < no0v; ++e) Vrow[e] = 0.0;
      continue;
    }
    // emission order: cube corner 7-j (trellis_node.hpp:143-147), tetrahedron corner j (:285-287); pivot = first emitted
    const uint32_t slot = is_cube ? 7u - i : i, pslot = is_cube ? 7u : 0u;
    const uint32_t* cv = is_cube ? cube_vertices + (size_t)cellidx * 8 : tet_vertices + (size_t)cellidx * 4;
    const uint32_t v = cv[slot], v0 = cv[pslot];
    uint32_t pb = bb, pb0 = bb;
    if (dd.n_perm_rows > 1) {
      const uint32_t* pt = is_cube ? dd.cube_perm + (size_t)cellidx * 64 + pslot * 8 : dd.tet_perm + (size_t)cellidx * 16 + pslot * 4;
      pb = dd.perm_rows[(size_t)pt[slot] * M + bb];
      pb0 = dd.perm_rows[(size_t)pt[pslot] * M + bb];
    }
    const double2* src = reinterpret_cast<const double2*>(vecs.data) + (size_t)v * wrow + (size_t)pb * S;
    if (i == 0) {
      for (uint32_t e = 0; e < S; ++e) Drow[e] = src[e];
    } else {
      const double2* piv = reinterpret_cast<const double2*>(vecs.data) + (size_t)v0 * wrow + (size_t)pb0 * S;
      double re = 0.0, im = 0.0;
      for (uint32_t e = 0; e < S; ++e) align_accumulate(piv[e], src[e], re, im);
      const double2 f = align_factor(re, im);
      for (uint32_t e = 0; e < S; ++e) Drow[e] = align_apply(f, src[e]);
    }
    const double* vsrc = vals.data + (size_t)v * vrow + (size_t)pb * no0v;
    for (uint32_t e = 0; e < no0v; ++e) Vrow[e] = vsrc[e];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// shared memory plan of the pipelined kernel
// ---------------------------------------------------------------------------------------------------------------
struct TmaPlan {
  uint32_t D0, D1, RW, W, PH, RS, F0, GV, QI, RI, QV, BAR, per_set, total;
};
__host__ __device__ inline TmaPlan plan_smem_tma(uint32_t nvmax, uint32_t mpp, uint32_t S, uint32_t no0v, uint32_t chunk,
                                                 uint32_t n_at, uint32_t G, bool gamma, bool sf = false) {
  TmaPlan p;
  uint32_t o = 0;
  auto take = [&](size_t bytes) { uint32_t at = o; o += (uint32_t)((bytes + 15) / 16 * 16); return at; };
  const size_t tile = cell_tile_bytes(nvmax, mpp, S, no0v);
  p.D0 = take(tile);
  p.D1 = take(tile);
  p.RW = take((size_t)chunk * REC_USED_BYTES);  // raw per-point records of the next item (per-point bulk copies), packed: a 128-byte
                                                 // stride in shared memory would put the same field of every record in one bank
  // two sets of per-item tables (W, PH, QI, RI): the set of item n is read while the set of item n+1 is written
  const uint32_t set0 = o;
  p.W = take((size_t)nvmax * chunk * 8);
  p.PH = take(gamma ? (size_t)chunk * n_at * 16 : 0);
  p.QI = take((size_t)chunk * 4);
  p.RI = take((size_t)chunk * 4);
  p.QV = take(sf ? (size_t)chunk * 24 : 0);  // fused structure factor: g = (T Q)^T R of every point
  p.per_set = o - set0;
  o += p.per_set;
  p.RS = take((size_t)G * 9 * 8);
  p.F0 = take(gamma ? (size_t)n_at * G * 4 : 0);
  p.GV = take(gamma ? (size_t)n_at * G * 24 : 0);
  p.BAR = take(32);
  p.total = o;
  return p;
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t NO_ITEM = 0xffffffffu;
constexpr uint32_t ITEM_BLOCK = 8;  // (4 and 16 measured: 5.19 / 5.27 ms against 5.17)

// SF: fused structure-factor finish (cell_sf_pass): |F|^2 per (Q, mode) instead of the eigenvectors (a.sf_out, a.Q, a.sf)
template <int TQ, bool SF>
__global__ void __launch_bounds__(256, TQ == 2 ? 3 : 2) k_interp_cell_tma(const __grid_constant__ CellArgs a, const __grid_constant__ CellTableDev ct,
                                                            const unsigned char* __restrict__ table, const __grid_constant__ TmaPlan pl) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const InterpDev& vals = a.dd.values;
  const InterpDev& vecs = a.dd.vectors;
  const uint32_t M = vecs.branches, S = vecs.span, NAT = vecs.no1, no0v = vals.span, G = a.dd.n_ops;
  const int kind = a.ir ? vecs.rot_kind : -1;
  const bool gamma = kind >= 3;
  const uint32_t mpp = ct.mpp, n_pass = ct.n_pass, CH = a.bk.chunk;
  unsigned char* const D0p = smem + pl.D0;
  unsigned char* const D1p = smem + pl.D1;
  double* const RW = reinterpret_cast<double*>(smem + pl.RW);
  double* const RS = reinterpret_cast<double*>(smem + pl.RS);
  uint32_t* const F0 = reinterpret_cast<uint32_t*>(smem + pl.F0);
  double* const GV = reinterpret_cast<double*>(smem + pl.GV);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + pl.BAR);
  __shared__ uint32_t s_task_ctr[2];

  // ---- this CTA's work items: blocks of ITEM_BLOCK consecutive items (same-cell reuse), dealt round-robin to the CTAs
  // (the item list is ordered cubes first, then tetrahedra: a cyclic deal gives every CTA the same mix of both) ----------
  const uint32_t n_items = a.bk.n_items[0];
  if (blockIdx.x * ITEM_BLOCK >= n_items) return;
  // local item l -> global item ((l / ITEM_BLOCK) * gridDim.x + blockIdx.x) * ITEM_BLOCK + l % ITEM_BLOCK (monotonic in l)
#define GLOBAL_ITEM(l_) ((((l_) / ITEM_BLOCK) * gridDim.x + blockIdx.x) * ITEM_BLOCK + ((l_) % ITEM_BLOCK))

  // ---- per-CTA constants ---------------------------------------------------------------------------------------------
  {
    const double* src = nullptr;
    switch (kind) {
      case 0: src = a.dd.rot_int; break;                  // R
      case 1: src = a.dd.rot_int + 9 * (size_t)G; break;  // R^T
      case 2: src = a.dd.rot_int; break;                  // R^-1 = R[invridx]
      case 3: src = a.dd.rot_int; break;
      case 4: src = a.dd.rot_cart; break;
      default: break;
    }
    if (src)
      for (uint32_t i = tid; i < G * 9; i += nthr) RS[i] = src[i];
    if (gamma) {
      for (uint32_t i = tid; i < NAT * G; i += nthr) {
        F0[i] = a.dd.gamma_F0[i];
        const double* gv = a.dd.gamma_vectors + 3 * (size_t)a.dd.gamma_vidx[i];
        GV[3 * i] = gv[0];
        GV[3 * i + 1] = gv[1];
        GV[3 * i + 2] = gv[2];
      }
    }
    if (tid == 0) {
      s_task_ctr[0] = s_task_ctr[1] = 0u;
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      mbar_init(&bar[2], 1);
      mbar_init(&bar[3], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();

  uint32_t parity = 0u;  // bit b = phase parity of bar[b]
  // an item is (key, start, len); scalars, so that the descriptors stay in registers
#define LOAD_ITEM(it_, key_, start_, len_)                                \
  do {                                                                    \
    key_ = NO_ITEM; start_ = 0; len_ = 0;                                 \
    const uint32_t gi_ = GLOBAL_ITEM(it_);                                \
    if (gi_ < n_items) {                                                  \
      const CellItem* ci_ = a.bk.items + gi_;                             \
      key_ = ci_->key; start_ = ci_->start; len_ = ci_->len;              \
    }                                                                     \
  } while (0)
  auto issue_tile = [&](uint32_t key, uint32_t pass, int b) {  // one thread
    const uint32_t nv = key < a.n_cubes ? 8u : 4u;
    const uint32_t bytes = (uint32_t)cell_tile_bytes(nv, mpp, S, no0v);
    const unsigned char* src = table + cell_record_offset(ct, key) + (size_t)pass * bytes;
    mbar_expect_tx(bar + b, bytes);
    bulk_g2s(b ? D1p : D0p, src, bytes, bar + b);
  };
  // raw records of an item: every thread fetches the record (REC_USED_BYTES of a 128-byte line) of its own point with one bulk copy; all of them
  // complete on bar[2 + (item & 1)], on which thread 0 announces the total.  Two barriers, because threads issue the copies
  // of item n+2 right after they have seen item n+1 complete, without a CTA barrier in between: with a single mbarrier a
  // copy of item n+2 could be counted in the phase of item n+1, and a late thread could miss a whole phase.
  auto issue_raw = [&](uint32_t q, uint32_t len, uint32_t item) {
    uint64_t* const rb = bar + 2 + (item & 1u);
    if (tid == 0 && len) mbar_expect_tx(rb, len * REC_USED_BYTES);
    if ((uint32_t)tid < len) bulk_g2s(RW + REC_SMEM_DOUBLES * (size_t)tid, a.weight + REC_DOUBLES * (size_t)q, REC_USED_BYTES, rb);
  };
  auto wait_raw = [&](uint32_t item) {  // every thread, exactly once per item
    const uint32_t b = 2u + (item & 1u);
    mbar_wait(bar + b, (parity >> b) & 1u);
    parity ^= 1u << b;
  };
  // this thread's point of an item: raw record -> transposed weights, Gamma phases, indices in table set `set`.
  // The sort already ordered the points of a cell by operation: position in the item == position in the tables.
  auto make_tables = [&](uint32_t len, int nv, int set, double qx, double qy, double qz) {
    unsigned char* const base = smem + (set ? pl.per_set : 0);
    double* const W = reinterpret_cast<double*>(base + pl.W);
    double2* const PH = reinterpret_cast<double2*>(base + pl.PH);
    uint32_t* const QI = reinterpret_cast<uint32_t*>(base + pl.QI);
    uint32_t* const RI = reinterpret_cast<uint32_t*>(base + pl.RI);
    if (SF) {
      // Fused structure factor.  A point is shared by nthr / CH threads (t = tid % CH; all of them hold its input point in
      // registers, loaded an item ahead), which split its atoms.  Per SOURCE atom k (destination l = F0(k, R)) the combined factor
      //     coef_l e^{-qv.W_l.qv} e^{2 pi i (Q.r_l -/+ q_ir.(R^-1 r_l - r_k))}
      // -- the Gamma phase of interpolator_gamma.tpp:18-32,56-58 (conjugated when the eigenvectors are) and the atom's own phase
      // share one sincos -- and, once per point, the row vector g = qv^T R, so that the finish is a plain dot product g . a.
      const uint32_t t = (uint32_t)tid % CH, part = (uint32_t)tid / CH, nrep = (uint32_t)nthr / CH;
      if (t < len) {
        const double* rec = RW + REC_SMEM_DOUBLES * (size_t)t;
        const uint32_t rot = *reinterpret_cast<const uint32_t*>(rec + 11);
        const uint32_t mi = (kind == 0 || kind == 1) ? (rot & 0xffffu) : (rot >> 16);
        const double* T = a.sf.T;
        const double v0 = T[0] * qx + T[1] * qy + T[2] * qz, v1 = T[3] * qx + T[4] * qy + T[5] * qz, v2 = T[6] * qx + T[7] * qy + T[8] * qz;
        if (part == 0) {
          double* const QV = reinterpret_cast<double*>(base + pl.QV);
          const double* R = RS + 9 * mi;
          QV[3 * t] = __fma_rn(v2, R[6], __fma_rn(v1, R[3], __dmul_rn(v0, R[0])));
          QV[3 * t + 1] = __fma_rn(v2, R[7], __fma_rn(v1, R[4], __dmul_rn(v0, R[1])));
          QV[3 * t + 2] = __fma_rn(v2, R[8], __fma_rn(v1, R[5], __dmul_rn(v0, R[2])));
        }
        for (uint32_t k = part; k < NAT; k += nrep) {
          const uint32_t l = F0[k * G + mi];
          const double* gv = GV + 3 * ((size_t)k * G + mi);
          const double gdot = rec[8] * gv[0] + rec[9] * gv[1] + rec[10] * gv[2];
          double arg = a.sf.conjugate ? -gdot : gdot;
          if (a.sf.pos) {
            const double* r = a.sf.pos + 3 * l;
            arg += qx * r[0] + qy * r[1] + qz * r[2];
          }
          double sn, cs;
          sincos(6.283185307179586476925286766559 * arg, &sn, &cs);
          const double cr = a.sf.coef[2 * l], ci = a.sf.coef[2 * l + 1];
          double2 f = make_double2(cr * cs - ci * sn, cr * sn + ci * cs);
          if (a.sf.dw) {
            const double* Wl = a.sf.dw + 9 * l;
            const double w = v0 * (Wl[0] * v0 + Wl[1] * v1 + Wl[2] * v2) + v1 * (Wl[3] * v0 + Wl[4] * v1 + Wl[5] * v2) +
                             v2 * (Wl[6] * v0 + Wl[7] * v1 + Wl[8] * v2);
            const double e = exp(-w);
            f.x *= e;
            f.y *= e;
          }
          PH[(size_t)t * NAT + k] = f;
        }
      }
    }
    if ((uint32_t)tid < len) {
      const double* rec = RW + REC_SMEM_DOUBLES * (size_t)tid;  // weight[8] | q_ir[3] | rot, index
      const uint2 ri2 = *reinterpret_cast<const uint2*>(rec + 11);
      const uint32_t my_r = ri2.x & 0xffffu, my_inv = ri2.x >> 16;
      // which matrix multiplies the interpolated vectors: gamma/axial use R^-1, real/recip use R (interpolator_*.tpp)
      const uint32_t mi = (kind == 0 || kind == 1) ? my_r : my_inv;
      QI[tid] = ri2.y;
      RI[tid] = mi | (my_r << 16);
      const double2* rw = reinterpret_cast<const double2*>(rec);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (2 * j < nv) {
          const double2 w2 = rw[j];
          W[(size_t)(2 * j) * CH + tid] = w2.x;
          W[(size_t)(2 * j + 1) * CH + tid] = w2.y;
        }
      }
    } else if ((uint32_t)tid < ((len + 3u) & ~3u)) {  // padding points of the last register tile carry zero weight
      for (int i = 0; i < nv; ++i) W[(size_t)i * CH + tid] = 0.0;
      QI[tid] = 0;
      RI[tid] = 0;
    }
    if (gamma && !SF) {
      // e^{2 pi i q_ir . (R^-1 r_l - r_k)} once per (point, atom)   interpolator_gamma.tpp:18-32,56-58.  The (point, atom)
      // pairs are dealt to ALL threads (every record of the item is visible once its mbarrier phase completed), so that
      // short items do not leave most warps idle here.
      const uint32_t nat_magic = 0xffffffffu / NAT + 1u;  // floor(u / NAT) == umulhi(u, magic) for u * NAT < 2^32
      for (uint32_t u = tid; u < len * NAT; u += nthr) {
        const uint32_t t = NAT == 1u ? u : __umulhi(u, nat_magic), k = u - t * NAT;
        const double* rec = RW + REC_SMEM_DOUBLES * (size_t)t;
        const uint32_t rot = *reinterpret_cast<const uint32_t*>(rec + 11);
        const uint32_t mi = (kind == 0 || kind == 1) ? (rot & 0xffffu) : (rot >> 16);
        const double* gv = GV + 3 * ((size_t)k * G + mi);
        const double dot = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(rec[8], gv[0])), __dmul_rn(rec[9], gv[1])), __dmul_rn(rec[10], gv[2]));
        double sn, cs;
        sincos(6.283185307179586476925286766559 * dot, &sn, &cs);
        PH[u] = make_double2(cs, sn);
      }
    }
  };

  // Pipeline (item n is being computed):
  //   raw records of item n+1: issued at the top of iteration n (the raw buffer is free: its readers finished before the
  //                            barrier that ended iteration n-1), consumed at the bottom of iteration n
  //   point indices (sort order) of item n+2: loaded into a register at the bottom of iteration n
  //   descriptor of item n+3: loaded at the end of iteration n
  // so that no global load waits on the one before it.
  uint32_t cur_key, cur_start, cur_len, nxt_key, nxt_start, nxt_len, nn_key, nn_start, nn_len;
  LOAD_ITEM(0u, cur_key, cur_start, cur_len);
  LOAD_ITEM(1u, nxt_key, nxt_start, nxt_len);
  LOAD_ITEM(2u, nn_key, nn_start, nn_len);
  int buf = 0, set = 0;
  bool fresh = true;
  uint32_t n_done = 0;  // passes done by this CTA (selects the task counter)
  if (tid == 0) issue_tile(cur_key, 0, 0);
  // prologue: tables of the first item
  double qx = 0.0, qy = 0.0, qz = 0.0;  // (SF) the input point of this thread's point of the next item, loaded an item ahead
  {
    // (SF: thread tid also serves point tid % CH, see make_tables; only the threads tid < len issue the record copies)
    const uint32_t tp = SF ? (uint32_t)tid % CH : (uint32_t)tid;
    const uint32_t q0 = tp < cur_len ? a.bk.order[cur_start + tp] : 0u;
    issue_raw(q0, cur_len, 0u);
    if (SF && tp < cur_len) { qx = a.Q[3 * (size_t)q0]; qy = a.Q[3 * (size_t)q0 + 1]; qz = a.Q[3 * (size_t)q0 + 2]; }
  }
  const uint32_t tp = SF ? (uint32_t)tid % CH : (uint32_t)tid;
  uint32_t q_nx = tp < nxt_len ? a.bk.order[nxt_start + tp] : 0u;
  wait_raw(0u);
  make_tables(cur_len, cur_key < a.n_cubes ? 8 : 4, 0, qx, qy, qz);
  __syncthreads();
  (void)nxt_start;

  for (uint32_t it = 0; cur_key != NO_ITEM; ++it) {
    const int NV = cur_key < a.n_cubes ? 8 : 4;
    unsigned char* const tbase = smem + (set ? pl.per_set : 0);
    if (nxt_key != NO_ITEM) {
      issue_raw(q_nx, nxt_len, it + 1);
      if (SF && tp < nxt_len) { qx = a.Q[3 * (size_t)q_nx]; qy = a.Q[3 * (size_t)q_nx + 1]; qz = a.Q[3 * (size_t)q_nx + 2]; }
    }
    for (uint32_t pass = 0; pass < n_pass; ++pass) {
      const uint32_t b0 = pass * mpp, mb = min(mpp, M - b0);
      // ---- prefetch the next tile into the other buffer (its last readers finished before the previous barrier) ------
      uint32_t nkey = NO_ITEM, npass = 0;
      if (pass + 1 < n_pass) { nkey = cur_key; npass = pass + 1; }
      else if (nxt_key != NO_ITEM) { nkey = nxt_key; npass = 0; }
      const bool reuse_next = n_pass == 1 && nkey == cur_key;
      const bool load_next = nkey != NO_ITEM && !reuse_next;
      if (load_next && tid == 0) issue_tile(nkey, npass, buf ^ 1);
      if (fresh) {
        mbar_wait(bar + buf, (parity >> buf) & 1u);
        parity ^= 1u << buf;
      }
      // ---- eigenvalues and eigenvectors of this pass (cell_common.cuh) ---------------------------------------------------
      CellPass cp;
      unsigned char* const Dcur = buf ? D1p : D0p;
      cp.D = reinterpret_cast<const double2*>(Dcur);
      cp.V = reinterpret_cast<const double*>(Dcur + (size_t)NV * mpp * S * 16);
      cp.W = reinterpret_cast<const double*>(tbase + pl.W);
      cp.PH = reinterpret_cast<const double2*>(tbase + pl.PH);
      cp.QI = reinterpret_cast<const uint32_t*>(tbase + pl.QI);
      cp.RI = reinterpret_cast<const uint32_t*>(tbase + pl.RI);
      cp.RS = RS; cp.F0 = F0;
      cp.CH = CH; cp.mpp = mpp; cp.mb = mb; cp.b0 = b0; cp.len = cur_len; cp.M = M; cp.S = S; cp.NAT = NAT; cp.no0v = no0v; cp.G = G;
      cp.NV = NV; cp.kind = kind; cp.gamma = gamma; cp.rot_det = a.dd.rot_det; cp.vals_out = a.vals_out; cp.vecs_out = a.vecs_out;
      cp.task_ctr = s_task_ctr + (n_done & 1u);
      if (tid == 0) s_task_ctr[(n_done + 1u) & 1u] = 0u;  // the counter of the next pass: idle since the previous barrier
      if (SF) {
        cp.QV = reinterpret_cast<const double*>(tbase + pl.QV);
        cp.sf_out = a.sf_out;
        cp.conjugate = a.sf.conjugate;
        cell_sf_pass<TQ>(cp, tid, nthr);
      } else {
        cell_compute_pass<TQ>(cp, tid, nthr);
      }
      if (pass + 1 == n_pass) {
        if (nxt_key != NO_ITEM) {  // tables of item n+1 into the other set
          wait_raw(it + 1);
          make_tables(nxt_len, nxt_key < a.n_cubes ? 8 : 4, set ^ 1, qx, qy, qz);
        }
        q_nx = tp < nn_len ? a.bk.order[nn_start + tp] : 0u;
      }
      __syncthreads();  // every reader of this tile, this table set and the raw records is done; the other set is complete
      if (load_next) buf ^= 1;
      fresh = load_next;
      ++n_done;
    }
    cur_key = nxt_key; cur_len = nxt_len;
    nxt_key = nn_key; nxt_len = nn_len;
    LOAD_ITEM(it + 3, nn_key, nn_start, nn_len);
    set ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// diagnostic: the output stores of the kernel above on their own -- same grid, same deal of the work items, same rows, same
// 48-byte pieces per lane in the same order, constants instead of results, nothing else (no staging, no tables, no barriers).
// What it reaches is the ceiling of the store pattern; the difference to the real kernel is what the rest of the kernel costs.
// (option "replay_stores"; profiles/README.md)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_store_replay(const __grid_constant__ CellArgs a) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const InterpDev& vecs = a.dd.vectors;
  const uint32_t M = vecs.branches, S = vecs.span, NAT = vecs.no1;
  const size_t wrow = (size_t)M * S;
  const uint32_t n_items = a.bk.n_items[0];
  const uint32_t per_q = M * NAT;
  for (uint32_t l = 0;; ++l) {
    const uint32_t gi = GLOBAL_ITEM(l);
    if (gi >= n_items) break;
    const CellItem it = a.bk.items[gi];
    const uint32_t ntile = (it.len + 3) / 4;
    for (uint32_t task = tid; task < ntile * per_q; task += nthr) {
      const uint32_t tile = task / per_q, r = task - tile * per_q, t0 = tile * 4;
      double2* const out_base = reinterpret_cast<double2*>(a.vecs_out) + 3 * (size_t)r;
      const double2 c0 = make_double2((double)task, 1.0), c1 = make_double2(2.0, (double)l);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t0 + t >= it.len) break;
        const uint32_t q = __ldg(a.bk.order + it.start + t0 + t);
        store48(out_base + (size_t)q * wrow, c0, c1, c0);
      }
    }
  }
}

cudaError_t launch_store_replay(const CellArgs& args, size_t n, int sm_count, cudaStream_t stream) {
  const size_t max_items = (n + args.bk.chunk - 1) / args.bk.chunk + (args.bk.n_buckets - 1);
  size_t grid = (size_t)sm_count * 2;
  const size_t max_blocks = (max_items + ITEM_BLOCK - 1) / ITEM_BLOCK;
  if (grid > max_blocks) grid = max_blocks;
  if (grid == 0) return cudaSuccess;
  k_store_replay<<<(unsigned)grid, 256, 0, stream>>>(args);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// modes per pass / points per item of the pipelined kernel for `budget` bytes of dynamic shared memory
bool cell_sf_fusable(const DataDev& dd, const SFDev& sf) {
  const InterpDev& v = dd.vectors;
  const uint32_t n = sf.n_atoms;
  return v.rot_kind >= 3 && v.is_complex && v.no0 == 0 && v.no2 == 0 && v.no1 == n && n >= 1 && n <= 32;
}

uint32_t cell_tma_pick(const DataDev& dd, bool has_cubes, uint32_t preferred, size_t budget, uint32_t* mpp_out, bool sf) {
  const bool gamma = dd.vectors.rot_kind >= 3;
  const uint32_t M = dd.vectors.branches;
  auto modes_that_fit = [&](uint32_t chunk) {
    for (uint32_t m = M; m >= 1; --m)
      if (plan_smem_tma(has_cubes ? 8u : 4u, m, dd.vectors.span, dd.values.span, chunk, dd.vectors.no1, dd.n_ops, gamma, sf).total <= budget) return m;
    return 0u;
  };
  uint32_t best_chunk = 0, best_mpp = 0;
  // all modes in one pass with a decent number of points per item, if that fits ...
  for (uint32_t chunk = preferred; chunk >= 64 && !best_chunk; chunk /= 2)
    if (modes_that_fit(chunk) == M) { best_chunk = chunk; best_mpp = M; }
  // ... otherwise shrink the chunk until at most 8 passes are needed (else: fewest passes)
  for (uint32_t chunk = preferred; chunk >= 32 && !(best_mpp && 8 * best_mpp >= M); chunk /= 2) {
    const uint32_t mpp = modes_that_fit(chunk);
    if (mpp > best_mpp) { best_chunk = chunk; best_mpp = mpp; }
  }
  if (best_mpp) {  // equalise the passes (same count, smaller padding)
    const uint32_t n_pass = (M + best_mpp - 1) / best_mpp;
    best_mpp = (M + n_pass - 1) / n_pass;
  }
  *mpp_out = best_mpp;
  return best_chunk;
}

CellTableDev cell_table_layout(const DataDev& dd, uint32_t n_cubes, uint32_t n_tets, uint32_t mpp) {
  CellTableDev ct{};
  ct.n_cubes = n_cubes;
  ct.n_tets = n_tets;
  ct.mpp = mpp;
  ct.n_pass = (dd.vectors.branches + mpp - 1) / mpp;
  ct.cube_bytes = (uint64_t)ct.n_pass * cell_tile_bytes(8, mpp, dd.vectors.span, dd.values.span);
  ct.tet_bytes = (uint64_t)ct.n_pass * cell_tile_bytes(4, mpp, dd.vectors.span, dd.values.span);
  ct.total_bytes = (uint64_t)n_cubes * ct.cube_bytes + (uint64_t)n_tets * ct.tet_bytes;
  return ct;
}

cudaError_t launch_build_cell_table(const DataDev& dd, const uint32_t* cube_vertices, const uint32_t* tet_vertices,
                                    const CellTableDev& ct, unsigned char* table, int sm_count, cudaStream_t stream) {
  const size_t rows = ((size_t)ct.n_cubes * 8 + (size_t)ct.n_tets * 4) * ct.n_pass * ct.mpp;
  if (rows == 0) return cudaSuccess;
  const size_t want = (rows + 255) / 256, cap = (size_t)sm_count * 32;
  k_build_cell_table<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(dd, cube_vertices, tet_vertices, ct, table);
  return cudaGetLastError();
}

template <int TQ, bool SF>
static cudaError_t launch_tma_tile(const CellArgs& args, const CellTableDev& ct, const unsigned char* table, size_t n, int sm_count,
                                   const TmaPlan& plan, cudaStream_t stream) {
  const size_t smem = plan.total;
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_interp_cell_tma<TQ, SF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const size_t max_items = (n + args.bk.chunk - 1) / args.bk.chunk + (args.bk.n_buckets - 1);
  static int ctas_dev[MAX_DEVICES] = {};
  static size_t occ_smem_dev[MAX_DEVICES] = {};
  int& ctas_per_sm = ctas_dev[current_device_slot()];
  size_t& occ_smem = occ_smem_dev[current_device_slot()];
  if (occ_smem != smem) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_interp_cell_tma<TQ, SF>, 256, smem);
    if (e != cudaSuccess) return e;
    if (ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
    occ_smem = smem;
  }
  size_t grid = (size_t)sm_count * (size_t)ctas_per_sm;
  const size_t max_blocks = (max_items + ITEM_BLOCK - 1) / ITEM_BLOCK;
  if (grid > max_blocks) grid = max_blocks;
  if (grid == 0) return cudaSuccess;
  k_interp_cell_tma<TQ, SF><<<(unsigned)grid, 256, smem, stream>>>(args, ct, table, plan);
  return cudaGetLastError();
}

cudaError_t launch_interp_cell_tma(const CellArgs& args, const CellTableDev& ct, const unsigned char* table, size_t n,
                                   int sm_count, cudaStream_t stream, int tile) {
  const DataDev& dd = args.dd;
  const bool gamma = (args.ir ? dd.vectors.rot_kind : -1) >= 3;
  const bool sf = args.sf_out != nullptr;
  const TmaPlan plan = plan_smem_tma(args.n_cubes ? 8u : 4u, ct.mpp, dd.vectors.span, dd.values.span, args.bk.chunk, dd.vectors.no1, dd.n_ops, gamma, sf);
  if (sf) return tile == 2 ? launch_tma_tile<2, true>(args, ct, table, n, sm_count, plan, stream) : launch_tma_tile<4, true>(args, ct, table, n, sm_count, plan, stream);
  if (tile == 2) return launch_tma_tile<2, false>(args, ct, table, n, sm_count, plan, stream);
  return launch_tma_tile<4, false>(args, ct, table, n, sm_count, plan, stream);
}

#undef LOAD_ITEM
#undef GLOBAL_ITEM
}  // namespace b200
