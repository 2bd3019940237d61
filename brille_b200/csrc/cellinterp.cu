// cellinterp.cu -- the cell-batched interpolation stage (the fast path of stage 2).
//
// Same arithmetic as interp.cu (interpolator_at.tpp:91-127 + interpolator_gamma.tpp:49-139), re-organised around the
// observation that everything expensive in the reference's per-Q loop depends only on the CELL the point fell into:
//   * which 4 (tetrahedron) or 8 (cube) vertex rows are gathered                       trellis_node.hpp:130-149,273-308
//   * the permutation of the modes of every vertex relative to the pivot vertex          interpolatordual.hpp:374-382
//   * the phase e^{-i arg<d_pivot|d_v>} that aligns every vertex' eigenvector to the pivot's   utilities.tpp:567-579
// The locate stage therefore buckets the points by cell (counting sort: k_locate counts, k_bucket_scan scans,
// k_bucket_scatter scatters).  One CTA then stages the cell's vertex rows ONCE in shared memory -- already permuted and
// pre-multiplied by their alignment phase -- and streams up to `chunk` points of that cell through them: per point only
// the barycentric/trilinear weights, the rotation matrix index and the per-atom Gamma phase differ.  The vertex table
// is read from L2 once per (cell, chunk) instead of once per point, and the per-point work drops to
// 2*n_v real-complex FMAs per element plus the 3x3 rotation.
//
// Points whose pivot is not the cell's first corner (some weight ~ 0: on a face/edge/vertex of the cell) and failed
// points go to a last bucket that is processed by the general kernel of interp.cu.
#include "cell_common.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------------
// counting sort, part 2: single-CTA exclusive scan of the bucket populations + work-item table
// ---------------------------------------------------------------------------------------------------------------
// part 2a: population of every bucket (cell) = sum over its sub-buckets (one per point group operation); one thread per
// bucket, spread over many CTAs so that the strided loads overlap
__global__ void __launch_bounds__(256) k_bucket_totals(BucketDev b) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n_buckets) return;
  const uint32_t* c = b.cell_count + (size_t)i * b.sub;
  uint32_t t = 0;
  for (uint32_t r = 0; r < b.sub; ++r) t += c[r];
  b.cell_total[i] = t;
}

// part 2b: single-CTA exclusive scan of the bucket populations + work-item table.  A thread owns SCAN_E consecutive buckets per
// round (serial prefix in registers), the thread totals are scanned over the CTA with shuffles: the 2.6e5 spatial bins of a Nest /
// Mesh regrouping take 32 rounds (with one bucket per thread and round: 256 rounds of five barriers, 0.26 ms).
constexpr int SCAN_E = 8;
__global__ void __launch_bounds__(1024) k_bucket_scan(BucketDev b) {
  __shared__ uint32_t w_pts[32], w_its[32];
  __shared__ uint32_t carry_pts, carry_its;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry_pts = carry_its = 0;
  __syncthreads();
  const uint32_t nb = b.n_buckets, last = nb - 1;
  for (uint32_t base = 0; base < nb; base += 1024 * SCAN_E) {
    const uint32_t i0 = base + (uint32_t)tid * SCAN_E;
    uint32_t c[SCAN_E], it[SCAN_E], tc = 0, ti = 0;
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
      const uint32_t i = i0 + e;
      c[e] = i < nb ? b.cell_total[i] : 0u;
      it[e] = (i < last) ? (c[e] + b.chunk - 1) / b.chunk : 0u;  // the last bucket produces no cell items
      tc += c[e];
      ti += it[e];
    }
    // inclusive scan of the thread totals over the 1024 threads: within warps by shuffles, then over the 32 warp totals
    uint32_t xc = tc, xi = ti;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t yc = __shfl_up_sync(0xffffffffu, xc, o), yi = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) { xc += yc; xi += yi; }
    }
    if (lane == 31) { w_pts[w] = xc; w_its[w] = xi; }
    __syncthreads();
    if (w == 0) {
      uint32_t sc = w_pts[lane], si = w_its[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t yc = __shfl_up_sync(0xffffffffu, sc, o), yi = __shfl_up_sync(0xffffffffu, si, o);
        if (lane >= o) { sc += yc; si += yi; }
      }
      w_pts[lane] = sc;
      w_its[lane] = si;
    }
    __syncthreads();
    uint32_t p0 = carry_pts + xc - tc + (w ? w_pts[w - 1] : 0u), q0 = carry_its + xi - ti + (w ? w_its[w - 1] : 0u);
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
      const uint32_t i = i0 + e;
      if (i < nb) {
        b.cell_start[i] = p0;
        b.item_start[i] = q0;  // the work items themselves are written by k_bucket_items, one thread each
        if (i == last) {
          b.n_items[1] = p0;
          b.n_items[2] = c[e];
        }
      }
      p0 += c[e];
      q0 += it[e];
    }
    __syncthreads();
    if (tid == 0) {
      carry_pts += w_pts[31];
      carry_its += w_its[31];
    }
    __syncthreads();
  }
  if (tid == 0) b.n_items[0] = carry_its;
}

// part 2c: the work items (cell, <= chunk points), one thread per item: its cell is the last one whose first item index is
// not beyond the item (binary search in the exclusive scan of the item counts; empty cells share the index of their successor)
__global__ void __launch_bounds__(256) k_bucket_items(BucketDev b) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= b.n_items[0]) return;
  uint32_t lo = 0, hi = b.n_buckets - 1;  // the last bucket has no items: search [0, n_buckets - 1)
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (b.item_start[mid] <= k) lo = mid; else hi = mid;
  }
  const uint32_t j = k - b.item_start[lo], c = b.cell_total[lo];
  CellItem ci;
  ci.key = lo;
  ci.start = b.cell_start[lo] + j * b.chunk;
  ci.len = min(b.chunk, c - j * b.chunk);
  b.items[k] = ci;
}

// part 2d: first position of every sub-bucket
__global__ void __launch_bounds__(256) k_bucket_suboffsets(BucketDev b) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n_buckets) return;
  const uint32_t* c = b.cell_count + (size_t)i * b.sub;
  uint32_t* o = b.cell_offset + (size_t)i * b.sub;
  uint32_t run = b.cell_start[i];
  for (uint32_t r = 0; r < b.sub; ++r) {
    o[r] = run;
    run += __ldg(c + r);
  }
}

// counting sort, part 3.  Only the 4-byte point index is scattered (the 40 MB `order` array stays in L2); the
// per-point records themselves are gathered by the cell kernel, early enough to be hidden behind the cell staging.
__global__ void __launch_bounds__(256)
k_bucket_scatter(const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ cell_offset,
                 uint32_t* __restrict__ order, size_t n, const uint32_t* __restrict__ index) {
  // entry i of key/rank describes point index[i] (or i itself)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    order[cell_offset[key[i]] + rank[i]] = index ? index[i] : (uint32_t)i;
}

// ---------------------------------------------------------------------------------------------------------------
// the cell kernel
// ---------------------------------------------------------------------------------------------------------------
struct cplx2 { double re, im; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256, 2) k_interp_cell(CellArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (blockIdx.x >= a.bk.n_items[0]) return;
  const CellItem item = a.bk.items[blockIdx.x];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const InterpDev& vals = a.dd.values;
  const InterpDev& vecs = a.dd.vectors;
  const uint32_t M = vecs.branches, S = vecs.span, NAT = vecs.no1, no0v = vals.span, G = a.dd.n_ops;
  const int kind = a.ir ? vecs.rot_kind : -1;
  const bool gamma = kind >= 3;
  const bool is_cube = item.key < a.n_cubes;
  const int NV = is_cube ? 8 : 4;
  const uint32_t cellidx = is_cube ? item.key : item.key - a.n_cubes;
  const uint32_t mpp = a.modes_per_pass;
  const SmemPlan pl = plan_smem(a.n_cubes ? 8u : 4u, mpp, S, no0v, a.bk.chunk, NAT, G, gamma);
  double2* D = reinterpret_cast<double2*>(smem + pl.D);
  double* V = reinterpret_cast<double*>(smem + pl.V);
  double* W = reinterpret_cast<double*>(smem + pl.W);
  double2* PH = reinterpret_cast<double2*>(smem + pl.PH);
  double* RS = reinterpret_cast<double*>(smem + pl.RS);
  uint32_t* F0 = reinterpret_cast<uint32_t*>(smem + pl.F0);
  uint32_t* QI = reinterpret_cast<uint32_t*>(smem + pl.QI);
  uint32_t* RI = reinterpret_cast<uint32_t*>(smem + pl.RI);
  double2* PHI = reinterpret_cast<double2*>(smem + pl.PHI);
  __shared__ uint32_t s_vtx[8], s_prow[8];

  // ---- per-cell constants: vertices in emission order, permutation rows relative to the pivot ------------------
  if (tid < NV) {
    // emission order: cube corner 7-j (trellis_node.hpp:143-147), tetrahedron corner j (:285-287)
    const uint32_t slot = is_cube ? 7u - tid : (uint32_t)tid;
    s_vtx[tid] = is_cube ? a.cube_vertices[(size_t)cellidx * 8 + slot] : a.tet_vertices[(size_t)cellidx * 4 + slot];
    uint32_t prow = 0;
    if (a.dd.n_perm_rows > 1) {
      const uint32_t pivot = is_cube ? 7u : 0u;
      prow = is_cube ? a.dd.cube_perm[(size_t)cellidx * 64 + pivot * 8 + slot] : a.dd.tet_perm[(size_t)cellidx * 16 + pivot * 4 + slot];
    }
    s_prow[tid] = prow;
  }
  // rotation matrices used by this call, and the atom permutation table
  {
    const double* src = nullptr;
    switch (kind) {
      case 0: src = a.dd.rot_int; break;              // R
      case 1: src = a.dd.rot_int + 9 * (size_t)G; break;  // R^T
      case 2: src = a.dd.rot_int; break;              // R^-1 = R[invridx]
      case 3: src = a.dd.rot_int; break;
      case 4: src = a.dd.rot_cart; break;
      default: break;
    }
    if (src)
      for (uint32_t i = tid; i < G * 9; i += nthr) RS[i] = src[i];
    if (gamma)
      for (uint32_t i = tid; i < NAT * G; i += nthr) F0[i] = a.dd.gamma_F0[i];
  }
  // ---- per-point records: one point per thread (chunk <= blockDim) --------------------------------------------------
  // The gathers are issued here and consumed only after the first staging pass below, so their latency (point index ->
  // record, two dependent DRAM/L2 accesses) overlaps the staging and phase alignment of the cell's vertex rows.
  const uint32_t CH = a.bk.chunk;  // row length of the transposed weight array W[corner][point]
  __shared__ uint32_t s_hist[64];
  if (tid < 64) s_hist[tid] = 0;
  for (uint32_t t = tid; t < (uint32_t)NV * CH; t += nthr) W[t] = 0.0;  // padding points carry zero weight
  const bool has = (uint32_t)tid < item.len;
  uint32_t my_q = 0;
  int my_r = 0, my_inv = 0;
  double2 my_w[4];
  double my_qir[3] = {0.0, 0.0, 0.0};
  if (has) {
    my_q = a.bk.order[item.start + tid];
    my_r = a.ridx[my_q];
    my_inv = a.invridx[my_q];
    const double2* wp = reinterpret_cast<const double2*>(a.weight + REC_DOUBLES * (size_t)my_q);
#pragma unroll
    for (int j = 0; j < 4; ++j) my_w[j] = wp[j];
    if (gamma) {
      my_qir[0] = a.q_ir[3 * (size_t)my_q];
      my_qir[1] = a.q_ir[3 * (size_t)my_q + 1];
      my_qir[2] = a.q_ir[3 * (size_t)my_q + 2];
    }
  }

  const size_t vrow = (size_t)M * no0v, wrow = (size_t)M * S;
  for (uint32_t b0 = 0; b0 < M; b0 += mpp) {
    const uint32_t mb = min(mpp, M - b0);
    __syncthreads();  // previous pass finished reading D / V
    // ---- stage the permuted vertex rows ---------------------------------------------------------------------------
    for (uint32_t idx = tid; idx < (uint32_t)NV * mb * S; idx += nthr) {
      const uint32_t i = idx / (mb * S), r = idx - i * (mb * S), b = r / S, e = r - b * S;
      const uint32_t pb = a.dd.n_perm_rows > 1 ? a.dd.perm_rows[(size_t)s_prow[i] * M + b0 + b] : b0 + b;
      D[((size_t)i * mpp + b) * S + e] = reinterpret_cast<const double2*>(vecs.data)[(size_t)s_vtx[i] * wrow + (size_t)pb * S + e];
    }
    for (uint32_t idx = tid; idx < (uint32_t)NV * mb * no0v; idx += nthr) {
      const uint32_t i = idx / (mb * no0v), r = idx - i * (mb * no0v), b = r / no0v, e = r - b * no0v;
      const uint32_t pb = a.dd.n_perm_rows > 1 ? a.dd.perm_rows[(size_t)s_prow[i] * M + b0 + b] : b0 + b;
      V[((size_t)i * mpp + b) * no0v + e] = vals.data[(size_t)s_vtx[i] * vrow + (size_t)pb * no0v + e];
    }
    __syncthreads();
    // ---- align every vertex' eigenvector to the pivot's (utilities.tpp:567-579) -------------------------------------
    // One THREAD per (vertex, mode): z = <d_pivot|d_v>, factor e^{-i arg z} = conj(z)/|z| (the reference evaluates
    // polar(1, -atan2(Im z, Re z)), the same number up to rounding; z == 0 gives 1 in both).  All threads then scale.
    for (uint32_t pr = tid; pr < (uint32_t)(NV - 1) * mb; pr += nthr) {
      const uint32_t i = 1 + pr / mb, b = pr - (i - 1) * mb;
      const double2* d0 = D + (size_t)b * S;  // pivot (emission index 0) keeps its own branch
      const double2* dx = D + ((size_t)i * mpp + b) * S;
      double re = 0.0, im = 0.0;
      for (uint32_t e = 0; e < S; ++e) align_accumulate(d0[e], dx[e], re, im);
      const double2 f = align_factor(re, im);
      PHI[pr] = f;
    }
    __syncthreads();
    for (uint32_t idx = tid; idx < (uint32_t)(NV - 1) * mb * S; idx += nthr) {
      const uint32_t pr = idx / S, e = idx - pr * S, i = 1 + pr / mb, b = pr - (i - 1) * mb;
      const double2 f = PHI[pr];
      double2* dx = D + ((size_t)i * mpp + b) * S + e;
      *dx = align_apply(f, *dx);
    }
    __syncthreads();
    if (b0 == 0) {
      // ---- consume the per-point records --------------------------------------------------------------------------
      // The points of the chunk are re-ordered by the index of the rotation matrix they need (a counting sort over <= 48
      // bins in shared memory), so that the TQ consecutive points a thread works on almost always share the matrix.
      // which matrix multiplies the interpolated vectors: gamma/axial use R^-1, real/recip use R (interpolator_*.tpp)
      // (the empty volatile asm pins the first use of the gathered values here: without it the compiler schedules the
      // trivial arithmetic on them right after the loads, i.e. before the staging that is meant to hide their latency)
      asm volatile("" : "+r"(my_r), "+r"(my_inv), "+r"(my_q));
      const uint32_t mi = (uint32_t)((kind == 0 || kind == 1) ? my_r : my_inv);
      uint32_t my_rank = 0;
      if (has) my_rank = atomicAdd(&s_hist[mi], 1u);
      __syncthreads();
      if (tid == 0) {  // exclusive scan of <= 48 bins
        uint32_t run = 0;
        for (uint32_t j = 0; j < G; ++j) {
          const uint32_t c = s_hist[j];
          s_hist[j] = run;
          run += c;
        }
      }
      __syncthreads();
      if (has) {
        const uint32_t t = s_hist[mi] + my_rank;
        QI[t] = my_q;
        RI[t] = mi | ((uint32_t)my_r << 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (2 * j < NV) {
            W[(size_t)(2 * j) * CH + t] = my_w[j].x;
            W[(size_t)(2 * j + 1) * CH + t] = my_w[j].y;
          }
        }
        if (gamma) {
          // e^{2 pi i q_ir . (R^-1 r_l - r_k)} once per (point, atom)   interpolator_gamma.tpp:18-32,56-58
          for (uint32_t k = 0; k < NAT; ++k) {
            const double* gv = a.dd.gamma_vectors + 3 * (size_t)a.dd.gamma_vidx[(size_t)k * G + mi];
            const double dot = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(my_qir[0], gv[0])), __dmul_rn(my_qir[1], gv[1])), __dmul_rn(my_qir[2], gv[2]));
            double sn, cs;
            sincos(6.283185307179586476925286766559 * dot, &sn, &cs);
            PH[(size_t)t * NAT + k] = make_double2(cs, sn);
          }
        }
      }
      __syncthreads();
    }
    // ---- eigenvalues and eigenvectors of this pass (cell_common.cuh) -----------------------------------------------------
    CellPass cp;
    cp.D = D; cp.V = V; cp.W = W; cp.PH = PH; cp.RS = RS; cp.F0 = F0; cp.QI = QI; cp.RI = RI;
    cp.CH = CH; cp.mpp = mpp; cp.mb = mb; cp.b0 = b0; cp.len = item.len; cp.M = M; cp.S = S; cp.NAT = NAT; cp.no0v = no0v; cp.G = G;
    cp.NV = NV; cp.kind = kind; cp.gamma = gamma; cp.rot_det = a.dd.rot_det; cp.vals_out = a.vals_out; cp.vecs_out = a.vecs_out; cp.task_ctr = nullptr;
    cell_compute_pass<4>(cp, tid, nthr);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
bool cell_path_eligible(const DataDev& dd) {
  const InterpDev& v = dd.values;
  const InterpDev& w = dd.vectors;
  return !v.is_complex && v.no1 == 0 && v.no2 == 0 && v.no0 >= 1 && w.is_complex && w.no0 == 0 && w.no2 == 0 && w.no1 >= 1 &&
         v.rot_kind == -1;
}

// modes staged per pass so that the carve-up fits `budget` bytes of dynamic shared memory
uint32_t cell_modes_per_pass(const DataDev& dd, bool has_cubes, uint32_t chunk, size_t budget) {
  const bool gamma = dd.vectors.rot_kind >= 3;
  for (uint32_t mpp = dd.vectors.branches; mpp >= 1; --mpp)
    if (plan_smem(has_cubes ? 8u : 4u, mpp, dd.vectors.span, dd.values.span, chunk, dd.vectors.no1, dd.n_ops, gamma).total <= budget) return mpp;
  return 0;
}

// Points per work item: the per-point tables (weights, per-atom Gamma phases) grow with chunk * n_atoms, the staged rows with
// modes_per_pass * span.  Take the largest chunk that still leaves room for at least a quarter of the modes per pass.
uint32_t cell_pick_chunk(const DataDev& dd, bool has_cubes, uint32_t preferred, size_t budget, uint32_t* mpp_out) {
  uint32_t best_chunk = 0, best_mpp = 0;
  for (uint32_t chunk = preferred; chunk >= 32; chunk /= 2) {
    const uint32_t mpp = cell_modes_per_pass(dd, has_cubes, chunk, budget);
    if (mpp == 0) continue;
    if (best_chunk == 0) { best_chunk = chunk; best_mpp = mpp; }
    if (4 * mpp >= dd.vectors.branches || mpp == dd.vectors.branches) { best_chunk = chunk; best_mpp = mpp; break; }
    if (mpp > best_mpp) { best_chunk = chunk; best_mpp = mpp; }
  }
  *mpp_out = best_mpp;
  return best_chunk;
}

cudaError_t launch_bucket_sort(const BucketDev& bk, const uint32_t* key, const uint32_t* rank, size_t n, int sm_count,
                               cudaStream_t stream, const uint32_t* index) {
  const unsigned cell_blocks = (bk.n_buckets + 255) / 256;
  k_bucket_totals<<<cell_blocks, 256, 0, stream>>>(bk);
  k_bucket_scan<<<1, 1024, 0, stream>>>(bk);
  if (bk.max_items) k_bucket_items<<<(bk.max_items + 255) / 256, 256, 0, stream>>>(bk);
  k_bucket_suboffsets<<<cell_blocks, 256, 0, stream>>>(bk);
  size_t want = (n + 255) / 256, cap = (size_t)sm_count * 16;
  k_bucket_scatter<<<(int)(want < cap ? want : cap), 256, 0, stream>>>(key, rank, bk.cell_offset, bk.order, n, index);
  return cudaGetLastError();
}

cudaError_t launch_interp_cell(const CellArgs& args, size_t n, cudaStream_t stream) {
  const DataDev& dd = args.dd;
  const bool gamma = (args.ir ? dd.vectors.rot_kind : -1) >= 3;
  const size_t smem = plan_smem(args.n_cubes ? 8u : 4u, args.modes_per_pass, dd.vectors.span, dd.values.span, args.bk.chunk, dd.vectors.no1, dd.n_ops, gamma).total;
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_interp_cell, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const size_t max_items = (n + args.bk.chunk - 1) / args.bk.chunk + (args.bk.n_buckets - 1);
  const unsigned threads = args.bk.chunk <= 128 ? 128 : 256;  // one point per thread in the prologue: chunk <= threads
  k_interp_cell<<<(unsigned)max_items, threads, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace b200
