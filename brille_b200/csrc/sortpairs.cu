// sortpairs.cu -- brille's sort() on the device: mode assignment between connected grid vertices.
//
// DualInterpolator::sort() (interpolatordual.hpp:398-434) visits every connected vertex pair (i < j) of the grid, builds a
// modes x modes cost matrix from the eigenvalues (no phase freedom) and eigenvectors (arbitrary phase allowed)
// (interpolator_cost.tpp:18-58, interpolator.hpp:246-299, utilities.tpp) and solves the linear assignment problem with the
// Jonker-Volgenant algorithm (lapjv.hpp:281-538); the row solution is stored for (i, j), the column solution for (j, i).
// Every pair is independent: a grid has 1e4 - 1e6 pairs.
//
//   k_pair_costs   one thread per cost-matrix entry (pair, mode i, mode j): the anti-phase e^{-i arg<a|b>}, then the scalar,
//                  vector and matrix costs of the two rows, in the reference's operation order (compiled without FMA
//                  contraction; the transcendental functions are CUDA's, not glibc's: costs agree to ~1 ulp)
//   k_pair_match   one warp per pair: the assignment problem with the reference's decisions (ties included), every scan over
//                  the columns lane-parallel, work arrays in shared memory (see match_pair)
//
// Permutations are integers: the tests demand equality with the reference's for every pair (ties of the cost are the only
// place where the 1-ulp difference of the transcendental functions could show; none occurs in the test grids).
#include <algorithm>
#include <cfloat>

#include "device_tables.cuh"
#include "brille_b200.h"

namespace b200 {

struct CostCfg {
  double v_mult[3], w_mult[3];
  int v_vfun, w_vfun;
};

__device__ __forceinline__ bool approx_default(double a, double b) {  // approx_float::scalar with (tol 0, digit 1)
  const double rel = DBL_EPSILON * 10000.0, abs_ = 5.0 / 1000000000000000.0;
  const double x = fabs(a - b);
  return x <= abs_ + rel * fabs(a + b) || x < DBL_MIN;
}
__device__ __forceinline__ double clamp_acos(double c_t) {  // tail of vector_angle / euclidean_angle / hermitian_angle
  double act = fabs(c_t);
  if (approx_default(act, 1.0) && act > 1) {
    c_t /= act;
    act = fabs(c_t);
  }
  if (act > 1) return nan("");  // the reference throws
  return acos(c_t);
}
__device__ __forceinline__ double cos_of(double num, double nA, double nB) {
  if (nA != 0.0 && nB != 0.0) return num / (nA * nB);
  return (nA != 0.0 || nB != 0.0) ? 0.0 : 1.0;
}

// ---- real rows ------------------------------------------------------------------------------------------------------
__device__ double vector_angle_d(uint32_t n, const double* A, const double* B) {
  double AA = 0, BB = 0, AB = 0;
  for (uint32_t i = 0; i < n; ++i) {
    AA += A[i] * A[i];
    BB += B[i] * B[i];
    AB += A[i] * B[i];
  }
  return clamp_acos(cos_of(AB, sqrt(AA), sqrt(BB)));
}
__device__ double vector_distance_d(uint32_t n, const double* a, const double* b) {
  double s = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const double d = a[i] - b[i];
    s += d * d;
  }
  return sqrt(s);
}
__device__ double vectorfun_d(int vfun, uint32_t n, const double* i, const double* j) {
  switch (vfun) {
    case 1: return vector_distance_d(n, i, j);
    case 2: {
      double h = 0;
      for (uint32_t e = 0; e < n; ++e) h += i[e] * j[e];
      return 1 - h;
    }
    case 3:
    case 4: return vector_angle_d(n, i, j);
    default: {
      const double s = sin(vector_angle_d(n, i, j));
      return s * s;
    }
  }
}

// ---- complex rows: a is read from memory, b is a row multiplied by the phase factor f on the fly ---------------------
struct PhasedRow {
  const double2* b;
  double2 f;
  bool phased;
  __device__ __forceinline__ double2 operator[](uint32_t e) const {
    const double2 x = b[e];
    if (!phased) return x;
    return make_double2(f.x * x.x - f.y * x.y, f.x * x.y + f.y * x.x);  // eith * b[e]
  }
};
__device__ double2 hermitian_product_c(uint32_t n, const double2* a, const PhasedRow& b, uint32_t off) {  // sum conj(a) b
  double hr = 0, hi = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const double2 x = a[i], y = b[off + i];
    hr += x.x * y.x - (-x.y) * y.y;
    hi += x.x * y.y + (-x.y) * y.x;
  }
  return make_double2(hr, hi);
}
__device__ double norm2_a(uint32_t n, const double2* a) {  // real(hermitian_product(a, a))
  double hr = 0;
  for (uint32_t i = 0; i < n; ++i) hr += a[i].x * a[i].x - (-a[i].y) * a[i].y;
  return hr;
}
__device__ double norm2_b(uint32_t n, const PhasedRow& b, uint32_t off) {
  double hr = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const double2 y = b[off + i];
    hr += y.x * y.x - (-y.y) * y.y;
  }
  return hr;
}
__device__ double vector_product_c(uint32_t n, const double2* a, const PhasedRow& b, uint32_t off) {
  const double2 h = hermitian_product_c(n, a, b, off);
  return h.x * h.x - h.y * (-h.y);
}
__device__ double hermitian_angle_c(uint32_t n, const double2* A, const PhasedRow& B, uint32_t off) {
  const double nAB = sqrt(vector_product_c(n, A, B, off));
  const double nA = sqrt(norm2_a(n, A));
  const double nB = sqrt(norm2_b(n, B, off));
  return clamp_acos(cos_of(nAB, nA, nB));
}
__device__ double euclidean_angle_c(uint32_t n, const double2* A, const PhasedRow& B, uint32_t off) {
  double AB = 0, nA = 0, nB = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const double2 x = A[i], y = B[off + i];
    AB += x.x * y.x + x.y * y.y;
    nA += x.x * x.x + x.y * x.y;
    nB += y.x * y.x + y.y * y.y;
  }
  return clamp_acos(cos_of(AB, sqrt(nA), sqrt(nB)));
}
__device__ double vector_distance_c(uint32_t n, const double2* a, const PhasedRow& b, uint32_t off) {
  double s = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const double2 y = b[off + i];
    const double dr = a[i].x - y.x, di = a[i].y - y.y;
    s += dr * dr - di * (-di);
  }
  return sqrt(s);
}
__device__ double vectorfun_c(int vfun, uint32_t n, const double2* a, const PhasedRow& b, uint32_t off) {
  switch (vfun) {
    case 1: return vector_distance_c(n, a, b, off);
    case 2: return 1 - vector_product_c(n, a, b, off);
    case 3: return euclidean_angle_c(n, a, b, off);
    case 4: return hermitian_angle_c(n, a, b, off);
    default: {
      const double s = sin(hermitian_angle_c(n, a, b, off));
      return s * s;
    }
  }
}

// Interpolator::add_cost for one entry: `row_i` is mode i of the first vertex, `row_j` mode j of the second (span elements each,
// in global or shared memory)
__device__ double entry_cost(const InterpDev& t, const double* mult, int vfun, const void* row_i, const void* row_j, bool arbitrary_phase) {
  const uint32_t e0 = t.no0, e1 = 3u * t.no1, e2 = 9u * t.no2, s_ = t.span, mo_ = e0 + e1;
  if (s_ == 0) return 0.0;
  double s_cost = 0, v_cost = 0, m_cost = 0;
  if (t.is_complex) {
    const double2* x0i = static_cast<const double2*>(row_i);
    PhasedRow rhs;
    rhs.b = static_cast<const double2*>(row_j);
    rhs.phased = arbitrary_phase;
    rhs.f = make_double2(1.0, 0.0);
    if (arbitrary_phase) {  // antiphase (utilities.tpp:567-579): polar(1, -atan2(Im <a|b>, Re <a|b>))
      double real_dot = 0, imag_dot = 0;
      for (uint32_t e = 0; e < s_; ++e) {
        const double2 a = x0i[e], b = rhs.b[e];
        real_dot += a.x * b.x + a.y * b.y;
        imag_dot += a.x * b.y - a.y * b.x;
      }
      const double th = -1.0 * atan2(imag_dot, real_dot);
      rhs.f = make_double2(cos(th), sin(th));
    }
    if (e0) {
      double s = 0;
      for (uint32_t z = 0; z < e0; ++z) {
        const double2 y = rhs[z];
        const double dr = x0i[z].x - y.x, di = x0i[z].y - y.y;
        s += sqrt(dr * dr - di * (-di));  // magnitude
      }
      s_cost = s;
    }
    if (e1) v_cost = vectorfun_c(vfun, e1, x0i + e0, rhs, e0);
    if (e2)
      for (uint32_t m = 0; m < e2 / 9; ++m) m_cost += vector_distance_c(9, x0i + mo_ + 9u * m, rhs, mo_ + 9u * m);
  } else {
    const double* x0i = static_cast<const double*>(row_i);
    const double* x1j = static_cast<const double*>(row_j);
    if (e0) {
      double s = 0;
      for (uint32_t z = 0; z < e0; ++z) s += fabs(x0i[z] - x1j[z]);
      s_cost = s;
    }
    if (e1) v_cost = vectorfun_d(vfun, e1, x0i + e0, x1j + e0);
    if (e2)
      for (uint32_t m = 0; m < e2 / 9; ++m) m_cost += vector_distance_d(9, x0i + mo_ + 9u * m, x1j + mo_ + 9u * m);
  }
  return mult[0] * s_cost + mult[1] * v_cost + mult[2] * m_cost;
}

__global__ void __launch_bounds__(256) k_pair_costs(InterpDev values, InterpDev vectors, CostCfg cfg, const uint32_t* __restrict__ pairs,
                                                    size_t n_pairs, double* __restrict__ cost) {
  const uint32_t B = vectors.branches;
  const size_t total = n_pairs * B * B;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const size_t p = g / ((size_t)B * B);
    const uint32_t r = (uint32_t)(g - p * B * B), i = r / B, j = r - i * B;
    const uint32_t i0 = pairs[2 * p], i1 = pairs[2 * p + 1];
    double c = 0.0;
    if (i0 == i1) {
      c = i == j ? -1.0 : 0.0;
    } else {
      const size_t ev = values.is_complex ? 16 : 8, ew = vectors.is_complex ? 16 : 8;
      const char* vd = reinterpret_cast<const char*>(values.data);
      const char* wd = reinterpret_cast<const char*>(vectors.data);
      c += entry_cost(values, cfg.v_mult, cfg.v_vfun, vd + ((size_t)i0 * B + i) * values.span * ev, vd + ((size_t)i1 * B + j) * values.span * ev, false);
      c += entry_cost(vectors, cfg.w_mult, cfg.w_vfun, wd + ((size_t)i0 * B + i) * vectors.span * ew, wd + ((size_t)i1 * B + j) * vectors.span * ew, true);
    }
    cost[g] = c;
  }
}

// The same costs with the rows of the two vertices of a pair staged in shared memory: one CTA per pair copies the 2 x modes rows
// of both interpolators once (coalesced), every thread then walks "its" (i, j) entries over rows that sit in shared memory -- in
// k_pair_costs every entry re-reads two rows from global memory several times, and the threads of a warp (consecutive j) read
// them a whole row apart.  Row strides are padded to an odd number of 16-byte units, so that the rows j, j+1, ... that the lanes
// of a warp read fall into different banks.  Same arithmetic, same order: identical costs.
struct StagePlan {
  uint32_t v_stride, w_stride;  // bytes between rows
  uint32_t v0, v1, w0, w1;      // byte offsets of the four row blocks
  uint32_t total;
};
__host__ __device__ inline StagePlan stage_plan(const InterpDev& values, const InterpDev& vectors) {
  auto stride = [](const InterpDev& t) {
    const uint32_t bytes = t.span * (t.is_complex ? 16u : 8u), units = (bytes + 15u) / 16u;
    return (units | 1u) * 16u;
  };
  StagePlan p;
  p.v_stride = stride(values);
  p.w_stride = stride(vectors);
  const uint32_t B = vectors.branches;
  p.v0 = 0;
  p.v1 = p.v0 + B * p.v_stride;
  p.w0 = p.v1 + B * p.v_stride;
  p.w1 = p.w0 + B * p.w_stride;
  p.total = p.w1 + B * p.w_stride;
  return p;
}
__global__ void __launch_bounds__(256) k_pair_costs_staged(InterpDev values, InterpDev vectors, CostCfg cfg, const uint32_t* __restrict__ pairs,
                                                           size_t n_pairs, double* __restrict__ cost, StagePlan pl) {
  extern __shared__ __align__(16) unsigned char pc_smem[];
  const uint32_t B = vectors.branches;
  auto stage = [&](const InterpDev& t, uint32_t vertex, uint32_t at, uint32_t stride) {
    // rows of `vertex` are contiguous in global memory: 8-byte words, coalesced; padded rows in shared memory
    const uint32_t words_per_row = t.span * (t.is_complex ? 2u : 1u);
    const double* src = t.data + (size_t)vertex * B * words_per_row;
    for (uint32_t k = threadIdx.x; k < B * words_per_row; k += blockDim.x) {
      const uint32_t row = k / words_per_row, e = k - row * words_per_row;
      *reinterpret_cast<double*>(pc_smem + at + row * stride + e * 8u) = src[k];
    }
  };
  for (size_t p = blockIdx.x; p < n_pairs; p += gridDim.x) {
    const uint32_t i0 = pairs[2 * p], i1 = pairs[2 * p + 1];
    double* out = cost + p * B * B;
    if (i0 == i1) {
      for (uint32_t r = threadIdx.x; r < B * B; r += blockDim.x) out[r] = (r / B == r % B) ? -1.0 : 0.0;
      continue;
    }
    __syncthreads();  // (the previous pair's readers are done)
    stage(values, i0, pl.v0, pl.v_stride);
    stage(values, i1, pl.v1, pl.v_stride);
    stage(vectors, i0, pl.w0, pl.w_stride);
    stage(vectors, i1, pl.w1, pl.w_stride);
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < B * B; r += blockDim.x) {
      const uint32_t i = r / B, j = r - i * B;
      double c = entry_cost(values, cfg.v_mult, cfg.v_vfun, pc_smem + pl.v0 + i * pl.v_stride, pc_smem + pl.v1 + j * pl.v_stride, false);
      c += entry_cost(vectors, cfg.w_mult, cfg.w_vfun, pc_smem + pl.w0 + i * pl.w_stride, pc_smem + pl.w1 + j * pl.w_stride, true);
      out[r] = c;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Assignment solver: one WARP per vertex pair, work arrays in shared memory.
//
// The reference solves every pair with a sequential Jonker-Volgenant routine (lapjv.hpp:281-538).  Its result depends on the
// order in which ties are met, so a different algorithm (or a different scan order) would give other -- equally cheap --
// permutations.  Here every scan over the columns is replaced by a lane-parallel formulation that provably takes the same
// decisions (oracle/lap_model.py states them in numpy and is checked against the sequential restatement on tie-heavy
// matrices; tests/test_sort_oracle.py):
//
//   column minima   per-column arg-min, lowest row among equals; a row keeps the LARGEST column that chose it (the reference
//                   walks the columns downwards, first claim wins)                                     lapjv.hpp:311-335
//   dual transfer   the reference's tolerant running minimum `if (h < mn + eps) mn = h` is a recurrence, not a minimum: 32
//                   columns at a time, ballot the columns below mn + eps, the first one sets mn, re-ballot behind it
//                                                                                                     lapjv.hpp:340-358
//   row bidding     (umin, j1) = lexicographic minimum of (h_j, j); (usubmin, j2) the same without j1 -- what the scan of
//                   find_umins_plain (lapjv.hpp:74-99) ends with                                       lapjv.hpp:365-410
//   augmenting path a scan moves the prefix-minimum records of d over the to-do list (new minimum, or tie with the running
//                   minimum) to the ready list: found with a warp prefix-min, applied in list order; a relaxation step
//                   flags (improved, ties the minimum, unassigned) per column: the first unassigned tie ends the path, the
//                   ties before it join the ready list in list order                                   lapjv.hpp:416-520
//
// The tolerance eps = sum(cost) / (10000 dim) is summed in storage order with one accumulator, as the reference does
// (lapjv.hpp:305-307): it enters comparisons.
// ---------------------------------------------------------------------------------------------------------------------
struct PairWork {  // per-warp arrays of `dim` entries each, carved from dynamic shared memory
  double* v;       // column duals
  double* d;       // path distances (scratch h_j during the bidding)
  int* rowsol;
  int* colsol;
  int* pred;       // row before a column on the alternating path (row chosen per column during the column minima)
  int* todo;       // column to-do list of a path: [0, low) scanned, [low, up) ready, [up, dim) to do
  int* freerow;    // unassigned rows
  int* claims;     // how many columns chose a row
};
__host__ __device__ inline size_t pair_work_bytes(uint32_t dim) { return (size_t)dim * (2 * sizeof(double) + 6 * sizeof(int)); }

constexpr unsigned FULL = 0xffffffffu;
constexpr int SERIAL_BID_MAX_DIM = 32;
__device__ __forceinline__ void lexmin_reduce(double& val, int& idx) {  // minimum value, lowest index among equals, in every lane
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(FULL, val, o);
    const int oi = __shfl_xor_sync(FULL, idx, o);
    if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
}

__device__ void match_pair(const int dim, const double* __restrict__ C, const PairWork& w, const int lane) {
  if (dim == 1) {
    if (lane == 0) w.rowsol[0] = w.colsol[0] = 0;
    return;
  }
  // ---- tolerance: the entries summed in storage order (every lane runs the same chain on coalesced loads) ---------------
  double total = 0.0;
  for (int base = 0; base < dim * dim; base += 32) {
    const double x = base + lane < dim * dim ? C[base + lane] : 0.0;
    const int cnt = min(32, dim * dim - base);
    for (int k = 0; k < cnt; ++k) total += __shfl_sync(FULL, x, k);
  }
  const double eps = total / (double)(10000 * dim);
  // ---- column minima ------------------------------------------------------------------------------------------------
  for (int j = lane; j < dim; j += 32) { w.claims[j] = 0; w.rowsol[j] = -1; }
  __syncwarp();
  for (int j = lane; j < dim; j += 32) {
    double mn = C[j];
    int at = 0;
    for (int i = 1; i < dim; ++i) {
      const double c = C[(size_t)i * dim + j];
      if (c < mn) { mn = c; at = i; }
    }
    w.v[j] = mn;
    w.pred[j] = at;
    atomicAdd(w.claims + at, 1);
    atomicMax(w.rowsol + at, j);
  }
  __syncwarp();
  for (int j = lane; j < dim; j += 32) {
    const int at = w.pred[j];
    w.colsol[j] = w.rowsol[at] == j ? at : -1;
  }
  __syncwarp();
  // ---- dual transfer (rows in order: later rows see the duals lowered by earlier ones) ----------------------------------
  int nfree = 0;
  for (int i = 0; i < dim; ++i) {
    const int c = w.claims[i];
    if (c == 0) {
      if (lane == 0) w.freerow[nfree] = i;
      ++nfree;
    } else if (c == 1) {
      const int j1 = w.rowsol[i];
      double mn = DBL_MAX;
      for (int base = 0; base < dim; base += 32) {
        const int j = base + lane;
        const bool in = j < dim && j != j1;
        const double h = in ? C[(size_t)i * dim + j] - w.v[j] : 0.0;
        int start = 0;
        for (;;) {
          const unsigned below = __ballot_sync(FULL, in && lane >= start && h < mn + eps);
          if (!below) break;
          const int first = __ffs(below) - 1;
          mn = __shfl_sync(FULL, h, first);
          start = first + 1;
        }
      }
      __syncwarp();
      if (lane == 0) w.v[j1] = w.v[j1] - mn;
      __syncwarp();
    }
  }
  // ---- row bidding: two sweeps over the unassigned rows (the list is consumed and refilled in place) ----------------------
  // The tolerance is small against the costs, so a pair can take thousands of bids, each a dependent step: for few modes a
  // bid done by the whole warp is two 5-stage shuffle reductions for a dozen columns (the slowest pair of a grid then sets the
  // kernel time, whatever the occupancy: C3 6 ms).  Up to SERIAL_BID_MAX_DIM columns one lane walks the row instead -- the same
  // two lexicographic minima, read off a single pass.
  if (dim <= SERIAL_BID_MAX_DIM) {
    if (lane == 0) {
      for (int sweep = 0; sweep < 2; ++sweep) {
        int k = 0;
        const int n_todo = nfree;
        nfree = 0;
        while (k < n_todo) {
          const int i = w.freerow[k++];
          const double* ci = C + (size_t)i * dim;
          double umin = ci[0] - w.v[0], usub = DBL_MAX;
          int j1 = 0, j2 = dim;
          for (int j = 1; j < dim; ++j) {  // (value, index) minimum and runner-up in one pass, lowest index among equals
            const double h = ci[j] - w.v[j];
            if (h < umin) { usub = umin; j2 = j1; umin = h; j1 = j; }
            else if (h < usub) { usub = h; j2 = j; }
          }
          int i0 = w.colsol[j1];
          const double vj = w.v[j1], lowered = vj - (usub + eps - umin);
          const bool lowers = lowered < vj;
          if (lowers) {
            w.v[j1] = lowered;
          } else if (i0 != -1) {
            j1 = j2;
            i0 = w.colsol[j2];
          }
          w.rowsol[i] = j1;
          w.colsol[j1] = i;
          if (i0 != -1) {
            if (lowers) w.freerow[--k] = i0;
            else w.freerow[nfree++] = i0;
          }
        }
      }
    }
    nfree = __shfl_sync(FULL, nfree, 0);
    __syncwarp();
  } else
  for (int sweep = 0; sweep < 2; ++sweep) {
    int k = 0;
    const int n_todo = nfree;
    nfree = 0;
    __syncwarp();
    while (k < n_todo) {
      const int i = w.freerow[k++];
      double umin = DBL_MAX;
      int j1 = dim;
      for (int j = lane; j < dim; j += 32) {  // (ascending per lane: the first of equal values stays)
        const double h = C[(size_t)i * dim + j] - w.v[j];
        w.d[j] = h;
        if (h < umin) { umin = h; j1 = j; }
      }
      lexmin_reduce(umin, j1);
      double usub = DBL_MAX;
      int j2 = dim;
      for (int j = lane; j < dim; j += 32) {
        const double h = w.d[j];  // (written by this lane)
        if (j != j1 && h < usub) { usub = h; j2 = j; }
      }
      lexmin_reduce(usub, j2);
      int i0 = w.colsol[j1];
      const double vj = w.v[j1];
      const double lowered = vj - (usub + eps - umin);
      const bool lowers = lowered < vj;
      __syncwarp();
      if (!lowers && i0 != -1) {  // the dual cannot move and the column is taken: bid for the second best instead
        j1 = j2;
        i0 = w.colsol[j2];
      }
      __syncwarp();
      if (lane == 0) {
        if (lowers) w.v[j1] = lowered;
        w.rowsol[i] = j1;
        w.colsol[j1] = i;
        if (i0 != -1) {
          if (lowers) w.freerow[k - 1] = i0;  // the displaced row bids next
          else w.freerow[nfree] = i0;
        }
      }
      if (i0 != -1) {
        if (lowers) --k; else ++nfree;
      }
      __syncwarp();
    }
  }
  // ---- one shortest augmenting path per row that is still unassigned --------------------------------------------------
  const double INF = __longlong_as_double(0x7ff0000000000000ll);
  for (int f = 0; f < nfree; ++f) {
    const int start_row = w.freerow[f];
    for (int j = lane; j < dim; j += 32) {
      w.d[j] = C[(size_t)start_row * dim + j] - w.v[j];
      w.pred[j] = start_row;
      w.todo[j] = j;
    }
    __syncwarp();
    int low = 0, up = 0, last = 0, end = -1;
    double mn = 0.0;
    while (end < 0) {
      if (up == low) {
        // scan: the columns of the to-do list whose distance is a new minimum, or ties the running minimum, in list order
        last = low - 1;
        double run = INF;
        for (int base = low; base < dim; base += 32) {
          const int k = base + lane;
          const bool in = k < dim;
          const int j = in ? w.todo[k] : 0;
          const double h = in ? w.d[j] : INF;
          double pm = h;  // inclusive prefix minimum over the lanes
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(FULL, pm, o);
            if (lane >= o) pm = fmin(pm, t);
          }
          double before = __shfl_up_sync(FULL, pm, 1);
          if (lane == 0) before = INF;
          before = fmin(before, run);
          unsigned record = __ballot_sync(FULL, in && h <= before);
          const unsigned fresh = __ballot_sync(FULL, in && h < before);
          while (record) {
            const int p = __ffs(record) - 1;
            record &= record - 1u;
            const int jj = __shfl_sync(FULL, j, p);
            if ((fresh >> p) & 1u) {
              up = low;
              mn = __shfl_sync(FULL, h, p);
            }
            if (lane == 0) {
              w.todo[base + p] = w.todo[up];
              w.todo[up] = jj;
            }
            ++up;
          }
          run = fmin(run, __shfl_sync(FULL, pm, 31));
        }
        __syncwarp();
        for (int base = low; base < up && end < 0; base += 32) {  // an unassigned column among the ready ones ends the path
          const int k = base + lane;
          const int j = k < up ? w.todo[k] : 0;
          const unsigned open = __ballot_sync(FULL, k < up && w.colsol[j] == -1);
          if (open) end = __shfl_sync(FULL, j, __ffs(open) - 1);
        }
        if (end >= 0) break;
      }
      // relax through the row of the next ready column
      const int j1 = w.todo[low++];
      const int i = w.colsol[j1];
      const double h1 = C[(size_t)i * dim + j1] - w.v[j1] - mn;
      const int first = up;
      for (int base = first; base < dim && end < 0; base += 32) {
        const int k = base + lane;
        const bool in = k < dim;
        const int j = in ? w.todo[k] : 0;
        const double v2 = in ? C[(size_t)i * dim + j] - w.v[j] - h1 : 0.0;
        const bool better = in && v2 < w.d[j];
        const bool tie = better && v2 == mn;
        const unsigned closing = __ballot_sync(FULL, tie && w.colsol[j] == -1);
        const int stop = closing ? __ffs(closing) - 1 : 32;
        if (better && lane <= stop) w.pred[j] = i;
        if (better && lane < stop) w.d[j] = v2;
        unsigned joins = __ballot_sync(FULL, tie) & (stop >= 32 ? FULL : ((1u << stop) - 1u));
        while (joins) {
          const int p = __ffs(joins) - 1;
          joins &= joins - 1u;
          const int jj = __shfl_sync(FULL, j, p);
          if (lane == 0) {
            w.todo[base + p] = w.todo[up];
            w.todo[up] = jj;
          }
          ++up;
        }
        if (closing) end = __shfl_sync(FULL, j, stop);
      }
      __syncwarp();
    }
    // duals of the scanned columns, then flip the assignments along the path back to the starting row
    for (int k = lane; k <= last; k += 32) {
      const int j = w.todo[k];
      w.v[j] = w.v[j] + w.d[j] - mn;
    }
    __syncwarp();
    if (lane == 0) {
      int at = end, i;
      do {
        i = w.pred[at];
        w.colsol[at] = i;
        const int was = w.rowsol[i];
        w.rowsol[i] = at;
        at = was;
      } while (i != start_row);
    }
    __syncwarp();
  }
}

// bytes of one warp's slice of shared memory: work arrays, then (small problems) the cost matrix itself
__host__ __device__ inline size_t pair_slice_bytes(uint32_t dim, bool stage_cost) {
  return (pair_work_bytes(dim) + 15) / 16 * 16 + (stage_cost ? (size_t)dim * dim * sizeof(double) : 0);
}
constexpr uint32_t STAGE_COST_MAX_DIM = 40;  // cost matrices up to 40 x 40 (12.8 KB) are copied into shared memory first

__global__ void __launch_bounds__(256) k_pair_match(uint32_t B, size_t n_pairs, const double* __restrict__ cost, int* __restrict__ row,
                                                    int* __restrict__ col) {
  extern __shared__ __align__(16) unsigned char pm_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const bool stage_cost = B <= STAGE_COST_MAX_DIM;
  unsigned char* base = pm_smem + (size_t)wid * pair_slice_bytes(B, stage_cost);
  double* const staged = reinterpret_cast<double*>(base + (pair_work_bytes(B) + 15) / 16 * 16);
  PairWork w;
  w.v = reinterpret_cast<double*>(base);
  w.d = w.v + B;
  w.rowsol = reinterpret_cast<int*>(w.d + B);
  w.colsol = w.rowsol + B;
  w.pred = w.colsol + B;
  w.todo = w.pred + B;
  w.freerow = w.todo + B;
  w.claims = w.freerow + B;
  for (size_t p = (size_t)blockIdx.x * wpc + wid; p < n_pairs; p += (size_t)gridDim.x * wpc) {
    const double* C = cost + p * B * B;
    if (stage_cost) {  // the solver reads the matrix many times, a row or a column at a time
      for (uint32_t k = lane; k < B * B; k += 32) staged[k] = C[k];
      __syncwarp();
      C = staged;
    }
    match_pair((int)B, C, w, lane);
    __syncwarp();
    for (uint32_t j = lane; j < B; j += 32) {
      row[p * B + j] = w.rowsol[j];
      col[p * B + j] = w.colsol[j];
    }
    __syncwarp();
  }
}

// warps per CTA and dynamic shared memory of k_pair_match for `B` modes (0: too many modes for one warp's work arrays)
static int pair_match_config(uint32_t B, size_t* smem) {
  const size_t per_warp = pair_slice_bytes(B, B <= STAGE_COST_MAX_DIM);
  int warps = 8;
  while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
  if (per_warp * warps > 200 * 1024) return 0;
  *smem = per_warp * warps;
  return warps;
}
static cudaError_t launch_pair_match(uint32_t B, size_t n, const double* cost, int* row, int* col, int sm_count) {
  size_t smem = 0;
  const int warps = pair_match_config(B, &smem);
  if (!warps) return cudaErrorInvalidConfiguration;
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_pair_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const size_t want = (n + warps - 1) / warps, cap = (size_t)sm_count * 16;
  k_pair_match<<<(unsigned)(want < cap ? want : cap), warps * 32, smem>>>(B, n, cost, row, col);
  return cudaGetLastError();
}

// the solver on its own: cost matrices from the host, permutations back (diagnostic entry point b200_solve_assignments)
cudaError_t run_match_only(const double* h_cost, size_t n, uint32_t B, int32_t* h_row, int32_t* h_col, int sm_count) {
  if (n == 0 || B == 0) return cudaSuccess;
  double* dc = nullptr;
  int *dr = nullptr, *dl = nullptr;
  cudaError_t e = cudaMalloc(&dc, n * B * B * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dr, n * B * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&dl, n * B * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(dc, h_cost, n * B * B * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_pair_match(B, n, dc, dr, dl, sm_count);
  if (e == cudaSuccess) e = cudaMemcpy(h_row, dr, n * B * sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(h_col, dl, n * B * sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(dc); cudaFree(dr); cudaFree(dl);
  return e;
}

// host side: pairs are processed in batches so that the cost matrices stay within `max_ws_bytes`; the work space is kept
// by the caller between calls (cudaMalloc / cudaFree cost more than the kernels)
void SortWorkspace::release() {
  cudaFree(pairs); cudaFree(cost); cudaFree(row); cudaFree(col);
  pairs = nullptr; cost = nullptr; row = col = nullptr;
  batch = 0; branches = 0;
}
cudaError_t SortWorkspace::ensure(size_t n, uint32_t B) {
  if (n <= batch && B == branches) return cudaSuccess;
  release();
  cudaError_t e;
  if ((e = cudaMalloc(&pairs, n * 2 * sizeof(uint32_t))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&cost, n * B * B * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&row, n * B * sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&col, n * B * sizeof(int))) != cudaSuccess) return e;
  batch = n;
  branches = B;
  return cudaSuccess;
}

cudaError_t run_sort_pairs(const DataDev& dd, const double v_mult[3], int v_vfun, const double w_mult[3], int w_vfun, const uint32_t* h_pairs,
                           size_t n_pairs, int32_t* h_row, int32_t* h_col, double* h_cost, int sm_count, size_t max_ws_bytes,
                           SortWorkspace& ws, uint64_t* launches) {
  const uint32_t B = dd.vectors.branches;
  if (n_pairs == 0 || B == 0) return cudaSuccess;
  CostCfg cfg;
  for (int i = 0; i < 3; ++i) { cfg.v_mult[i] = v_mult[i]; cfg.w_mult[i] = w_mult[i]; }
  cfg.v_vfun = v_vfun;
  cfg.w_vfun = w_vfun;
  const size_t per_pair = (size_t)B * B * 8 + (size_t)B * (2 * 4) + 8;
  size_t batch = max_ws_bytes / per_pair;
  if (batch < 1) batch = 1;
  if (batch > n_pairs) batch = n_pairs;
  cudaError_t e = ws.ensure(batch, B);
  if (e != cudaSuccess) { ws.release(); return e; }
  batch = ws.batch;
  for (size_t lo = 0; lo < n_pairs; lo += batch) {
    const size_t n = n_pairs - lo < batch ? n_pairs - lo : batch;
    if ((e = cudaMemcpy(ws.pairs, h_pairs + 2 * lo, n * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    const size_t entries = n * B * B, want = (entries + 255) / 256, cap = (size_t)sm_count * 32;
    const StagePlan pl = stage_plan(dd.values, dd.vectors);
    if (pl.total <= 200u * 1024u) {  // rows of both vertices fit in shared memory: one CTA per pair
      static size_t configured_dev[MAX_DEVICES] = {};
      size_t& configured = configured_dev[current_device_slot()];
      if (pl.total > 48u * 1024u && pl.total > configured) {
        if ((e = cudaFuncSetAttribute(k_pair_costs_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total)) != cudaSuccess) return e;
        configured = pl.total;
      }
      const unsigned threads = (unsigned)std::min<size_t>(256, ((size_t)B * B + 31) / 32 * 32);
      const size_t cap_c = (size_t)sm_count * 8;
      k_pair_costs_staged<<<(unsigned)(n < cap_c ? n : cap_c), threads, pl.total>>>(dd.values, dd.vectors, cfg, ws.pairs, n, ws.cost, pl);
    } else {
      k_pair_costs<<<(unsigned)(want < cap ? want : cap), 256>>>(dd.values, dd.vectors, cfg, ws.pairs, n, ws.cost);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = launch_pair_match(B, n, ws.cost, ws.row, ws.col, sm_count)) != cudaSuccess) return e;
    if (launches) *launches += 2;
    if ((e = cudaMemcpy(h_row + lo * B, ws.row, n * B * sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(h_col + lo * B, ws.col, n * B * sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
    if (h_cost && (e = cudaMemcpy(h_cost + lo * B * B, ws.cost, n * B * B * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace b200
