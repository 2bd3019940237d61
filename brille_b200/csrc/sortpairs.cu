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
//   k_pair_assign  one thread per pair: the Jonker-Volgenant solver, same control flow as the reference (idx = int)
//
// Permutations are integers: the tests demand equality with the reference's for every pair (ties of the cost are the only
// place where the 1-ulp difference of the transcendental functions could show; none occurs in the test grids).
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

// Interpolator::add_cost for one entry (mode i of vertex i0, mode j of vertex i1)
__device__ double entry_cost(const InterpDev& t, const double* mult, int vfun, uint32_t i0, uint32_t i1, uint32_t i, uint32_t j,
                             bool arbitrary_phase) {
  const uint32_t e0 = t.no0, e1 = 3u * t.no1, e2 = 9u * t.no2, s_ = t.span, B = t.branches, mo_ = e0 + e1;
  if (s_ == 0) return 0.0;
  double s_cost = 0, v_cost = 0, m_cost = 0;
  if (t.is_complex) {
    const double2* x0i = reinterpret_cast<const double2*>(t.data) + ((size_t)i0 * B + i) * s_;
    PhasedRow rhs;
    rhs.b = reinterpret_cast<const double2*>(t.data) + ((size_t)i1 * B + j) * s_;
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
    const double* x0i = t.data + ((size_t)i0 * B + i) * s_;
    const double* x1j = t.data + ((size_t)i1 * B + j) * s_;
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
      c += entry_cost(values, cfg.v_mult, cfg.v_vfun, i0, i1, i, j, false);
      c += entry_cost(vectors, cfg.w_mult, cfg.w_vfun, i0, i1, i, j, true);
    }
    cost[g] = c;
  }
}

// lapjv (lapjv.hpp:281-538) for one pair; work arrays of `dim` entries each live in global memory
__device__ void lapjv_one(int dim, const double* assign_cost, int* rowsol, int* colsol, double* v, double* d, int* freerows,
                          int* collist, int* matches, int* pred) {
  if (1 == dim) {
    rowsol[0] = colsol[0] = 0;
    return;
  }
  for (int i = 0; i < dim; i++) matches[i] = 0;
  double total_cost = 0;
  for (int tc = 0; tc < dim * dim; ++tc) total_cost += assign_cost[tc];
  const double cost_epsilon = total_cost / (double)(10000 * dim);
  // COLUMN REDUCTION
  for (int j = dim; j-- > 0;) {
    double mn = assign_cost[j];
    int imin = 0;
    for (int i = 1; i < dim; i++) {
      const double c = assign_cost[i * dim + j];
      if (c < mn) {
        mn = c;
        imin = i;
      }
    }
    v[j] = mn;
    if (++matches[imin] == 1) {
      rowsol[imin] = j;
      colsol[j] = imin;
    } else {
      colsol[j] = -1;
    }
  }
  // REDUCTION TRANSFER
  int numfree = 0;
  for (int i = 0; i < dim; i++) {
    const double* local_cost = assign_cost + i * dim;
    if (matches[i] == 0) {
      freerows[numfree++] = i;
    } else if (matches[i] == 1) {
      const int j1 = rowsol[i];
      double mn = DBL_MAX;
      for (int j = 0; j < dim; j++)
        if (j != j1)
          if (local_cost[j] - v[j] < mn + cost_epsilon) mn = local_cost[j] - v[j];
      v[j1] = v[j1] - mn;
    }
  }
  // AUGMENTING ROW REDUCTION
  for (int loopcnt = 0; loopcnt < 2; loopcnt++) {
    int k = 0;
    const int prevnumfree = numfree;
    numfree = 0;
    while (k < prevnumfree) {
      const int i = freerows[k++];
      const double* local_cost = assign_cost + i * dim;  // find_umins_plain (lapjv.hpp:74-99)
      double umin = local_cost[0] - v[0];
      int j1 = 0, j2 = -1;
      double usubmin = DBL_MAX;
      for (int j = 1; j < dim; j++) {
        const double h = local_cost[j] - v[j];
        if (h < usubmin) {
          if (h >= umin) {
            usubmin = h;
            j2 = j;
          } else {
            usubmin = umin;
            umin = h;
            j2 = j1;
            j1 = j;
          }
        }
      }
      int i0 = colsol[j1];
      const double vj1_new = v[j1] - (usubmin + cost_epsilon - umin);
      const bool vj1_lowers = vj1_new < v[j1];
      if (vj1_lowers) {
        v[j1] = vj1_new;
      } else if (i0 != -1) {
        j1 = j2;
        i0 = colsol[j2];
      }
      rowsol[i] = j1;
      colsol[j1] = i;
      if (i0 != -1) {
        if (vj1_lowers) freerows[--k] = i0;
        else freerows[numfree++] = i0;
      }
    }
  }
  // AUGMENT SOLUTION for each free row
  for (int f = 0; f < numfree; f++) {
    int endofpath = 0;
    const int freerow = freerows[f];
    for (int j = 0; j < dim; j++) {
      d[j] = assign_cost[freerow * dim + j] - v[j];
      pred[j] = freerow;
      collist[j] = j;
    }
    int low = 0, up = 0;
    bool unassigned_found = false;
    int last = 0;
    double mn = 0;
    do {
      if (up == low) {
        last = low - 1;
        mn = d[collist[up++]];
        for (int k = up; k < dim; k++) {
          const int j = collist[k];
          const double h = d[j];
          if (h <= mn) {
            if (h < mn) {
              up = low;
              mn = h;
            }
            collist[k] = collist[up];
            collist[up++] = j;
          }
        }
        for (int k = low; k < up; k++)
          if (colsol[collist[k]] == -1) {
            endofpath = collist[k];
            unassigned_found = true;
            break;
          }
      }
      if (!unassigned_found) {
        const int j1 = collist[low];
        low++;
        const int i = colsol[j1];
        const double* local_cost = assign_cost + i * dim;
        const double h = local_cost[j1] - v[j1] - mn;
        for (int k = up; k < dim; k++) {
          const int j = collist[k];
          const double v2 = local_cost[j] - v[j] - h;
          if (v2 < d[j]) {
            pred[j] = i;
            if (v2 == mn) {
              if (colsol[j] == -1) {
                endofpath = j;
                unassigned_found = true;
                break;
              } else {
                collist[k] = collist[up];
                collist[up++] = j;
              }
            }
            d[j] = v2;
          }
        }
      }
    } while (!unassigned_found);
    for (int k = 0; k <= last; k++) {
      const int j1 = collist[k];
      v[j1] = v[j1] + d[j1] - mn;
    }
    int i;
    do {
      i = pred[endofpath];
      colsol[endofpath] = i;
      const int j1 = endofpath;
      endofpath = rowsol[i];
      rowsol[i] = j1;
    } while (i != freerow);
  }
}

__global__ void __launch_bounds__(128) k_pair_assign(uint32_t B, size_t n_pairs, const double* __restrict__ cost, int* __restrict__ row,
                                                     int* __restrict__ col, double* __restrict__ fwork, int* __restrict__ iwork) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (size_t)gridDim.x * blockDim.x) {
    double* fw = fwork + p * 2 * B;  // v, d
    int* iw = iwork + p * 4 * B;     // free rows, column list, matches, predecessors
    lapjv_one((int)B, cost + p * B * B, row + p * B, col + p * B, fw, fw + B, iw, iw + B, iw + 2 * B, iw + 3 * B);
  }
}

// host side: pairs are processed in batches so that the cost matrices stay within `max_ws_bytes`; the work space is kept
// by the caller between calls (cudaMalloc / cudaFree cost more than the kernels)
void SortWorkspace::release() {
  cudaFree(pairs); cudaFree(cost); cudaFree(fwork); cudaFree(row); cudaFree(col); cudaFree(iwork);
  pairs = nullptr; cost = fwork = nullptr; row = col = iwork = nullptr;
  batch = 0; branches = 0;
}
cudaError_t SortWorkspace::ensure(size_t n, uint32_t B) {
  if (n <= batch && B == branches) return cudaSuccess;
  release();
  cudaError_t e;
  if ((e = cudaMalloc(&pairs, n * 2 * sizeof(uint32_t))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&cost, n * B * B * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&fwork, n * 2 * B * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&iwork, n * 4 * B * sizeof(int))) != cudaSuccess) return e;
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
  const size_t per_pair = (size_t)B * B * 8 + (size_t)B * (2 * 4 + 2 * 8 + 4 * 4) + 8;
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
    k_pair_costs<<<(unsigned)(want < cap ? want : cap), 256>>>(dd.values, dd.vectors, cfg, ws.pairs, n, ws.cost);
    const size_t want2 = (n + 127) / 128;
    k_pair_assign<<<(unsigned)(want2 < cap ? want2 : cap), 128>>>(B, n, ws.cost, ws.row, ws.col, ws.fwork, ws.iwork);
    if (launches) *launches += 2;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cudaMemcpy(h_row + lo * B, ws.row, n * B * sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(h_col + lo * B, ws.col, n * B * sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
    if (h_cost && (e = cudaMemcpy(h_cost + lo * B * B, ws.cost, n * B * B * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace b200
