/* sort_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * Plain-C restatement of brille's DualInterpolator::sort() for one vertex pair: the mode-assignment cost matrix and the
 * Jonker-Volgenant linear assignment that turns it into a pair of permutations.  Follows, line by line,
 *   interpolatordual.hpp:384-396   cost_matrix (values without, vectors with an arbitrary phase)
 *   interpolatordual.hpp:423-433   determine_permutation_ij (row -> (i,j), col -> (j,i))
 *   interpolator_cost.tpp:18-58    Interpolator::add_cost
 *   interpolator.hpp:246-299       the scalar / vector cost functions selected by set_cost_info
 *   utilities.tpp:249-312,354-381  vector_angle / euclidean_angle / hermitian_angle
 *   utilities.tpp:383-431          vector_distance, vector_product, magnitude
 *   utilities.tpp:567-593          antiphase / inplace_antiphase
 *   permutation.hpp:600-609        jv_permutation_fill
 *   lapjv.hpp:74-99,281-538        find_umins_plain, lapjv (idx = int, cost = double, the non-AVX2 path)
 * Compiled with -ffp-contract=off (the reference is x86-64 baseline code).  Pinned against the reference's own sort() in
 * tests/test_sort_oracle.py.  Real-valued eigenvectors with an arbitrary phase ("dd" classes) are not restated: the
 * reference reads an uninitialised buffer there (utilities.tpp:530 is an empty function).                              */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cx;

static int approx_default(double a, double b) { /* approx_float::scalar(a, b) with (tol = 0, digit = 1) */
  const double rel = DBL_EPSILON * 10000.0, abs_ = 5.0 / 1000000000000000.0;
  double x = fabs(a - b);
  return x <= abs_ + rel * fabs(a + b) || x < DBL_MIN;
}
static double clamp_acos(double c_t) { /* the tail shared by vector_angle / euclidean_angle / hermitian_angle */
  double act = fabs(c_t);
  if (approx_default(act, 1.0) && act > 1) {
    c_t /= act;
    act = fabs(c_t);
  }
  if (act > 1) return NAN; /* the reference throws */
  return acos(c_t);
}
static double cos_of(double num, double nA, double nB) {
  if (nA && nB) return num / (nA * nB);
  return (nA || nB) ? 0.0 : 1.0;
}

/* ---- real data -------------------------------------------------------------------------------------------------- */
static double vector_angle_d(uint32_t n, const double* A, const double* B) {
  double AA = 0, BB = 0, AB = 0;
  for (uint32_t i = 0; i < n; ++i) {
    AA += A[i] * A[i];
    BB += B[i] * B[i];
    AB += A[i] * B[i];
  }
  return clamp_acos(cos_of(AB, sqrt(AA), sqrt(BB)));
}
static double vector_distance_d(uint32_t n, const double* a, const double* b) {
  double s = 0;
  for (uint32_t i = 0; i < n; ++i) {
    double d = a[i] - b[i];
    s += d * d;
  }
  return sqrt(s);
}
static double vector_product_d(uint32_t n, const double* a, const double* b) {
  double h = 0;
  for (uint32_t i = 0; i < n; ++i) h += a[i] * b[i];
  return h;
}
/* ---- complex data ----------------------------------------------------------------------------------------------- */
static cx hermitian_product_c(uint32_t n, const cx* a, const cx* b) { /* sum conj(a) * b */
  double hr = 0, hi = 0;
  for (uint32_t i = 0; i < n; ++i) {
    hr += a[i].re * b[i].re - (-a[i].im) * b[i].im;
    hi += a[i].re * b[i].im + (-a[i].im) * b[i].re;
  }
  cx h = {hr, hi};
  return h;
}
static double vector_product_c(uint32_t n, const cx* a, const cx* b) { /* real(h * conj(h)) */
  cx h = hermitian_product_c(n, a, b);
  return h.re * h.re - h.im * (-h.im);
}
static double hermitian_angle_c(uint32_t n, const cx* A, const cx* B) {
  double nAB = sqrt(vector_product_c(n, A, B));
  double nA = sqrt(hermitian_product_c(n, A, A).re);
  double nB = sqrt(hermitian_product_c(n, B, B).re);
  return clamp_acos(cos_of(nAB, nA, nB));
}
static double euclidean_angle_c(uint32_t n, const cx* A, const cx* B) {
  double AB = 0, nA = 0, nB = 0;
  for (uint32_t i = 0; i < n; ++i) {
    AB += A[i].re * B[i].re + A[i].im * B[i].im;
    nA += A[i].re * A[i].re + A[i].im * A[i].im;
    nB += B[i].re * B[i].re + B[i].im * B[i].im;
  }
  return clamp_acos(cos_of(AB, sqrt(nA), sqrt(nB)));
}
static double vector_distance_c(uint32_t n, const cx* a, const cx* b) {
  double s = 0;
  for (uint32_t i = 0; i < n; ++i) {
    double dr = a[i].re - b[i].re, di = a[i].im - b[i].im;
    s += dr * dr - di * (-di); /* real(d * conj(d)) */
  }
  return sqrt(s);
}
static double magnitude_c(cx a) { return sqrt(a.re * a.re - a.im * (-a.im)); }

typedef struct {
  const double* data; /* (n_vertices, branches * span) of double or complex */
  int is_complex;
  uint32_t el[3];     /* scalars, vector elements, matrix elements per branch */
  double costmult[3];
  int vfun;           /* 0 sin^2(hermitian angle), 1 distance, 2 1 - product, 3 vector angle, 4 hermitian angle */
} interp_t;

static double vectorfun_d(int vfun, uint32_t n, const double* i, const double* j) {
  switch (vfun) {
    case 1: return vector_distance_d(n, i, j);
    case 2: return 1 - vector_product_d(n, i, j);
    case 3: return vector_angle_d(n, i, j);
    case 4: return vector_angle_d(n, i, j); /* hermitian_angle of real data is vector_angle */
    default: { double s = sin(vector_angle_d(n, i, j)); return s * s; }
  }
}
static double vectorfun_c(int vfun, uint32_t n, const cx* i, const cx* j) {
  switch (vfun) {
    case 1: return vector_distance_c(n, i, j);
    case 2: return 1 - vector_product_c(n, i, j);
    case 3: return euclidean_angle_c(n, i, j);
    case 4: return hermitian_angle_c(n, i, j);
    default: { double s = sin(hermitian_angle_c(n, i, j)); return s * s; }
  }
}

/* Interpolator::add_cost (interpolator_cost.tpp:18-58); returns -1 for the combination that is not restated */
static int add_cost(const interp_t* t, uint32_t branches, uint32_t i0, uint32_t i1, double* cost, int arbitrary_phase) {
  const uint32_t s_ = t->el[0] + t->el[1] + t->el[2], b_ = branches, mo_ = t->el[0] + t->el[1];
  if (s_ == 0) return 0;
  double s_cost = 0, v_cost = 0, m_cost = 0;
  if (t->is_complex) {
    const cx* x0 = (const cx*)t->data + (size_t)i0 * b_ * s_;
    const cx* x1 = (const cx*)t->data + (size_t)i1 * b_ * s_;
    cx* phased = (cx*)malloc(sizeof(cx) * s_);
    for (uint32_t i = 0; i < b_; ++i) {
      const cx* x0i = x0 + (size_t)i * s_;
      for (uint32_t j = 0; j < b_; ++j) {
        const cx* rhs = x1 + (size_t)j * s_;
        if (arbitrary_phase) { /* inplace_antiphase: e^{-i arg<a|b>} b over the whole span */
          double real_dot = 0, imag_dot = 0;
          for (uint32_t e = 0; e < s_; ++e) {
            real_dot += x0i[e].re * rhs[e].re + x0i[e].im * rhs[e].im;
            imag_dot += x0i[e].re * rhs[e].im - x0i[e].im * rhs[e].re;
          }
          const double th = -1.0 * atan2(imag_dot, real_dot);
          const cx eith = {1.0 * cos(th), 1.0 * sin(th)}; /* std::polar(1, th) */
          for (uint32_t e = 0; e < s_; ++e) {
            phased[e].re = eith.re * rhs[e].re - eith.im * rhs[e].im;
            phased[e].im = eith.re * rhs[e].im + eith.im * rhs[e].re;
          }
          rhs = phased;
        }
        if (t->el[0]) {
          double s = 0;
          for (uint32_t z = 0; z < t->el[0]; ++z) {
            cx d = {x0i[z].re - rhs[z].re, x0i[z].im - rhs[z].im};
            s += magnitude_c(d);
          }
          s_cost = s;
        }
        if (t->el[1]) v_cost = vectorfun_c(t->vfun, t->el[1], x0i + t->el[0], rhs + t->el[0]);
        if (t->el[2]) {
          m_cost = 0;
          for (uint32_t m = 0; m < t->el[2] / 9; ++m) m_cost += vector_distance_c(9, x0i + mo_ + 9u * m, rhs + mo_ + 9u * m);
        }
        cost[(size_t)i * b_ + j] += t->costmult[0] * s_cost + t->costmult[1] * v_cost + t->costmult[2] * m_cost;
      }
    }
    free(phased);
  } else {
    if (arbitrary_phase) return -1;
    const double* x0 = t->data + (size_t)i0 * b_ * s_;
    const double* x1 = t->data + (size_t)i1 * b_ * s_;
    for (uint32_t i = 0; i < b_; ++i) {
      const double* x0i = x0 + (size_t)i * s_;
      for (uint32_t j = 0; j < b_; ++j) {
        const double* x1j = x1 + (size_t)j * s_;
        if (t->el[0]) {
          double s = 0;
          for (uint32_t z = 0; z < t->el[0]; ++z) s += fabs(x0i[z] - x1j[z]);
          s_cost = s;
        }
        if (t->el[1]) v_cost = vectorfun_d(t->vfun, t->el[1], x0i + t->el[0], x1j + t->el[0]);
        if (t->el[2]) {
          m_cost = 0;
          for (uint32_t m = 0; m < t->el[2] / 9; ++m) m_cost += vector_distance_d(9, x0i + mo_ + 9u * m, x1j + mo_ + 9u * m);
        }
        cost[(size_t)i * b_ + j] += t->costmult[0] * s_cost + t->costmult[1] * v_cost + t->costmult[2] * m_cost;
      }
    }
  }
  return 0;
}

/* lapjv (lapjv.hpp:281-538) with idx = int, cost = double */
static void lapjv(int dim, const double* assign_cost, int* rowsol, int* colsol, double* u, double* v) {
  int* freerows = (int*)malloc(sizeof(int) * dim);
  int* collist = (int*)malloc(sizeof(int) * dim);
  int* matches = (int*)malloc(sizeof(int) * dim);
  double* d = (double*)malloc(sizeof(double) * dim);
  int* pred = (int*)malloc(sizeof(int) * dim);
  if (1 == dim) {
    rowsol[0] = colsol[0] = 0;
    goto done;
  }
  for (int i = 0; i < dim; i++) matches[i] = 0;
  double total_cost = 0;
  for (int tc = 0; tc < dim * dim; ++tc) total_cost += assign_cost[tc];
  const double cost_epsilon = total_cost / (double)(10000 * dim);
  /* COLUMN REDUCTION */
  for (int j = dim; j-- > 0;) {
    double min = assign_cost[j];
    int imin = 0;
    for (int i = 1; i < dim; i++) {
      const double* local_cost = &assign_cost[i * dim];
      if (local_cost[j] < min) {
        min = local_cost[j];
        imin = i;
      }
    }
    v[j] = min;
    if (++matches[imin] == 1) {
      rowsol[imin] = j;
      colsol[j] = imin;
    } else {
      colsol[j] = -1;
    }
  }
  /* REDUCTION TRANSFER */
  int numfree = 0;
  for (int i = 0; i < dim; i++) {
    const double* local_cost = &assign_cost[i * dim];
    if (matches[i] == 0) {
      freerows[numfree++] = i;
    } else if (matches[i] == 1) {
      int j1 = rowsol[i];
      double min = DBL_MAX;
      for (int j = 0; j < dim; j++)
        if (j != j1)
          if (local_cost[j] - v[j] < min + cost_epsilon) min = local_cost[j] - v[j];
      v[j1] = v[j1] - min;
    }
  }
  /* AUGMENTING ROW REDUCTION */
  for (int loopcnt = 0; loopcnt < 2; loopcnt++) {
    int k = 0;
    int prevnumfree = numfree;
    numfree = 0;
    while (k < prevnumfree) {
      int i = freerows[k++];
      /* find_umins_plain (lapjv.hpp:74-99) */
      const double* local_cost = &assign_cost[i * dim];
      double umin = local_cost[0] - v[0];
      long long j1 = 0, j2 = -1;
      double usubmin = DBL_MAX;
      for (int j = 1; j < dim; j++) {
        double h = local_cost[j] - v[j];
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
      double vj1_new = v[j1] - (usubmin + cost_epsilon - umin);
      int vj1_lowers = vj1_new < v[j1];
      if (vj1_lowers) {
        v[j1] = vj1_new;
      } else if (i0 != -1) {
        j1 = j2;
        i0 = colsol[j2];
      }
      rowsol[i] = (int)j1;
      colsol[j1] = i;
      if (i0 != -1) {
        if (vj1_lowers) freerows[--k] = i0;
        else freerows[numfree++] = i0;
      }
    }
  }
  /* AUGMENT SOLUTION for each free row */
  for (int f = 0; f < numfree; f++) {
    int endofpath = 0;
    int freerow = freerows[f];
    for (int j = 0; j < dim; j++) {
      d[j] = assign_cost[freerow * dim + j] - v[j];
      pred[j] = freerow;
      collist[j] = j;
    }
    int low = 0, up = 0;
    int unassigned_found = 0;
    long long last = 0;
    double min = 0;
    do {
      if (up == low) {
        last = (long long)low - 1;
        min = d[collist[up++]];
        for (int k = up; k < dim; k++) {
          int j = collist[k];
          double h = d[j];
          if (h <= min) {
            if (h < min) {
              up = low;
              min = h;
            }
            collist[k] = collist[up];
            collist[up++] = j;
          }
        }
        for (int k = low; k < up; k++)
          if (colsol[collist[k]] == -1) {
            endofpath = collist[k];
            unassigned_found = 1;
            break;
          }
      }
      if (!unassigned_found) {
        int j1 = collist[low];
        low++;
        int i = colsol[j1];
        const double* local_cost = &assign_cost[i * dim];
        double h = local_cost[j1] - v[j1] - min;
        for (int k = up; k < dim; k++) {
          int j = collist[k];
          double v2 = local_cost[j] - v[j] - h;
          if (v2 < d[j]) {
            pred[j] = i;
            if (v2 == min) {
              if (colsol[j] == -1) {
                endofpath = j;
                unassigned_found = 1;
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
    for (long long k = 0; k <= last; k++) {
      int j1 = collist[k];
      v[j1] = v[j1] + d[j1] - min;
    }
    {
      int i;
      do {
        i = pred[endofpath];
        colsol[endofpath] = i;
        int j1 = endofpath;
        endofpath = rowsol[i];
        rowsol[i] = j1;
      } while (i != freerow);
    }
  }
  for (int i = 0; i < dim; i++) u[i] = assign_cost[i * dim + rowsol[i]] - v[rowsol[i]];
done:
  free(freerows);
  free(collist);
  free(matches);
  free(d);
  free(pred);
}

/* One call = DualInterpolator::sort() over the given pairs: row (n_pairs, branches) is the permutation stored for (i,j),
 * col the one stored for (j,i); cost_out (optional) receives the cost matrices (n_pairs, branches, branches).          */
int oracle_sort_pairs(const double* values, int v_cplx, const uint32_t* v_el, const double* v_cm, int v_vfun, const double* vectors,
                      int w_cplx, const uint32_t* w_el, const double* w_cm, int w_vfun, uint32_t branches, const uint32_t* pairs,
                      size_t n_pairs, int32_t* row, int32_t* col, double* cost_out) {
  interp_t V = {values, v_cplx, {v_el[0], v_el[1], v_el[2]}, {v_cm[0], v_cm[1], v_cm[2]}, v_vfun};
  interp_t W = {vectors, w_cplx, {w_el[0], w_el[1], w_el[2]}, {w_cm[0], w_cm[1], w_cm[2]}, w_vfun};
  const uint32_t B = branches;
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 16)
  for (long long k = 0; k < (long long)n_pairs; ++k) {
    double* cost = (double*)calloc((size_t)B * B, sizeof(double));
    double* u = (double*)calloc(B, sizeof(double));
    double* v = (double*)calloc(B, sizeof(double));
    const uint32_t i0 = pairs[2 * k], i1 = pairs[2 * k + 1];
    if (i0 == i1) {
      for (uint32_t j = 0; j < B * B; j += B + 1) cost[j] = -1.0;
    } else {
      if (add_cost(&V, B, i0, i1, cost, 0) || add_cost(&W, B, i0, i1, cost, 1)) {
#pragma omp atomic write
        rc = -1;
      }
    }
    lapjv((int)B, cost, row + (size_t)k * B, col + (size_t)k * B, u, v);
    if (cost_out) memcpy(cost_out + (size_t)k * B * B, cost, sizeof(double) * B * B);
    free(cost);
    free(u);
    free(v);
  }
  return rc;
}

/* The assignment step alone on given cost matrices (n_pairs, branches, branches): lets the tests check the integer part of
 * the device path bit for bit on the device's own cost matrices. */
int oracle_lapjv_batch(const double* cost, size_t n_pairs, uint32_t branches, int32_t* row, int32_t* col) {
  const uint32_t B = branches;
#pragma omp parallel for schedule(dynamic, 16)
  for (long long k = 0; k < (long long)n_pairs; ++k) {
    double* u = (double*)calloc(B, sizeof(double));
    double* v = (double*)calloc(B, sizeof(double));
    lapjv((int)B, cost + (size_t)k * B * B, row + (size_t)k * B, col + (size_t)k * B, u, v);
    free(u);
    free(v);
  }
  return 0;
}
