/* oracle/brille_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar, plain-C restatement of brille's batched Q-point interpolation path working on the flat
 * tables of include/brille_b200.h.  It exists only so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py can check the CUDA kernels; nothing under brille_b200/ may call it.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py compares this file against the reference itself
 * (oracle/_ref, built from /root/reference by oracle/build_ref.sh) and against the reference's own
 * golden file wrap/tests/test_5_gamma.npz (fixtures under tests/golden/).
 *
 * Every function cites the reference lines it follows.  The floating-point operation ORDER of the
 * reference is kept (x86-64 without FMA contraction: compile with -ffp-contract=off), including the
 * metric-aware lattice-vector arithmetic of LVec, so that decisions agree bit for bit.
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "brille_b200.h"

#define TWO_PI 6.283185307179586476925286766559005768394338798750211641949889

/* ------------------------------------------------------------------------------------------------
 * approx_float  (src/approx_float.hpp:78-100 `tols`, :171-188 `_scalar`; TOL_MULT=10000 :52-56)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double rel, abs_; } tol_t;
static tol_t make_tol(double tol, int digit) {
  tol_t t;
  t.rel = DBL_EPSILON * (double)digit * 10000.0;
  t.abs_ = 5.0 / 1000000000000000.0;
  if (tol > t.rel) t.rel = tol;
  if (tol > t.abs_) t.abs_ = tol;
  return t;
}
static int approx(double a, double b, tol_t t) {
  double x = fabs(a - b);
  return x <= t.abs_ + t.rel * fabs(a + b) || x < DBL_MIN;
}

/* ------------------------------------------------------------------------------------------------
 * 3x3 helpers in the reference's accumulation order (utilities.tpp:39-44, 50-53)
 * ---------------------------------------------------------------------------------------------- */
static void matvec_dd(double* c, const double* A, const double* b) {
  for (int i = 0; i < 3; ++i) {
    c[i] = 0.0;
    for (int k = 0; k < 3; ++k) c[i] += A[i * 3 + k] * b[k];
  }
}
static void matvec_id(double* c, const int32_t* A, const double* b) {
  for (int i = 0; i < 3; ++i) {
    c[i] = 0.0;
    for (int k = 0; k < 3; ++k) c[i] += (double)A[i * 3 + k] * b[k];
  }
}
static void matvec_ii(int32_t* c, const int32_t* A, const int32_t* b) {
  for (int i = 0; i < 3; ++i) {
    c[i] = 0;
    for (int k = 0; k < 3; ++k) c[i] += A[i * 3 + k] * b[k];
  }
}
/* same_lattice_dot (array_functions.hpp:246-255): (metric * x) . y */
static double lat_dot(const double* metric, const double* x, const double* y) {
  double tmp[3];
  matvec_dd(tmp, metric, x);
  double out = 0.0;
  for (int i = 0; i < 3; ++i) out += tmp[i] * y[i];
  return out;
}
/* vector_cross (utilities.hpp:369-376) */
static void cross3(double* c, const double* a, const double* b) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* ------------------------------------------------------------------------------------------------
 * pseudo_orient3d for lattice vectors (geometry.hpp:118-138) =
 *   dot(a-d, cross(b-d, c-d)) with LVec cross (array_functions.hpp:190-201) and LVec::star
 *   (array_lvec_methods.tpp:53-67) and LVec dot (array_functions.hpp:257-290)
 * ---------------------------------------------------------------------------------------------- */
static double lat_orient3d(const double* recip_metric, const double* real_metric, double recip_volume,
                           const double* a, const double* b, const double* c, const double* d) {
  double u[3], v[3], w[3], cr[3], st[3];
  for (int i = 0; i < 3; ++i) {
    u[i] = a[i] - d[i];
    v[i] = b[i] - d[i];
    w[i] = c[i] - d[i];
  }
  cross3(cr, v, w);
  double s = recip_volume / TWO_PI;
  for (int i = 0; i < 3; ++i) cr[i] *= s;
  matvec_dd(st, real_metric, cr);
  for (int i = 0; i < 3; ++i) st[i] /= TWO_PI;
  return lat_dot(recip_metric, u, st);
}
/* point_inside_all_planes (geometry.hpp:412-418) */
static int inside_all_planes(const double* recip_metric, const double* real_metric, double recip_volume, int nf,
                             const double* pa, const double* pb, const double* pc, const double* x, tol_t t) {
  double m = 0.0;
  for (int f = 0; f < nf; ++f) {
    double o = lat_orient3d(recip_metric, real_metric, recip_volume, pa + 3 * f, pb + 3 * f, pc + 3 * f, x);
    if (f == 0 || o < m) m = o;
  }
  return m > 0 || approx(m, 0.0, t);
}

/* ------------------------------------------------------------------------------------------------
 * moveinto for one Q (bz_move.cpp:103-163 and part_moveinto_prim :14-53)
 * returns 0 if the final q is inside the first zone, B200_ST_OUTSIDE_BZ otherwise
 * ---------------------------------------------------------------------------------------------- */
static uint32_t moveinto_one(const b200_bz_tables_t* bz, const double* Q, double* q_out, int32_t* tau_out) {
  tol_t cfg = make_tol(bz->float_tolerance, bz->approx_tolerance);
  tol_t def = make_tol(0.0, 1);
  const int F = bz->n_faces;
  double Qp[3];
  if (bz->transform_needed) { /* transform.hpp:182-190 */
    matvec_id(Qp, bz->P6t, Q);
    for (int i = 0; i < 3; ++i) Qp[i] /= 6.0;
  } else {
    for (int i = 0; i < 3; ++i) Qp[i] = Q[i];
  }
  int32_t tau[3], last[3];
  double q[3];
  for (int i = 0; i < 3; ++i) {
    tau[i] = (int32_t)round(Qp[i]); /* array2.tpp:498-505: std::round, half away from zero */
    q[i] = Qp[i] - (double)tau[i];
    last[i] = tau[i];
  }
  double d[32];
  int N[32];
  int count = 0;
  while (count++ < F && !inside_all_planes(bz->w_recip_metric, bz->w_real_metric, bz->w_recip_volume, F, bz->pa, bz->pb, bz->pc, q, cfg)) {
    int any = 0;
    for (int j = 0; j < F; ++j) {
      d[j] = lat_dot(bz->w_recip_metric, q, bz->normals + 3 * j); /* dot(q_i, normals) */
      N[j] = (int)round(d[j] / bz->tau_lens[j]);
      if (N[j] > 0) any = 1;
    }
    if (any) {
      int max_nm = 0, max_at = 0;
      for (int j = 0; j < F; ++j) {
        if (N[j] > 0 && N[j] >= max_nm) {
          int ok = (0 == max_nm);
          if (!ok) {
            /* norm(taus.view(j) + last_shift).all(gt, 0.) && d[j] > d[max_at]  (bz_move.cpp:36-41) */
            double s[3];
            for (int i = 0; i < 3; ++i) s[i] = (double)(bz->taus[3 * j + i] + last[i]);
            double nrm = sqrt(lat_dot(bz->w_recip_metric, s, s));
            int gt0 = !approx(nrm, 0.0, def) && nrm > 0.0;
            ok = gt0 && d[j] > d[max_at];
          }
          if (ok) {
            max_at = j;
            max_nm = N[j];
          }
        }
      }
      for (int i = 0; i < 3; ++i) {
        int32_t t = bz->taus[3 * max_at + i];
        q[i] -= (double)t * (double)max_nm;
        tau[i] += t * max_nm;
        last[i] = t * max_nm;
      }
    }
  }
  if (bz->transform_needed) { /* transform.hpp:222-230 */
    matvec_id(q_out, bz->invPt, q);
    matvec_ii(tau_out, bz->invPt, tau);
  } else {
    for (int i = 0; i < 3; ++i) {
      q_out[i] = q[i];
      tau_out[i] = tau[i];
    }
  }
  /* isinside re-check in the conventional lattice (bz_move.cpp:149, bz.hpp:631-642, polyhedron_faces.hpp:314-333) */
  if (!inside_all_planes(bz->o_recip_metric, bz->o_real_metric, bz->o_recip_volume, F, bz->ca, bz->cb, bz->cc, q_out, cfg))
    return B200_ST_OUTSIDE_BZ;
  return 0;
}

/* _inside_wedge_outer (bz.hpp:757-763) with Array2::all (array2.tpp:664-672) */
/* BrillouinZone::isinside (bz.hpp:631-640): the conventional-lattice plane test (the one moveinto re-checks its result with) */
static int isinside_one(const b200_bz_tables_t* bz, const double* q) {
  tol_t cfg = make_tol(bz->float_tolerance, bz->approx_tolerance);
  return inside_all_planes(bz->o_recip_metric, bz->o_real_metric, bz->o_recip_volume, bz->n_faces, bz->ca, bz->cb, bz->cc, q, cfg);
}

static int inside_wedge(const b200_bz_tables_t* bz, const double* q) {
  const int K = bz->n_wedge;
  if (K == 0) return 1;
  double dots[16];
  for (int k = 0; k < K; ++k) dots[k] = lat_dot(bz->o_recip_metric, bz->wedge_normals + 3 * k, q);
  if (bz->no_ir_mirroring) {
    tol_t cfg = make_tol(bz->float_tolerance, bz->approx_tolerance);
    for (int k = 0; k < K; ++k)
      if (!(approx(dots[k], 0.0, cfg) || dots[k] > 0.0)) return 0;
    return 1;
  }
  /* le_ge: tolerances are dropped (array2.tpp:666-667) */
  tol_t def = make_tol(0.0, 1);
  int all_le = 1, all_ge = 1;
  for (int k = 0; k < K; ++k) {
    if (!(approx(dots[k], 0.0, def) || dots[k] < 0.0)) all_le = 0;
    if (!(approx(dots[k], 0.0, def) || dots[k] > 0.0)) all_ge = 0;
  }
  return all_le || all_ge;
}

/* ir_moveinto for one Q (bz_move.cpp:165-296) */
static uint32_t wedge_search(const b200_bz_tables_t* bz, double* q, int* ridx, int* invridx, uint32_t st);
static uint32_t ir_moveinto_one(const b200_bz_tables_t* bz, const double* Q, double* q, int32_t* tau, int* ridx, int* invridx) {
  return wedge_search(bz, q, ridx, invridx, moveinto_one(bz, Q, q, tau));
}
/* ir_moveinto_wedge (bz_move.cpp:299-356): the same rotation search applied to Q itself */
static uint32_t wedge_rotate_one(const b200_bz_tables_t* bz, const double* Q, double* q, int* ridx, int* invridx) {
  for (int i = 0; i < 3; ++i) q[i] = Q[i];
  return wedge_search(bz, q, ridx, invridx, 0u);
}
/* the wedge loop shared by both (bz_move.cpp:257-285, 330-347) */
static uint32_t wedge_search(const b200_bz_tables_t* bz, double* q, int* ridx, int* invridx, uint32_t st) {
  if (inside_wedge(bz, q)) {
    *ridx = *invridx = bz->identity_index;
    return st;
  }
  for (int j = 0; j < bz->n_ops; ++j) {
    const int32_t* R = bz->rotations + 9 * j;
    int32_t Rt[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) Rt[a * 3 + b] = R[b * 3 + a];
    double qj[3];
    matvec_id(qj, Rt, q);
    if (inside_wedge(bz, qj)) {
      for (int i = 0; i < 3; ++i) q[i] = qj[i];
      *invridx = j;
      *ridx = bz->inverse_index[j];
      return st;
    }
  }
  *ridx = *invridx = 0; /* the reference leaves the zero-initialised entries (bz_trellis.hpp:158) */
  return st | B200_ST_OUTSIDE_WEDGE;
}

/* ------------------------------------------------------------------------------------------------
 * trellis point location
 * ---------------------------------------------------------------------------------------------- */
/* find_bin (trellis_poly.hpp:67-73) */
static size_t find_bin(const double* k, int n, double x) {
  int d = 0;
  while (d < n && !(k[d] > x)) ++d;
  if (d > n - 1 && x < k[0]) d = 0;
  return d > 0 ? (size_t)(d - 1) : (size_t)d;
}
/* on_boundary (trellis_poly.hpp:75-81) */
static int on_boundary(const double* k, int n, double x, size_t i, tol_t def) {
  if (i + 2 < (size_t)n && approx(k[i + 1], x, def)) return 1;
  if (i > 0 && approx(k[i], x, def)) return -1;
  return 0;
}
static int node_is_null(const b200_trellis_tables_t* t, size_t idx) {
  uint8_t ty = t->node_type[idx];
  return ty == B200_NODE_NULL || ty == B200_NODE_ASSUMED_NULL || ty == B200_NODE_FOUND_NULL;
}
static int sub_ok_not_null(const b200_trellis_tables_t* t, const size_t* sub) {
  for (int d = 0; d < 3; ++d)
    if (sub[d] >= (size_t)(t->n_knots[d] - 1)) return 0;
  size_t n0 = t->n_knots[0] - 1, n1 = t->n_knots[1] - 1;
  return !node_is_null(t, sub[0] + n0 * sub[1] + n0 * n1 * sub[2]);
}
/* node_subscript (trellis_poly.hpp:382-434); returns 0 when no usable node exists */
static int node_subscript(const b200_trellis_tables_t* t, const double* x, size_t* sub, uint32_t* st) {
  tol_t def = make_tol(0.0, 1);
  for (int d = 0; d < 3; ++d) sub[d] = find_bin(t->knots[d], t->n_knots[d], x[d]);
  int bad = !sub_ok_not_null(t, sub);
  if (bad) {
    int close[3], num_close = 0;
    for (int i = 0; i < 3; ++i) {
      close[i] = on_boundary(t->knots[i], t->n_knots[i], x[i], sub[i], def);
      if (close[i]) ++num_close;
    }
    size_t ns[3] = {sub[0], sub[1], sub[2]};
    if (num_close > 0)
      for (int i = 0; i < 3 && bad; ++i)
        if (close[i]) {
          ns[0] = sub[0]; ns[1] = sub[1]; ns[2] = sub[2];
          ns[i] += close[i];
          bad = !sub_ok_not_null(t, ns);
        }
    if (bad && num_close > 1)
      for (int i = 0; i < 3 && bad; ++i)
        if (close[i])
          for (int j = 0; j < 3 && bad; ++j)
            if (close[j]) {
              ns[0] = sub[0]; ns[1] = sub[1]; ns[2] = sub[2];
              ns[i] += close[i];
              ns[j] += close[j];
              bad = !sub_ok_not_null(t, ns);
            }
    if (bad && num_close > 2) {
      for (int i = 0; i < 3; ++i) ns[i] = sub[i] + close[i];
      bad = !sub_ok_not_null(t, ns);
    }
    if (!bad) {
      sub[0] = ns[0]; sub[1] = ns[1]; sub[2] = ns[2];
      *st |= B200_ST_NEIGHBOUR;
    }
  }
  return !bad;
}

/* pseudo_orient3d for bare arrays (geometry.hpp:118-138): (a-d).((b-d)x(c-d)) */
static double orient3d_plain(const double* a, const double* b, const double* c, const double* d) {
  double u[3], v[3], w[3], cr[3];
  for (int i = 0; i < 3; ++i) {
    u[i] = a[i] - d[i];
    v[i] = b[i] - d[i];
    w[i] = c[i] - d[i];
  }
  cross3(cr, v, w);
  double out = 0.0;
  for (int i = 0; i < 3; ++i) out += u[i] * cr[i];
  return out;
}

typedef struct {
  int n;
  uint32_t vertex[8];
  double weight[8];
  uint8_t slot[8]; /* corner of the cell (cube 0-7, tet 0-3) each emitted vertex came from */
  uint32_t cell;   /* node linear index */
  int32_t tet;     /* global tet index or -1 */
} iw_t;

/* CubeNode::indices_weights (trellis_node.hpp:130-149) */
static void cube_weights(const b200_trellis_tables_t* t, uint32_t cube, const double* x, iw_t* iw) {
  tol_t def = make_tol(0.0, 1);
  const uint32_t* vi = t->cube_vertices + 8 * (size_t)cube;
  const double* v0 = t->vertices + 3 * (size_t)vi[0];
  const double* v7 = t->vertices + 3 * (size_t)vi[7];
  double vol = 1.0;
  for (int d = 0; d < 3; ++d) vol *= fabs(v0[d] - v7[d]);
  iw->n = 0;
  for (int i = 0; i < 8; ++i) {
    const double* v = t->vertices + 3 * (size_t)vi[i];
    double w = 1.0;
    for (int d = 0; d < 3; ++d) w *= fabs(x[d] - v[d]);
    w = w / vol;
    if (!approx(w, 0.0, def) && w > 0.0) { /* w.is(gt, 0.) */
      iw->vertex[iw->n] = vi[7 - i];
      iw->weight[iw->n] = w;
      iw->slot[iw->n] = (uint8_t)(7 - i);
      ++iw->n;
    }
  }
}

/* PolyNode::tetrahedra_contains (trellis_node.hpp:319-339) + tetrahedra_might_contain (:349-364) */
static double tet_contains(const b200_trellis_tables_t* t, uint32_t tet, const double* x, double* w, int shortcut) {
  tol_t def = make_tol(0.0, 1);
  if (shortcut) {
    const double* ci = t->tet_circum + 4 * (size_t)tet;
    double v[3] = {ci[0] - x[0], ci[1] - x[1], ci[2] - x[2]};
    double d2 = 0.0, r2 = ci[3] * ci[3];
    for (int i = 0; i < 3; ++i) d2 += v[i] * v[i];
    double away = (d2 < r2 || approx(d2, r2, def)) ? 0.0 : -d2;
    if (away < 0.0) return away;
  }
  const uint32_t* vi = t->tet_vertices + 4 * (size_t)tet;
  const double* p0 = t->vertices + 3 * (size_t)vi[0];
  const double* p1 = t->vertices + 3 * (size_t)vi[1];
  const double* p2 = t->vertices + 3 * (size_t)vi[2];
  const double* p3 = t->vertices + 3 * (size_t)vi[3];
  double vol6 = t->tet_volume[tet] * 6.0;
  w[0] = orient3d_plain(x, p1, p2, p3) / vol6;
  w[1] = orient3d_plain(p0, x, p2, p3) / vol6;
  w[2] = orient3d_plain(p0, p1, x, p3) / vol6;
  w[3] = orient3d_plain(p0, p1, p2, x) / vol6;
  int neg = 0;
  for (int j = 0; j < 4; ++j)
    if (w[j] < 0.0 && !approx(w[j], 0.0, def)) neg = 1;
  if (neg) {
    double m = w[0];
    for (int j = 1; j < 4; ++j)
      if (w[j] < m) m = w[j];
    return m;
  }
  return 0.0;
}
/* PolyNode::indices_weights (trellis_node.hpp:273-308), should_contain = true (trellis_poly.hpp:252-253) */
static void poly_weights(const b200_trellis_tables_t* t, uint32_t poly, const double* x, iw_t* iw, uint32_t* st) {
  tol_t def = make_tol(0.0, 1);
  uint32_t t0 = t->poly_offsets[poly], t1 = t->poly_offsets[poly + 1];
  double w[4] = {0, 0, 0, 0};
  iw->n = 0;
  double best = 0.0;
  uint32_t best_at = t0;
  for (uint32_t k = t0; k < t1; ++k) {
    double mn = tet_contains(t, k, x, w, 1);
    if (mn >= 0.0) {
      for (int j = 0; j < 4; ++j)
        if (!approx(w[j], 0.0, def)) {
          iw->vertex[iw->n] = t->tet_vertices[4 * (size_t)k + j];
          iw->weight[iw->n] = w[j];
          iw->slot[iw->n] = (uint8_t)j;
          ++iw->n;
        }
      iw->tet = (int32_t)k;
      return;
    }
    if (k == t0 || mn > best) { /* std::max_element: first maximum */
      best = mn;
      best_at = k;
    }
  }
  if (t1 == t0) return;
  *st |= B200_ST_FALLBACK_TET;
  tet_contains(t, best_at, x, w, 0);
  for (int j = 0; j < 4; ++j)
    if (!approx(w[j], 0.0, def)) {
      iw->vertex[iw->n] = t->tet_vertices[4 * (size_t)best_at + j];
      iw->weight[iw->n] = w[j];
      iw->slot[iw->n] = (uint8_t)j;
      ++iw->n;
    }
  iw->tet = (int32_t)best_at;
}

/* PolyTrellis::indices_weights (trellis_poly.hpp:245-255) */
static uint32_t trellis_locate(const b200_trellis_tables_t* t, const double* x, iw_t* iw) {
  uint32_t st = 0;
  size_t sub[3];
  iw->n = 0;
  iw->tet = -1;
  iw->cell = 0xffffffffu;
  if (!node_subscript(t, x, sub, &st)) return st | B200_ST_NOT_FOUND; /* null-node access -> logic_error */
  size_t n0 = t->n_knots[0] - 1, n1 = t->n_knots[1] - 1;
  size_t idx = sub[0] + n0 * sub[1] + n0 * n1 * sub[2];
  iw->cell = (uint32_t)idx;
  if (t->node_type[idx] == B200_NODE_CUBE)
    cube_weights(t, t->node_index[idx], x, iw);
  else
    poly_weights(t, t->node_index[idx], x, iw, &st);
  if (iw->n < 1) st |= B200_ST_NOT_FOUND;
  return st;
}

/* ------------------------------------------------------------------------------------------------
 * orient3d of the vendored TetGen (lib/tetgen/predicates.cxx:1997-2050, Shewchuk's adaptive predicate).
 * Fast path: the reference's floating-point determinant in its own operation order, returned as is when
 * |det| > o3derrboundA * permanent (predicates.cxx:436-443: o3derrboundA = (7 + 56 eps) eps with eps = 2^-53).
 * Slow path: the reference refines the value in exact expansion arithmetic (orient3dadapt).  Here the determinant
 * is re-evaluated in double-double arithmetic (error ~1e-31 * permanent, far below every tolerance brille
 * applies to the result: the weights are only ever compared through approx_float with |w| <= 5e-15).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double hi, lo; } dd_t;
static dd_t dd_two_sum(double a, double b) {
  double s = a + b, bb = s - a;
  dd_t r = {s, (a - (s - bb)) + (b - bb)};
  return r;
}
static dd_t dd_two_diff(double a, double b) {
  double s = a - b, bb = s - a;
  dd_t r = {s, (a - (s - bb)) - (b + bb)};
  return r;
}
static dd_t dd_two_prod(double a, double b) {
  double p = a * b;
  dd_t r = {p, fma(a, b, -p)};
  return r;
}
static dd_t dd_norm(double hi, double lo) {
  double s = hi + lo;
  dd_t r = {s, lo - (s - hi)};
  return r;
}
static dd_t dd_add(dd_t a, dd_t b) {
  dd_t s = dd_two_sum(a.hi, b.hi), t = dd_two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = dd_norm(s.hi, s.lo);
  s.lo += t.lo;
  return dd_norm(s.hi, s.lo);
}
static dd_t dd_neg(dd_t a) { dd_t r = {-a.hi, -a.lo}; return r; }
static dd_t dd_mul(dd_t a, dd_t b) {
  dd_t p = dd_two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return dd_norm(p.hi, p.lo);
}
static double orient3d_tetgen(const double* pa, const double* pb, const double* pc, const double* pd) {
  const double eps = 1.1102230246251565e-16; /* 2^-53, exactinit(): predicates.cxx:380-443 */
  const double o3derrboundA = (7.0 + 56.0 * eps) * eps;
  double adx = pa[0] - pd[0], ady = pa[1] - pd[1], adz = pa[2] - pd[2];
  double bdx = pb[0] - pd[0], bdy = pb[1] - pd[1], bdz = pb[2] - pd[2];
  double cdx = pc[0] - pd[0], cdy = pc[1] - pd[1], cdz = pc[2] - pd[2];
  double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, cdxady = cdx * ady, adxcdy = adx * cdy, adxbdy = adx * bdy, bdxady = bdx * ady;
  double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
  double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * fabs(adz) + (fabs(cdxady) + fabs(adxcdy)) * fabs(bdz) +
                     (fabs(adxbdy) + fabs(bdxady)) * fabs(cdz);
  double errbound = o3derrboundA * permanent;
  if (det > errbound || -det > errbound) return det;
  /* double-double re-evaluation on the exact differences */
  dd_t ax = dd_two_diff(pa[0], pd[0]), ay = dd_two_diff(pa[1], pd[1]), az = dd_two_diff(pa[2], pd[2]);
  dd_t bx = dd_two_diff(pb[0], pd[0]), by = dd_two_diff(pb[1], pd[1]), bzz = dd_two_diff(pb[2], pd[2]);
  dd_t cx = dd_two_diff(pc[0], pd[0]), cy = dd_two_diff(pc[1], pd[1]), cz = dd_two_diff(pc[2], pd[2]);
  dd_t t1 = dd_mul(az, dd_add(dd_mul(bx, cy), dd_neg(dd_mul(cx, by))));
  dd_t t2 = dd_mul(bzz, dd_add(dd_mul(cx, ay), dd_neg(dd_mul(ax, cy))));
  dd_t t3 = dd_mul(cz, dd_add(dd_mul(ax, by), dd_neg(dd_mul(bx, ay))));
  dd_t r = dd_add(dd_add(t1, t2), t3);
  return r.hi + r.lo;
}

/* ------------------------------------------------------------------------------------------------
 * Nest point location  (nest.hpp:59-123 NestLeaf, :163-222 NestNode::indices_weights)
 * ---------------------------------------------------------------------------------------------- */
static int nest_weights(const b200_nest_tables_t* t, uint32_t node, const double* x, tol_t tol, double* w) {
  /* NestLeaf::weights (:78-88): {-1,-1,-1,-1} unless might_contain (:114-122) */
  const double* cr = t->node_circum + 4 * (size_t)node;
  double d[3] = {x[0] - cr[0], x[1] - cr[1], x[2] - cr[2]};
  double d2 = 0.0, r2 = cr[3] * cr[3];
  for (int i = 0; i < 3; ++i) d2 += d[i] * d[i];
  w[0] = w[1] = w[2] = w[3] = -1.0;
  if (d2 < r2 || approx(d2, r2, tol)) {
    const uint32_t* vi = t->node_vertices + 4 * (size_t)node;
    const double* p0 = t->vertices + 3 * (size_t)vi[0];
    const double* p1 = t->vertices + 3 * (size_t)vi[1];
    const double* p2 = t->vertices + 3 * (size_t)vi[2];
    const double* p3 = t->vertices + 3 * (size_t)vi[3];
    double vol6 = t->node_volume[node] * 6.0;
    w[0] = orient3d_tetgen(x, p1, p2, p3) / vol6;
    w[1] = orient3d_tetgen(p0, x, p2, p3) / vol6;
    w[2] = orient3d_tetgen(p0, p1, x, p3) / vol6;
    w[3] = orient3d_tetgen(p0, p1, p2, x) / vol6;
  }
  /* none_negative (:43-48) */
  for (int j = 0; j < 4; ++j)
    if (w[j] < 0 && !approx(w[j], 0.0, tol)) return 0;
  return 1;
}

static uint32_t nest_locate(const b200_nest_tables_t* t, const double* x, iw_t* iw) {
  tol_t tol = make_tol(t->tolerance, t->digit);
  iw->n = 0;
  iw->tet = -1;
  iw->cell = 0xffffffffu;
  /* breadth-first over containing nodes (the reference's std::deque, :175-192) */
  uint32_t cap = 64, head = 0, tail = 0;
  uint32_t* queue = (uint32_t*)malloc(sizeof(uint32_t) * cap);
  for (uint32_t c = t->child_begin[0]; c < t->child_end[0]; ++c) {
    if (tail == cap) queue = (uint32_t*)realloc(queue, sizeof(uint32_t) * (cap *= 2));
    queue[tail++] = c;
  }
  int nsol = 0;
  uint32_t first_node = 0, best_node = 0;
  double first_w[4], best_w[4], w[4];
  int have_best = 0;
  while (head < tail) {
    uint32_t node = queue[head++];
    if (nest_weights(t, node, x, tol, w)) {
      if (t->node_is_leaf[node]) {
        if (nsol == 0) { first_node = node; memcpy(first_w, w, sizeof(w)); }
        /* "best": the LAST solution whose weights are all > 0, else the first (:195-201) */
        int all_pos = 1;
        for (int j = 0; j < 4; ++j) if (w[j] <= 0) all_pos = 0;
        if (all_pos) { best_node = node; memcpy(best_w, w, sizeof(w)); have_best = 1; }
        ++nsol;
      } else {
        for (uint32_t c = t->child_begin[node]; c < t->child_end[node]; ++c) {
          if (tail == cap) queue = (uint32_t*)realloc(queue, sizeof(uint32_t) * (cap *= 2));
          queue[tail++] = c;
        }
      }
    }
  }
  free(queue);
  if (nsol == 0) return B200_ST_NOT_FOUND;
  uint32_t node = first_node;
  double sw[4];
  memcpy(sw, first_w, sizeof(sw));
  if (nsol > 1 && have_best) { node = best_node; memcpy(sw, best_w, sizeof(sw)); }
  /* fold ~0 weights into the largest (:203-217) */
  for (int i = 0; i < 4; ++i)
    if (approx(sw[i], 0.0, tol)) {
      int max_at = 0;
      for (int j = 0; j < 4; ++j)
        if (!approx(sw[j], 0.0, tol) && sw[j] > sw[max_at]) max_at = j;
      sw[max_at] += sw[i];
    }
  const uint32_t* vi = t->node_vertices + 4 * (size_t)node;
  for (int i = 0; i < 4; ++i)
    if (!approx(sw[i], 0.0, tol)) {
      iw->vertex[iw->n] = vi[i];
      iw->weight[iw->n] = sw[i];
      iw->slot[iw->n] = (uint8_t)i;
      ++iw->n;
    }
  iw->cell = node;
  iw->tet = (int32_t)node;
  return iw->n < 1 ? B200_ST_NOT_FOUND : 0;
}

/* ------------------------------------------------------------------------------------------------
 * Mesh point location  (triangulation_layers.hpp:177-189 unsafe_locate, :254-288, :401-417 TetTri::locate)
 * ---------------------------------------------------------------------------------------------- */
static int mesh_try(const b200_mesh_tables_t* t, uint32_t layer, uint32_t tet, const double* x, double* w) {
  tol_t def = make_tol(0.0, 1);
  size_t g = (size_t)t->tet_offset[layer] + tet;
  /* unsafe_might_contain (:254-256): norm(x - centre) <= radius with the default tolerance */
  const double* c = t->centres + 3 * g;
  double d[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
  double s = 0.0;
  for (int i = 0; i < 3; ++i) s += d[i] * d[i];
  double nrm = sqrt(s), r = t->radii[g];
  if (!(approx(nrm, r, def) || nrm < r)) return 0;
  const uint32_t* vi = t->tets + 4 * g;
  const double* vb = t->vertices + 3 * (size_t)t->vert_offset[layer];
  const double *p0 = vb + 3 * (size_t)vi[0], *p1 = vb + 3 * (size_t)vi[1], *p2 = vb + 3 * (size_t)vi[2], *p3 = vb + 3 * (size_t)vi[3];
  double vol6 = t->vol6[g];
  w[0] = orient3d_tetgen(x, p1, p2, p3) / vol6; /* weights (:265-288) */
  w[1] = orient3d_tetgen(p0, x, p2, p3) / vol6;
  w[2] = orient3d_tetgen(p0, p1, x, p3) / vol6;
  w[3] = orient3d_tetgen(p0, p1, p2, x) / vol6;
  for (int j = 0; j < 4; ++j)
    if (!(w[j] > 0.0 || approx(w[j], 0.0, def))) return 0; /* unsafe_contains (:261-264) */
  return 1;
}

static uint32_t mesh_locate(const b200_mesh_tables_t* t, const double* x, iw_t* iw) {
  tol_t def = make_tol(0.0, 1);
  iw->n = 0;
  iw->tet = -1;
  iw->cell = 0xffffffffu;
  double w[4];
  uint32_t idx = 0xffffffffu;
  for (uint32_t layer = 0; layer < t->n_layers; ++layer) {
    uint32_t found = 0xffffffffu;
    if (layer == 0) {
      uint32_t nt = t->tet_offset[1] - t->tet_offset[0];
      for (uint32_t k = 0; k < nt && found == 0xffffffffu; ++k)
        if (mesh_try(t, 0, k, x, w)) found = k;
    } else {
      size_t c = (size_t)t->tet_offset[layer - 1] + idx;
      for (uint32_t k = t->conn_offset[c]; k < t->conn_offset[c + 1] && found == 0xffffffffu; ++k)
        if (mesh_try(t, layer, t->conn_index[k], x, w)) found = t->conn_index[k];
    }
    if (found == 0xffffffffu) return B200_ST_NOT_FOUND; /* the reference indexes connections[..][nTetrahedra]: undefined */
    idx = found;
  }
  uint32_t last = t->n_layers - 1;
  const uint32_t* vi = t->tets + 4 * ((size_t)t->tet_offset[last] + idx);
  for (int i = 0; i < 4; ++i)
    if (!approx(w[i], 0.0, def)) {
      iw->vertex[iw->n] = vi[i];
      iw->weight[iw->n] = w[i];
      iw->slot[iw->n] = (uint8_t)i;
      ++iw->n;
    }
  iw->cell = idx;
  iw->tet = (int32_t)idx;
  return iw->n < 1 ? B200_ST_NOT_FOUND : 0;
}

/* ------------------------------------------------------------------------------------------------
 * interpolation  (interpolatordual.hpp:149-155,374-382; interpolator_at.tpp:91-127)
 * ---------------------------------------------------------------------------------------------- */
static uint32_t span_of(const b200_interp_desc_t* d) { return d->elements[0] + d->elements[1] + d->elements[2]; }

static const uint32_t* perm_row(const b200_data_tables_t* data, int kind_is_cube, uint32_t cellidx, int pivot_slot, int slot) {
  if (data->n_perm_rows <= 1 || !data->perm_rows) return NULL;
  uint32_t row;
  if (kind_is_cube)
    row = data->cube_perm ? data->cube_perm[(size_t)cellidx * 64 + pivot_slot * 8 + slot] : 0;
  else
    row = data->tet_perm ? data->tet_perm[(size_t)cellidx * 16 + pivot_slot * 4 + slot] : 0;
  return data->perm_rows + (size_t)row * data->values.branches;
}

/* utils::antiphase (utilities.tpp:567-579) */
static void antiphase(uint32_t n, const double* a, const double* b, double* re, double* im) {
  double rd = 0.0, id = 0.0;
  for (uint32_t i = 0; i < n; ++i) {
    double ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
    rd += ar * br + ai * bi;
    id += ar * bi - ai * br;
  }
  double th = -1.0 * atan2(id, rd);
  *re = 1.0 * cos(th); /* std::polar(1, theta) */
  *im = 1.0 * sin(th);
}

static void interpolate_one(const b200_interp_desc_t* d, const b200_data_tables_t* data, const iw_t* iw, int is_cube,
                            uint32_t cellidx, int phase, double* out) {
  const uint32_t B = d->branches, S = span_of(d);
  const size_t row = (size_t)B * S;
  if (d->is_complex) {
    const double* base = (const double*)d->data;
    memset(out, 0, row * 2 * sizeof(double));
    const double* d0 = base + (size_t)iw->vertex[0] * row * 2;
    for (int i = 0; i < iw->n; ++i) {
      const double* dx = base + (size_t)iw->vertex[i] * row * 2;
      const uint32_t* perm = perm_row(data, is_cube, cellidx, iw->slot[0], iw->slot[i]);
      for (uint32_t b = 0; b < B; ++b) {
        uint32_t p = perm ? perm[b] : b;
        double er = 1.0, ei = 0.0;
        if (phase) antiphase(S, d0 + (size_t)b * S * 2, dx + (size_t)p * S * 2, &er, &ei);
        /* ox += w*eith*dx : (double*complex)*complex */
        double wr = iw->weight[i] * er, wi = iw->weight[i] * ei;
        if (!phase) { wr = iw->weight[i]; wi = 0.0; }
        for (uint32_t s = 0; s < S; ++s) {
          double xr = dx[((size_t)p * S + s) * 2], xi = dx[((size_t)p * S + s) * 2 + 1];
          if (phase) {
            out[((size_t)b * S + s) * 2] += wr * xr - wi * xi;
            out[((size_t)b * S + s) * 2 + 1] += wr * xi + wi * xr;
          } else { /* double * complex */
            out[((size_t)b * S + s) * 2] += wr * xr;
            out[((size_t)b * S + s) * 2 + 1] += wr * xi;
          }
        }
      }
    }
  } else {
    const double* base = (const double*)d->data;
    memset(out, 0, row * sizeof(double));
    for (int i = 0; i < iw->n; ++i) {
      const double* dx = base + (size_t)iw->vertex[i] * row;
      const uint32_t* perm = perm_row(data, is_cube, cellidx, iw->slot[0], iw->slot[i]);
      for (uint32_t b = 0; b < B; ++b) {
        uint32_t p = perm ? perm[b] : b;
        for (uint32_t s = 0; s < S; ++s) {
          if (phase) /* real data: antiphase == T(1): w*1*x (utilities.tpp:529) */
            out[(size_t)b * S + s] += iw->weight[i] * 1.0 * dx[(size_t)p * S + s];
          else
            out[(size_t)b * S + s] += iw->weight[i] * dx[(size_t)p * S + s];
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * rotate_in_place  (interpolator.hpp:386-428) for one Q
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double re, im; } cx;
static cx cmul(cx a, cx b) { cx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }

/* generic "matrix (double) times vector/matrix of T" helpers; T real or complex stored as doubles */
static void mat_vec_T(double* out, const double* R, const double* v, int cplx) {
  int w = cplx ? 2 : 1;
  for (int i = 0; i < 3; ++i) {
    for (int c = 0; c < w; ++c) out[i * w + c] = 0.0;
    for (int k = 0; k < 3; ++k)
      for (int c = 0; c < w; ++c) out[i * w + c] += R[i * 3 + k] * v[k * w + c];
  }
}
/* C = A(T) * B(double)  (mul_mat_mat with complex A, real B: utilities.tpp:62-67) */
static void matT_matd(double* C, const double* A, const double* B, int cplx) {
  int w = cplx ? 2 : 1;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      for (int c = 0; c < w; ++c) C[(i * 3 + j) * w + c] = 0.0;
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < w; ++c) C[(i * 3 + j) * w + c] += A[(i * 3 + k) * w + c] * B[k * 3 + j];
    }
}
/* C = A(double) * B(T) */
static void matd_matT(double* C, const double* A, const double* B, int cplx) {
  int w = cplx ? 2 : 1;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      for (int c = 0; c < w; ++c) C[(i * 3 + j) * w + c] = 0.0;
      for (int k = 0; k < 3; ++k)
        for (int c = 0; c < w; ++c) C[(i * 3 + j) * w + c] += A[i * 3 + k] * B[(k * 3 + j) * w + c];
    }
}
static void rot_as_double(double* out, const int32_t* R, int transpose) {
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) out[a * 3 + b] = (double)(transpose ? R[b * 3 + a] : R[a * 3 + b]);
}
static int is_identity(const int32_t* R) {
  static const int32_t E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  return memcmp(R, E, sizeof(E)) == 0;
}
static int det3i(const int32_t* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* e_iqd (interpolator_gamma.tpp:18-32) */
static cx e_iqd(const double* q, const double* dvec) {
  double dot = 0.0;
  for (int k = 0; k < 3; ++k) dot += q[k] * dvec[k];
  double th = TWO_PI * dot;
  cx r = {cos(th), sin(th)}; /* std::exp(i*th): exp(0)*(cos,sin) */
  return r;
}

static int rotate_one(const b200_interp_desc_t* d, const b200_data_tables_t* data, const b200_bz_tables_t* bz,
                      const double* q_ir, int ridx, int invridx, double* x) {
  const uint32_t B = d->branches, S = span_of(d);
  const uint32_t no0 = d->elements[0], no1 = d->elements[1] / 3u, no2 = d->elements[2] / 9u;
  const int cplx = d->is_complex, w = cplx ? 2 : 1;
  const int G = bz->n_ops;
  int lu = d->length_unit, rl = d->rotates_like;
  int kind; /* 0 real, 1 recip, 2 axial, 3 gamma(int R), 4 gamma(cart R) */
  if (lu == B200_LEN_REAL_LATTICE) {
    if (rl == B200_ROT_VECTOR) kind = 0;
    else if (rl == B200_ROT_PSEUDOVECTOR) kind = 2;
    else if (rl == B200_ROT_GAMMA) kind = 3;
    else return B200_E_UNSUPPORTED;
  } else if (lu == B200_LEN_RECIPROCAL_LATTICE) {
    if (rl == B200_ROT_VECTOR) kind = 1;
    else return B200_E_UNSUPPORTED;
  } else if (lu == B200_LEN_ANGSTROM) {
    if (rl == B200_ROT_GAMMA) kind = 4;
    else return B200_E_UNSUPPORTED;
  } else
    return B200_E_UNSUPPORTED;
  if (kind >= 3 && !cplx) return B200_E_UNSUPPORTED; /* "RotatesLike == Gamma requires complex valued data!" */
  if (no1 == 0 && no2 == 0) return 0;                /* pure scalars: nothing to do */
  const int32_t* R = bz->rotations + 9 * ridx;
  const int32_t* iR = bz->rotations + 9 * invridx;
  double Rd[9], iRd[9];
  if (kind <= 2) {
    if (is_identity(R)) return 0; /* interpolator_real.tpp:41 */
    if (kind == 0) { rot_as_double(Rd, R, 0); rot_as_double(iRd, iR, 0); }
    if (kind == 1) { rot_as_double(Rd, R, 1); rot_as_double(iRd, iR, 1); }
    if (kind == 2) { rot_as_double(Rd, R, 0); rot_as_double(iRd, iR, 0); }
    double detR = (double)det3i(R);
    double tv[6], tm[18];
    for (uint32_t b = 0; b < B; ++b) {
      size_t o = (size_t)b * S + no0;
      for (uint32_t v = 0; v < no1; ++v) {
        if (kind == 2) { /* interpolator_axial.tpp:47-51: det(R) * R^-1 v */
          mat_vec_T(tv, iRd, x + o * w, cplx);
          for (int j = 0; j < 3 * w; ++j) x[o * w + j] = detR * tv[j];
        } else { /* real: R v (interpolator_real.tpp:46-50); recip: R^T v (interpolator_recip.tpp) */
          mat_vec_T(tv, Rd, x + o * w, cplx);
          for (int j = 0; j < 3 * w; ++j) x[o * w + j] = tv[j];
        }
        o += 3;
      }
      for (uint32_t m = 0; m < no2; ++m) {
        if (kind == 2) { /* R^-1 M R */
          matT_matd(tm, x + o * w, Rd, cplx);
          matd_matT(x + o * w, iRd, tm, cplx);
        } else { /* R M R^-1 (with transposes for recip) */
          matT_matd(tm, x + o * w, iRd, cplx);
          matd_matT(x + o * w, Rd, tm, cplx);
        }
        o += 9;
      }
    }
    return 0;
  }
  /* Gamma (interpolator_gamma.tpp:49-139) */
  if (!data->gamma_F0 || !data->gamma_vidx || !data->gamma_vectors) return B200_E_INVALID;
  uint32_t Nmat = (uint32_t)(sqrt((double)no2)) / 3u;
  if (no2 != 9 * Nmat * Nmat) return 0; /* "requires NxN 3x3 tensors" -> returns false, data untouched */
  if (kind == 3) {
    rot_as_double(Rd, R, 0);
    rot_as_double(iRd, iR, 0);
  } else {
    memcpy(Rd, data->rot_cart + 9 * ridx, sizeof(Rd));
    memcpy(iRd, data->rot_cart + 9 * invridx, sizeof(iRd));
  }
  size_t need = (size_t)(no1 * 3u > no2 * 9u ? no1 * 3u : no2 * 9u);
  /* zero-initialised like the reference's std::vector work array (:74,101,117): the elements behind the Nmat x Nmat rotated
   * matrices are copied back from it as they are -- leftovers of the vector pass, zero beyond */
  cx* tA = (cx*)calloc(need ? need : 1, sizeof(cx));
  cx* xc = (cx*)x;
  for (uint32_t b = 0; b < B; ++b) {
    size_t o = (size_t)b * S + no0;
    if (no1 > 0) {
      size_t o0 = o;
      for (uint32_t k = 0; k < no1; ++k) {
        double t0[6];
        mat_vec_T(t0, iRd, (const double*)(xc + o), 1);
        size_t v0 = 3u * data->gamma_F0[(size_t)k * G + invridx];
        cx ph = e_iqd(q_ir, data->gamma_vectors + 3 * (size_t)data->gamma_vidx[(size_t)k * G + invridx]);
        for (int j = 0; j < 3; ++j) {
          cx t = {t0[2 * j], t0[2 * j + 1]};
          tA[v0 + j] = cmul(ph, t);
        }
        o += 3;
      }
      for (uint32_t j = 0; j < no1 * 3u; ++j) xc[o0 + j] = tA[j];
    }
    if (no2 > 0) {
      for (uint32_t n = 0; n < Nmat; ++n) {
        cx Rph = e_iqd(q_ir, data->gamma_vectors + 3 * (size_t)data->gamma_vidx[(size_t)n * G + ridx]);
        uint32_t v = data->gamma_F0[(size_t)n * G + ridx];
        for (uint32_t m = 0; m < Nmat; ++m) {
          cx iRph = e_iqd(q_ir, data->gamma_vectors + 3 * (size_t)data->gamma_vidx[(size_t)m * G + invridx]);
          uint32_t k = data->gamma_F0[(size_t)m * G + invridx];
          double t0[18], t1[18];
          matT_matd(t0, (const double*)(xc + o + 9u * (n * Nmat + m)), Rd, 1);
          matd_matT(t1, iRd, t0, 1);
          cx pp = cmul(Rph, iRph);
          for (int j = 0; j < 9; ++j) {
            cx t = {t1[2 * j], t1[2 * j + 1]};
            tA[(v * Nmat + k) * 9u + j] = cmul(pp, t);
          }
        }
      }
      for (uint32_t j = 0; j < no2 * 9u; ++j) xc[o + j] = tA[j];
    }
  }
  free(tA);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * public entry points
 * ---------------------------------------------------------------------------------------------- */
static void probe_store(b200_probe_t* p, size_t i, const double* q, const double* x, const int32_t* tau, int ridx,
                        int invridx, const iw_t* iw, uint32_t st) {
  if (!p) return;
  if (p->q_ir && q) memcpy(p->q_ir + 3 * i, q, 3 * sizeof(double));
  if (p->x_ir && x) memcpy(p->x_ir + 3 * i, x, 3 * sizeof(double));
  if (p->tau && tau) memcpy(p->tau + 3 * i, tau, 3 * sizeof(int32_t));
  if (p->ridx) p->ridx[i] = ridx;
  if (p->invridx) p->invridx[i] = invridx;
  if (iw) {
    if (p->cell) p->cell[i] = iw->cell;
    if (p->tet) p->tet[i] = iw->tet;
    if (p->n_vert) p->n_vert[i] = iw->n;
    for (int j = 0; j < 8; ++j) {
      if (p->vertex) p->vertex[8 * i + j] = j < iw->n ? iw->vertex[j] : 0xffffffffu;
      if (p->weight) p->weight[8 * i + j] = j < iw->n ? iw->weight[j] : 0.0;
    }
  }
  if (p->status) p->status[i] = st;
}

/* ir: 0 moveinto (bz_move.cpp:103-163), 1 ir_moveinto (:165-296), 2 ir_moveinto_wedge (:299-356: the wedge rotation of Q
 * itself, no translation), 3 isinside (bz.hpp:631-640: status bit only, never an error) */
int oracle_moveinto(const b200_bz_tables_t* bz, const double* Q, size_t nQ, int ir, b200_probe_t* probe) {
  int rc = 0;
  for (size_t i = 0; i < nQ; ++i) {
    double q[3], x[3];
    int32_t tau[3] = {0, 0, 0};
    int r = bz->identity_index, ri = bz->identity_index;
    uint32_t st = 0;
    if (ir == 3) {
      memcpy(q, Q + 3 * i, sizeof(q));
      st = isinside_one(bz, q) ? 0u : (uint32_t)B200_ST_OUTSIDE_BZ;
      matvec_dd(x, bz->to_xyz, q);
      probe_store(probe, i, q, x, tau, r, ri, NULL, st);
      continue;
    }
    if (ir == 2) {
      st = wedge_rotate_one(bz, Q + 3 * i, q, &r, &ri);
      /* a point no operation places: the reference counts it but then only re-tests the OUTPUT row, which still holds its
       * zero initialisation and passes (bz_move.cpp:348-354) -- q = 0, operation 0, no error */
      if (st & B200_ST_OUTSIDE_WEDGE) q[0] = q[1] = q[2] = 0.0;
    } else {
      st = ir ? ir_moveinto_one(bz, Q + 3 * i, q, tau, &r, &ri) : moveinto_one(bz, Q + 3 * i, q, tau);
    }
    matvec_dd(x, bz->to_xyz, q);
    probe_store(probe, i, q, x, tau, r, ri, NULL, st);
    if ((st & B200_ST_OUTSIDE_BZ) && !rc) rc = B200_E_OUTSIDE_BZ;
    if ((st & B200_ST_OUTSIDE_WEDGE) && ir != 2 && !rc) rc = B200_E_OUTSIDE_WEDGE;
  }
  return rc;
}

/* mode: 1 = ir_interpolate_at (bz_trellis.hpp:153-196), 0 = interpolate_at (bz_trellis.hpp:105-121) */
int oracle_interpolate_at(int kind, const b200_bz_tables_t* bz, const void* structure, const b200_data_tables_t* data,
                          const double* Q, size_t nQ, uint32_t flags, int ir, void* vals_out, void* vecs_out,
                          b200_probe_t* probe) {
  const b200_trellis_tables_t* tr = (const b200_trellis_tables_t*)structure;
  const size_t vrow = (size_t)data->values.branches * span_of(&data->values) * (data->values.is_complex ? 2 : 1);
  const size_t wrow = (size_t)data->vectors.branches * span_of(&data->vectors) * (data->vectors.is_complex ? 2 : 1);
  int rc = 0;
  for (size_t i = 0; i < nQ; ++i) {
    double q[3], x[3];
    int32_t tau[3] = {0, 0, 0};
    int r = bz->identity_index, ri = bz->identity_index;
    uint32_t st = 0;
    if (flags & B200_FLAG_NO_MOVE) {
      memcpy(q, Q + 3 * i, sizeof(q));
    } else if (ir) {
      st = ir_moveinto_one(bz, Q + 3 * i, q, tau, &r, &ri);
    } else {
      st = moveinto_one(bz, Q + 3 * i, q, tau);
    }
    matvec_dd(x, bz->to_xyz, q);
    iw_t iw;
    if (kind == B200_GRID_TRELLIS) st |= trellis_locate(tr, x, &iw);
    else if (kind == B200_GRID_NEST) st |= nest_locate((const b200_nest_tables_t*)structure, x, &iw);
    else if (kind == B200_GRID_MESH) st |= mesh_locate((const b200_mesh_tables_t*)structure, x, &iw);
    else return B200_E_UNSUPPORTED;
    probe_store(probe, i, q, x, tau, r, ri, &iw, st);
    double* vo = (double*)vals_out + i * vrow;
    double* wo = (double*)vecs_out + i * wrow;
    if (st & (B200_ST_NOT_FOUND | B200_ST_OUTSIDE_BZ | B200_ST_OUTSIDE_WEDGE)) {
      if (!rc) rc = (st & B200_ST_OUTSIDE_BZ) ? B200_E_OUTSIDE_BZ : (st & B200_ST_OUTSIDE_WEDGE) ? B200_E_OUTSIDE_WEDGE : B200_E_NOT_FOUND;
      memset(vo, 0, vrow * sizeof(double));
      memset(wo, 0, wrow * sizeof(double));
      continue;
    }
    int is_cube = kind == B200_GRID_TRELLIS && tr->node_type[iw.cell] == B200_NODE_CUBE;
    uint32_t cellidx = is_cube ? tr->node_index[iw.cell] : (uint32_t)iw.tet;
    interpolate_one(&data->values, data, &iw, is_cube, cellidx, 0, vo);
    interpolate_one(&data->vectors, data, &iw, is_cube, cellidx, 1, wo);
    if (ir) {
      int e1 = rotate_one(&data->values, data, bz, q, r, ri, vo);
      int e2 = rotate_one(&data->vectors, data, bz, q, r, ri, wo);
      if (e1 && !rc) rc = e1;
      if (e2 && !rc) rc = e2;
    }
  }
  return rc;
}
