// locate.cu -- stage 1 of the path, one thread per Q:
//   BrillouinZone::ir_moveinto  (bz_move.cpp:14-53,103-163,165-296)  -> q_ir, tau, Ridx, invRidx
//   LVec::xyz                   (array_lvec_methods.tpp:40-52)       -> x = B q
//   PolyTrellis::indices_weights (trellis_poly.hpp:245-255,382-434; trellis_node.hpp:130-149,273-364)
//                                                                     -> (vertex, weight) list
//
// This file is compiled with -fmad=false: the reference is x86-64 baseline code (no fused
// multiply-add), and every branch decision below is taken on numbers computed in the reference's
// own operation order, so tau / R / node / tetrahedron / weights come out bit-identical.
// The one place where a cheaper formula is used (first-Brillouin-zone inside test through a
// pre-multiplied plane covector) is *certified*: when the cheap value lies within its rounding bound
// of the decision threshold the reference's lattice-aware triple product is evaluated instead.
#include "device_tables.cuh"
#include "brille_b200.h"

namespace b200 {

#define TWO_PI 6.283185307179586476925286766559005768394338798750211641949889
constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ bool approx_eq(double a, double b, double rel, double abs_) {
  double x = fabs(a - b);  // approx_float.hpp:171-188
  return x <= abs_ + rel * fabs(a + b) || x < 2.2250738585072014e-308;
}

__device__ __forceinline__ void matvec(double* c, const double* A, const double* b) {
  // utilities.tpp:39-44 : C[i] = ((0 + A0*b0) + A1*b1) + A2*b2
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = ((0.0 + A[i * 3] * b[0]) + A[i * 3 + 1] * b[1]) + A[i * 3 + 2] * b[2];
}

// reference-exact lattice-aware triple product (geometry.hpp:118-138 through the LVec cross/star/dot of
// array_functions.hpp:190-201,246-290 and array_lvec_methods.tpp:53-67); only used when the fast value
// is within its error bound of the tolerance threshold
__device__ __noinline__ double exact_face_min(const double* recip_metric, const double* real_metric, double recip_volume,
                                               int nf, const double (*pa)[3], const double (*pb)[3],
                                               const double (*pc)[3], const double* q) {
  double m = 0.0;
  const double s = recip_volume / TWO_PI;
  for (int f = 0; f < nf; ++f) {
    double u[3], v[3], w[3], cr[3], st[3], tmp[3];
    for (int i = 0; i < 3; ++i) {
      u[i] = pa[f][i] - q[i];
      v[i] = pb[f][i] - q[i];
      w[i] = pc[f][i] - q[i];
    }
    cr[0] = v[1] * w[2] - v[2] * w[1];
    cr[1] = v[2] * w[0] - v[0] * w[2];
    cr[2] = v[0] * w[1] - v[1] * w[0];
    for (int i = 0; i < 3; ++i) cr[i] *= s;
    matvec(st, real_metric, cr);
    for (int i = 0; i < 3; ++i) st[i] /= TWO_PI;
    matvec(tmp, recip_metric, u);
    double o = ((0.0 + tmp[0] * st[0]) + tmp[1] * st[1]) + tmp[2] * st[2];
    if (f == 0 || o < m) m = o;
  }
  return m;
}

// point_inside_all_planes (geometry.hpp:412-418) with tolerance (float_tolerance, approx_tolerance)
__device__ __forceinline__ bool inside_planes(const BZDev& bz, bool working, const double* q, double eps) {
  const double(*pa)[3] = working ? bz.pa : bz.ca;
  const double(*pm)[3] = working ? bz.pm : bz.cm;
  double m = 1e300;
  for (int f = 0; f < bz.n_faces; ++f) {
    double o = ((pa[f][0] - q[0]) * pm[f][0] + (pa[f][1] - q[1]) * pm[f][1]) + (pa[f][2] - q[2]) * pm[f][2];
    m = fmin(m, o);
  }
  const double thr = -bz.cfg_abs;
  if (m > thr + eps) return true;
  if (m < thr - eps - 4.0 * bz.cfg_rel * bz.cfg_abs) return false;
  // ambiguous: evaluate exactly what the reference evaluates
  double e = working ? exact_face_min(bz.w_recip_metric, bz.w_real_metric, bz.w_recip_volume, bz.n_faces, bz.pa, bz.pb, bz.pc, q)
                     : exact_face_min(bz.o_recip_metric, bz.o_real_metric, bz.o_recip_volume, bz.n_faces, bz.ca, bz.cb, bz.cc, q);
  return e > 0 || approx_eq(e, 0.0, bz.cfg_rel, bz.cfg_abs);
}

// _inside_wedge_outer (bz.hpp:757-763; Array2::all array2.tpp:664-672)
__device__ __forceinline__ bool inside_wedge(const BZDev& bz, const double* q) {
  const int K = bz.n_wedge;
  if (K == 0) return true;
  if (bz.no_ir_mirroring) {
    for (int k = 0; k < K; ++k) {
      double d = ((0.0 + bz.gw[k][0] * q[0]) + bz.gw[k][1] * q[1]) + bz.gw[k][2] * q[2];
      if (!(approx_eq(d, 0.0, bz.cfg_rel, bz.cfg_abs) || d > 0.0)) return false;
    }
    return true;
  }
  bool all_le = true, all_ge = true;  // le_ge drops the tolerances (array2.tpp:666-667)
  for (int k = 0; k < K; ++k) {
    double d = ((0.0 + bz.gw[k][0] * q[0]) + bz.gw[k][1] * q[1]) + bz.gw[k][2] * q[2];
    bool z = approx_eq(d, 0.0, bz.def_rel, bz.def_abs);
    if (!(z || d < 0.0)) all_le = false;
    if (!(z || d > 0.0)) all_ge = false;
  }
  return all_le || all_ge;
}

// moveinto for one Q: part_moveinto_prim (bz_move.cpp:14-53) between the two lattice transforms
__device__ __forceinline__ uint32_t moveinto_one(const BZDev& bz, double eps_w, double eps_o, const double* Q, double* qo, int* tauo) {
  double q[3];
  int tau[3], last[3];
  {
    double Qp[3];
    if (bz.transform_needed) {
      matvec(Qp, bz.P6t, Q);  // transform.hpp:182-190
      for (int i = 0; i < 3; ++i) Qp[i] /= 6.0;
    } else {
      for (int i = 0; i < 3; ++i) Qp[i] = Q[i];
    }
    for (int i = 0; i < 3; ++i) {
      double r = round(Qp[i]);  // half away from zero, array2.tpp:498-505
      tau[i] = (int)r;
      q[i] = Qp[i] - (double)tau[i];
      last[i] = tau[i];
    }
  }
  const int F = bz.n_faces;
  int count = 0;
  bool ended_inside = false;
  while (count++ < F && !(ended_inside = inside_planes(bz, true, q, eps_w))) {
    double tmp[3];
    matvec(tmp, bz.w_recip_metric, q);  // same_lattice_dot: (G q) . n
    int max_nm = 0, max_at = 0;
    double d_at = 0.0;
    for (int j = 0; j < F; ++j) {
      double d = ((0.0 + tmp[0] * bz.normals[j][0]) + tmp[1] * bz.normals[j][1]) + tmp[2] * bz.normals[j][2];
      // N = round(d/|tau_j|) (bz_move.cpp:30): multiply by the reciprocal, and divide exactly only when the quotient is
      // within 1e-9 of a rounding boundary (half-integers) where the last bit could matter
      double quo = d * bz.inv_tau_lens[j];
      if (quo < 0.499999) continue;  // N = round(d/|tau_j|) <= 0 for certain: the face plays no part (bz_move.cpp:31-34)
      const double fr = quo - floor(quo);
      if (fabs(fr - 0.5) < 1e-9) quo = d / bz.tau_lens[j];
      int N = (int)round(quo);
      if (N > 0 && N >= max_nm) {
        bool ok = (0 == max_nm);
        if (!ok) {
          // norm(taus.view(j)+last_shift) > 0 (bz_move.cpp:38): the norm of an integer lattice vector is
          // zero iff the vector is zero
          bool nz = (bz.taus[j][0] + last[0]) != 0 || (bz.taus[j][1] + last[1]) != 0 || (bz.taus[j][2] + last[2]) != 0;
          ok = nz && d > d_at;
        }
        if (ok) {
          max_at = j;
          max_nm = N;
          d_at = d;
        }
      }
    }
    if (max_nm > 0) {
      for (int i = 0; i < 3; ++i) {
        int t = bz.taus[max_at][i];
        q[i] -= (double)t * (double)max_nm;
        tau[i] += t * max_nm;
        last[i] = t * max_nm;
      }
    }
  }
  if (bz.transform_needed) {  // transform.hpp:222-230
    matvec(qo, bz.invPt, q);
    for (int i = 0; i < 3; ++i)
      tauo[i] = bz.invPt_i[i * 3] * tau[0] + bz.invPt_i[i * 3 + 1] * tau[1] + bz.invPt_i[i * 3 + 2] * tau[2];
  } else {
    for (int i = 0; i < 3; ++i) {
      qo[i] = q[i];
      tauo[i] = tau[i];
    }
  }
  // bz_move.cpp:149: every q is re-tested against the conventional-lattice planes.  Without a primitive transform these
  // are the very planes and the very q of the last test of the loop above, whose outcome is known when the loop ended
  // because the point was inside.
  if (!bz.transform_needed && ended_inside) return 0u;
  return inside_planes(bz, false, qo, eps_o) ? 0u : (uint32_t)B200_ST_OUTSIDE_BZ;
}

// ---------------------------------------------------------------------------------------------------
// trellis
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_bin(const double* k, int n, double x, double k0, double inv) {
  // trellis_poly.hpp:67-73: index d of the first knot > x (the reference scans; the knots ascend).  The knots of a trellis are
  // (nearly) uniform (trellis_poly.hpp:635-641): d - 1 is guessed from the spacing and corrected by comparing with the knots
  // themselves, so the result is the scan's for any ascending knot vector; usually two loads instead of a binary search.
  int g = (int)((x - k0) * inv);
  g = max(-1, min(g, n - 1));
  if (x != x) g = n - 1;  // (no knot compares greater than NaN: the scan ends at n)
  while (g + 1 < n && k[g + 1] <= x) ++g;
  while (g >= 0 && k[g] > x) --g;
  int d = g + 1;
  if (d > n - 1 && x < k[0]) d = 0;
  return d > 0 ? d - 1 : d;
}
__device__ __forceinline__ int on_boundary(const double* k, int n, double x, int i, double rel, double abs_) {
  if (i + 2 < n && approx_eq(k[i + 1], x, rel, abs_)) return 1;  // trellis_poly.hpp:75-81
  if (i > 0 && approx_eq(k[i], x, rel, abs_)) return -1;
  return 0;
}
__device__ __forceinline__ bool sub_ok(const TrellisDev& t, const int* sub) {
  for (int d = 0; d < 3; ++d)
    if (sub[d] < 0 || sub[d] >= t.n_knots[d] - 1) return false;
  const int n0 = t.n_knots[0] - 1, n1 = t.n_knots[1] - 1;
  uint8_t ty = t.node_type[sub[0] + n0 * (sub[1] + n1 * sub[2])];
  return !(ty == B200_NODE_NULL || ty == B200_NODE_ASSUMED_NULL || ty == B200_NODE_FOUND_NULL);
}

__device__ __forceinline__ double orient3d_plain(const double* a, const double* b, const double* c, const double* d) {
  // geometry.hpp:130-135 with bare-array dot/cross: (a-d).((b-d)x(c-d))
  double u0 = a[0] - d[0], u1 = a[1] - d[1], u2 = a[2] - d[2];
  double v0 = b[0] - d[0], v1 = b[1] - d[1], v2 = b[2] - d[2];
  double w0 = c[0] - d[0], w1 = c[1] - d[1], w2 = c[2] - d[2];
  double c0 = v1 * w2 - v2 * w1;
  double c1 = v2 * w0 - v0 * w2;
  double c2 = v0 * w1 - v1 * w0;
  return ((0.0 + u0 * c0) + u1 * c1) + u2 * c2;
}

// weights of x in tetrahedron tp (packed), tetrahedra_contains without the shortcut (trellis_node.hpp:326-338)
__device__ __forceinline__ double tet_weights(const double* tp, const double* x, double* w, double rel, double abs_) {
  const double* p0 = tp + 4;
  const double* p1 = tp + 7;
  const double* p2 = tp + 10;
  const double* p3 = tp + 13;
  const double vol6 = tp[16];
  w[0] = orient3d_plain(x, p1, p2, p3) / vol6;
  w[1] = orient3d_plain(p0, x, p2, p3) / vol6;
  w[2] = orient3d_plain(p0, p1, x, p3) / vol6;
  w[3] = orient3d_plain(p0, p1, p2, x) / vol6;
  bool neg = false;
  for (int j = 0; j < 4; ++j) neg |= (w[j] < 0.0 && !approx_eq(w[j], 0.0, rel, abs_));
  if (neg) return fmin(fmin(w[0], w[1]), fmin(w[2], w[3]));
  return 0.0;
}

// Result of the point location of one Q, written straight to the per-point output rows.  The common case (every
// corner of the cell carries weight) is stored with vector stores from registers; only points on a face / edge / vertex
// of their cell take the compacting path.
struct EmitOut {
  uint32_t* v;     // (8) vertex indices, emission order
  double* w;       // (8) weights
  uint64_t slots;  // byte j = cell corner of emitted vertex j
  int n;
};

// 32 bytes to a 32-byte aligned GLOBAL address with one store: the sector is written whole, so L2 never has to fetch it to
// merge a partial write (four 16-byte stores per weight row cost a 64-byte read per point)
__device__ __forceinline__ void st32(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __noinline__ void emit_compact(EmitOut& e, const uint32_t* vi, const double* w, const bool* keep, int count, bool cube) {
  e.n = 0;
  e.slots = 0;
  for (int i = 0; i < count; ++i)
    if (keep[i]) {
      const int slot = cube ? 7 - i : i;  // trellis_node.hpp:143-147 emits vertex_indices[7-i]
      e.v[e.n] = vi[slot];
      e.w[e.n] = w[i];
      e.slots |= (uint64_t)slot << (8 * e.n);
      ++e.n;
    }
  for (int j = e.n; j < 8; ++j) {
    e.v[j] = 0xffffffffu;
    e.w[j] = 0.0;
  }
}

// part 1 of the trellis location: the node of x (bins + neighbour search, trellis_poly.hpp:67-81,382-424)
__device__ __forceinline__ uint32_t trellis_find_node(const BZDev& bz, const TrellisDev& t, const double* knots, const double* x,
                                                      uint32_t& cell) {
  uint32_t st = 0;
  cell = 0xffffffffu;
  int sub[3];
  for (int d = 0; d < 3; ++d) sub[d] = find_bin(knots + t.knot_offset[d], t.n_knots[d], x[d], t.knot0[d], t.knot_inv[d]);
  bool bad = !sub_ok(t, sub);
  if (bad) {  // trellis_poly.hpp:391-424
    int close[3], num_close = 0, ns[3] = {sub[0], sub[1], sub[2]};
    for (int i = 0; i < 3; ++i) {
      close[i] = on_boundary(knots + t.knot_offset[i], t.n_knots[i], x[i], sub[i], bz.def_rel, bz.def_abs);
      num_close += close[i] != 0;
    }
    if (num_close > 0)
      for (int i = 0; i < 3 && bad; ++i)
        if (close[i]) {
          ns[0] = sub[0]; ns[1] = sub[1]; ns[2] = sub[2];
          ns[i] += close[i];
          bad = !sub_ok(t, ns);
        }
    if (bad && num_close > 1)
      for (int i = 0; i < 3 && bad; ++i)
        if (close[i])
          for (int j = 0; j < 3 && bad; ++j)
            if (close[j]) {
              ns[0] = sub[0]; ns[1] = sub[1]; ns[2] = sub[2];
              ns[i] += close[i];
              ns[j] += close[j];
              bad = !sub_ok(t, ns);
            }
    if (bad && num_close > 2) {
      for (int i = 0; i < 3; ++i) ns[i] = sub[i] + close[i];
      bad = !sub_ok(t, ns);
    }
    if (bad) return st | B200_ST_NOT_FOUND;  // reference: null-node access -> std::logic_error
    sub[0] = ns[0]; sub[1] = ns[1]; sub[2] = ns[2];
    st |= B200_ST_NEIGHBOUR;
  }
  const int n0 = t.n_knots[0] - 1, n1 = t.n_knots[1] - 1;
  cell = (uint32_t)(sub[0] + n0 * (sub[1] + n1 * sub[2]));
  return st;
}

// part 2: vertices and weights of x inside node `cell` (trellis_node.hpp:130-149,273-364)
__device__ __forceinline__ uint32_t trellis_in_node(const BZDev& bz, const TrellisDev& t, const double* x, uint32_t cell, EmitOut& e,
                                                    int& tet) {
  uint32_t st = 0;
  e.n = 0;
  e.slots = 0;
  tet = -1;
  const uint32_t payload = t.node_index[cell];
  uint4* vo = reinterpret_cast<uint4*>(e.v);
  if (t.node_type[cell] == B200_NODE_CUBE) {
    // CubeNode::indices_weights (trellis_node.hpp:130-149)
    const double2* cp = reinterpret_cast<const double2*>(t.cube_pack + 24 * (size_t)payload);
    const uint4* vip = reinterpret_cast<const uint4*>(t.cube_vertices + 8 * (size_t)payload);
    double c[24];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const double2 v = cp[i];
      c[2 * i] = v.x;
      c[2 * i + 1] = v.y;
    }
    const double vol = (fabs(c[0] - c[21]) * fabs(c[1] - c[22])) * fabs(c[2] - c[23]);
    double w[8];
    bool keep[8];
    bool all = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      w[i] = ((fabs(x[0] - c[3 * i]) * fabs(x[1] - c[3 * i + 1])) * fabs(x[2] - c[3 * i + 2])) / vol;
      keep[i] = !approx_eq(w[i], 0.0, bz.def_rel, bz.def_abs) && w[i] > 0.0;  // w.is(gt, 0.)
      all &= keep[i];
    }
    const uint4 va = vip[0], vb = vip[1];
    if (all) {
      vo[0] = make_uint4(vb.w, vb.z, vb.y, vb.x);  // vertex_indices[7], [6], [5], [4]
      vo[1] = make_uint4(va.w, va.z, va.y, va.x);  // vertex_indices[3], [2], [1], [0]
      st32(e.w, w[0], w[1], w[2], w[3]);
      st32(e.w + 4, w[4], w[5], w[6], w[7]);
      e.n = 8;
      e.slots = 0x0001020304050607ull;  // byte j = 7 - j
    } else {
      const uint32_t vi[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
      emit_compact(e, vi, w, keep, 8, true);
    }
  } else {
    // PolyNode::indices_weights (trellis_node.hpp:273-308), should_contain == true
    const uint32_t t0 = t.poly_offsets[payload], t1 = t.poly_offsets[payload + 1];
    double w[4];
    double best = 0.0;
    uint32_t best_at = t0;
    bool have_best = false;
    int found = -1;
    // Two phases per block of <= 64 tetrahedra, so that the expensive part is not executed once per loop trip of the
    // slowest lane of the warp: (1) the cheap circumsphere test of tetrahedra_might_contain for every tetrahedron
    // (:349-364), remembering the candidates in a bit mask; (2) the four weights of the candidates, in storage order,
    // until one contains the point (first accepted wins, :284-290).  `best` keeps max_element's first-maximum rule.
    for (uint32_t base = t0; base < t1 && found < 0; base += 64) {
      const uint32_t nblk = min(64u, t1 - base);
      unsigned long long cand = 0ull;
      for (uint32_t k = 0; k < nblk; ++k) {
        const double* tp = t.tet_pack + (size_t)TET_PACK * (base + k);
        const double2 c01 = reinterpret_cast<const double2*>(tp)[0];
        const double2 c23 = reinterpret_cast<const double2*>(tp)[1];
        const double v0 = c01.x - x[0], v1 = c01.y - x[1], v2 = c23.x - x[2];
        const double d2 = ((0.0 + v0 * v0) + v1 * v1) + v2 * v2;
        if (d2 < c23.y || approx_eq(d2, c23.y, bz.def_rel, bz.def_abs)) {
          cand |= 1ull << k;
        } else {
          const double mn = -d2;
          if (!have_best || mn > best || (mn == best && base + k < best_at)) {
            best = mn;
            best_at = base + k;
            have_best = true;
          }
        }
      }
      while (cand && found < 0) {
        const uint32_t k = base + (uint32_t)(__ffsll((long long)cand) - 1);
        cand &= cand - 1ull;
        const double mn = tet_weights(t.tet_pack + (size_t)TET_PACK * k, x, w, bz.def_rel, bz.def_abs);
        if (mn >= 0.0) {
          found = (int)k;
        } else if (!have_best || mn > best || (mn == best && k < best_at)) {
          best = mn;
          best_at = k;
          have_best = true;
        }
      }
    }
    if (found < 0) {
      if (t1 == t0) return st | B200_ST_NOT_FOUND;
      st |= B200_ST_FALLBACK_TET;  // trellis_node.hpp:295-306
      found = (int)best_at;
      tet_weights(t.tet_pack + (size_t)TET_PACK * best_at, x, w, bz.def_rel, bz.def_abs);
    }
    tet = found;
    const uint4 vi4 = *reinterpret_cast<const uint4*>(t.tet_vertices + 4 * (size_t)found);
    bool keep[4];
    bool all = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      keep[j] = !approx_eq(w[j], 0.0, bz.def_rel, bz.def_abs);
      all &= keep[j];
    }
    if (all) {
      vo[0] = vi4;
      vo[1] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      st32(e.w, w[0], w[1], w[2], w[3]);
      st32(e.w + 4, 0.0, 0.0, 0.0, 0.0);
      e.n = 4;
      e.slots = 0x03020100ull;
    } else {
      const uint32_t vi[4] = {vi4.x, vi4.y, vi4.z, vi4.w};
      emit_compact(e, vi, w, keep, 4, false);
    }
  }
  if (e.n < 1) st |= B200_ST_NOT_FOUND;
  return st;
}

__device__ __forceinline__ uint32_t trellis_locate(const BZDev& bz, const TrellisDev& t, const double* knots, const double* x,
                                                   EmitOut& e, uint32_t& cell, int& tet) {
  e.n = 0;
  e.slots = 0;
  tet = -1;
  const uint32_t st = trellis_find_node(bz, t, knots, x, cell);
  if (st & B200_ST_NOT_FOUND) return st;
  return st | trellis_in_node(bz, t, x, cell, e, tet);
}

// ---------------------------------------------------------------------------------------------------
// orient3d of the vendored TetGen (lib/tetgen/predicates.cxx:1997-2050): the floating-point determinant in the
// reference's operation order, returned as is when |det| > o3derrboundA*permanent; otherwise (the point is within
// ~1e-15 of the plane) re-evaluated in double-double arithmetic where the reference switches to exact expansions.
// ---------------------------------------------------------------------------------------------------
struct dd_t { double hi, lo; };
__device__ __forceinline__ dd_t dd_two_sum(double a, double b) { double s = a + b, bb = s - a; return {s, (a - (s - bb)) + (b - bb)}; }
__device__ __forceinline__ dd_t dd_two_diff(double a, double b) { double s = a - b, bb = s - a; return {s, (a - (s - bb)) - (b + bb)}; }
__device__ __forceinline__ dd_t dd_norm(double hi, double lo) { double s = hi + lo; return {s, lo - (s - hi)}; }
__device__ __forceinline__ dd_t dd_add(dd_t a, dd_t b) {
  dd_t s = dd_two_sum(a.hi, b.hi), t = dd_two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = dd_norm(s.hi, s.lo);
  s.lo += t.lo;
  return dd_norm(s.hi, s.lo);
}
__device__ __forceinline__ dd_t dd_neg(dd_t a) { return {-a.hi, -a.lo}; }
__device__ __forceinline__ dd_t dd_mul(dd_t a, dd_t b) {
  double p = a.hi * b.hi;
  dd_t r = {p, __fma_rn(a.hi, b.hi, -p)};
  r.lo += a.hi * b.lo + a.lo * b.hi;
  return dd_norm(r.hi, r.lo);
}
__device__ __noinline__ double orient3d_dd(const double* pa, const double* pb, const double* pc, const double* pd) {
  dd_t ax = dd_two_diff(pa[0], pd[0]), ay = dd_two_diff(pa[1], pd[1]), az = dd_two_diff(pa[2], pd[2]);
  dd_t bx = dd_two_diff(pb[0], pd[0]), by = dd_two_diff(pb[1], pd[1]), bz = dd_two_diff(pb[2], pd[2]);
  dd_t cx = dd_two_diff(pc[0], pd[0]), cy = dd_two_diff(pc[1], pd[1]), cz = dd_two_diff(pc[2], pd[2]);
  dd_t t1 = dd_mul(az, dd_add(dd_mul(bx, cy), dd_neg(dd_mul(cx, by))));
  dd_t t2 = dd_mul(bz, dd_add(dd_mul(cx, ay), dd_neg(dd_mul(ax, cy))));
  dd_t t3 = dd_mul(cz, dd_add(dd_mul(ax, by), dd_neg(dd_mul(bx, ay))));
  dd_t r = dd_add(dd_add(t1, t2), t3);
  return r.hi + r.lo;
}
__device__ __forceinline__ double orient3d_tetgen(const double* pa, const double* pb, const double* pc, const double* pd) {
  const double eps = 1.1102230246251565e-16;  // 2^-53 (exactinit, predicates.cxx:380-443)
  const double o3derrboundA = (7.0 + 56.0 * eps) * eps;
  const double adx = pa[0] - pd[0], ady = pa[1] - pd[1], adz = pa[2] - pd[2];
  const double bdx = pb[0] - pd[0], bdy = pb[1] - pd[1], bdz = pb[2] - pd[2];
  const double cdx = pc[0] - pd[0], cdy = pc[1] - pd[1], cdz = pc[2] - pd[2];
  const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, cdxady = cdx * ady, adxcdy = adx * cdy, adxbdy = adx * bdy, bdxady = bdx * ady;
  const double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
  const double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * fabs(adz) + (fabs(cdxady) + fabs(adxcdy)) * fabs(bdz) +
                           (fabs(adxbdy) + fabs(bdxady)) * fabs(cdz);
  const double errbound = o3derrboundA * permanent;
  if (det > errbound || -det > errbound) return det;
  return orient3d_dd(pa, pb, pc, pd);
}
// the four barycentric weights of x in a packed tetrahedron (NestLeaf::weights nest.hpp:78-88, TetTriLayer::weights
// triangulation_layers.hpp:265-288)
__device__ __forceinline__ void tetgen_weights(const double* tp, const double* x, double* w) {
  const double *p0 = tp + 4, *p1 = tp + 7, *p2 = tp + 10, *p3 = tp + 13;
  const double vol6 = tp[16];
  w[0] = orient3d_tetgen(x, p1, p2, p3) / vol6;
  w[1] = orient3d_tetgen(p0, x, p2, p3) / vol6;
  w[2] = orient3d_tetgen(p0, p1, x, p3) / vol6;
  w[3] = orient3d_tetgen(p0, p1, p2, x) / vol6;
}

// emit the weights that are not ~0, tetrahedron corner order
__device__ __forceinline__ void emit_tet(EmitOut& e, const uint32_t* vip, const double* w, double rel, double abs_) {
  const uint4 vi4 = *reinterpret_cast<const uint4*>(vip);
  bool keep[4];
  bool all = true;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    keep[j] = !approx_eq(w[j], 0.0, rel, abs_);
    all &= keep[j];
  }
  if (all) {
    reinterpret_cast<uint4*>(e.v)[0] = vi4;
    reinterpret_cast<uint4*>(e.v)[1] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    st32(e.w, w[0], w[1], w[2], w[3]);
    st32(e.w + 4, 0.0, 0.0, 0.0, 0.0);
    e.n = 4;
    e.slots = 0x03020100ull;
  } else {
    const uint32_t vi[4] = {vi4.x, vi4.y, vi4.z, vi4.w};
    emit_compact(e, vi, w, keep, 4, false);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fast path of the Nest / Mesh location.  The leaves of a Nest (the finest layer of a Mesh) tile the gridded volume, so a point
// STRICTLY inside one of them -- all four weights above FAST_MARGIN, a million times the tolerances -- has exactly one containing
// leaf, is inside every ancestor of it (every coarser tetrahedron on the way) with room to spare, and the reference's search,
// whatever its order, ends with that leaf and the weights computed below (same arithmetic: tetgen_weights).  The candidates
// come from a uniform grid of bins (BinDev::cand_*): the reference's weights are evaluated for a handful of tetrahedra instead of a descent
// through every level.  Anything else -- a point within the margin of a face, or in no candidate -- takes the reference's search
// as it is (nest_search / mesh_search).
// ---------------------------------------------------------------------------------------------------------------------
constexpr double FAST_MARGIN = 1e-9;
__device__ __forceinline__ int fast_candidate(const BinDev& b, const double* pack, bool squared_radius, const double* x, double* w) {
  if (!b.cand_offset) return -1;
  int ib[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double f = (x[d] - b.lo[d]) * b.inv[d];
    ib[d] = f > 0.0 ? (f < (double)b.n[d] ? (int)f : b.n[d] - 1) : 0;  // (NaN -> 0)
  }
  const uint32_t bin = (uint32_t)(ib[0] + b.n[0] * (ib[1] + b.n[1] * ib[2]));
  const uint32_t c0 = b.cand_offset[bin], c1 = b.cand_offset[bin + 1];
  // (measured and not kept: a bounding-box test in place of the circumsphere test with the weights of all lanes evaluated in
  // lock step -- 12 more loads and 24 min/max per candidate cost more than the weights they save: Nest 2.17 -> 3.11 ms)
  for (uint32_t c = c0; c < c1; ++c) {
    const uint32_t id = b.cand_index[c];
    const double* tp = pack + (size_t)TET_PACK * id;
    const double2 c01 = reinterpret_cast<const double2*>(tp)[0];
    const double2 c23 = reinterpret_cast<const double2*>(tp)[1];
    const double d0 = x[0] - c01.x, d1 = x[1] - c01.y, d2v = x[2] - c23.x;
    const double d2 = ((0.0 + d0 * d0) + d1 * d1) + d2v * d2v;
    // a point strictly inside a tetrahedron is strictly inside its circumsphere: candidates whose sphere (widened a little; the
    // record holds r^2 for a Nest, r for a Mesh) does not hold x cannot be the leaf looked for
    const double r2 = squared_radius ? c23.y : c23.y * c23.y;
    if (d2 > r2 * (1.0 + 1e-9)) continue;
    tetgen_weights(tp, x, w);
    const double lo = fmin(fmin(w[0], w[1]), fmin(w[2], w[3]));
    if (lo > FAST_MARGIN) return (int)id;  // strictly inside: the one containing leaf
    if (lo >= -FAST_MARGIN) return -1;     // within the margin of a face: the reference's search decides
  }
  return -1;
}

// Nest::indices_weights (nest.hpp:163-222) as the reference defines it: every node whose ancestors all contain x (circumsphere
// test, then no weight negative beyond tolerance) is visited; among the containing leaves the LAST one, in the reference's
// breadth-first order, whose weights are all > 0 wins, else the first; ~0 weights are folded into the largest one.  The tree is
// stored breadth-first, so "first" and "last" are the smallest and the largest node index: a depth-first walk with an explicit
// stack (bounded by the depth of the tree, not by how many nodes contain the point) finds the same two leaves.
__device__ __noinline__ uint32_t nest_search(const NestDev& t, const double* x, double* sw, uint32_t& node_out) {
  uint32_t sb[NEST_MAX_DEPTH], se[NEST_MAX_DEPTH];
  int depth = 0;
  sb[0] = t.child_begin[0];
  se[0] = t.child_end[0];
  int nsol = 0;
  bool have_best = false;
  uint32_t first_node = 0xffffffffu, best_node = 0;
  double fw[4] = {0, 0, 0, 0}, bw[4] = {0, 0, 0, 0}, w[4];
  while (depth >= 0) {
    if (sb[depth] >= se[depth]) {
      --depth;
      continue;
    }
    const uint32_t node = sb[depth]++;
    const double* tp = t.node_pack + (size_t)TET_PACK * node;
    const double2 c01 = reinterpret_cast<const double2*>(tp)[0];
    const double2 c23 = reinterpret_cast<const double2*>(tp)[1];
    const double d0 = x[0] - c01.x, d1 = x[1] - c01.y, d2v = x[2] - c23.x;
    const double d2 = ((0.0 + d0 * d0) + d1 * d1) + d2v * d2v;
    if (!(d2 < c23.y || approx_eq(d2, c23.y, t.rel, t.abs_))) continue;  // might_contain (:114-122)
    tetgen_weights(tp, x, w);
    bool ok = true;  // none_negative (:43-48)
#pragma unroll
    for (int j = 0; j < 4; ++j) ok &= !(w[j] < 0.0 && !approx_eq(w[j], 0.0, t.rel, t.abs_));
    if (!ok) continue;
    if (t.node_is_leaf[node]) {
      if (node < first_node) {
        first_node = node;
#pragma unroll
        for (int j = 0; j < 4; ++j) fw[j] = w[j];
      }
      if (w[0] > 0.0 && w[1] > 0.0 && w[2] > 0.0 && w[3] > 0.0 && (!have_best || node > best_node)) {
        best_node = node;
        have_best = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) bw[j] = w[j];
      }
      ++nsol;
    } else if (depth + 1 < NEST_MAX_DEPTH) {
      ++depth;
      sb[depth] = t.child_begin[node];
      se[depth] = t.child_end[node];
    }
  }
  if (nsol == 0) return B200_ST_NOT_FOUND;
  uint32_t node = first_node;
#pragma unroll
  for (int j = 0; j < 4; ++j) sw[j] = fw[j];
  if (nsol > 1 && have_best) {
    node = best_node;
#pragma unroll
    for (int j = 0; j < 4; ++j) sw[j] = bw[j];
  }
  for (int i = 0; i < 4; ++i)  // fold (:203-217)
    if (approx_eq(sw[i], 0.0, t.rel, t.abs_)) {
      int max_at = 0;
      for (int j = 0; j < 4; ++j)
        if (!approx_eq(sw[j], 0.0, t.rel, t.abs_) && sw[j] > sw[max_at]) max_at = j;
      sw[max_at] += sw[i];
    }
  node_out = node;
  return 0u;
}

__device__ __forceinline__ uint32_t nest_locate(const BZDev& bz, const GridDev& gd, const double* x, EmitOut& e, uint32_t& cell, int& tet) {
  const NestDev& t = gd.ne;
  e.n = 0;
  e.slots = 0;
  tet = -1;
  cell = 0xffffffffu;
  double sw[4];
  uint32_t node = 0;
  const int fast = fast_candidate(gd.bins, t.node_pack, true, x, sw);
  if (fast >= 0) {
    node = (uint32_t)fast;
  } else {
    const uint32_t st = nest_search(t, x, sw, node);
    if (st) return st;
  }
  emit_tet(e, t.node_vertices + 4 * (size_t)node, sw, t.rel, t.abs_);
  cell = node;
  tet = (int)node;
  return e.n < 1 ? (uint32_t)B200_ST_NOT_FOUND : 0u;
}

// TetTri::locate (triangulation_layers.hpp:401-417): first containing tetrahedron of the coarsest layer, then of the
// candidate list connections[l-1][idx] in every finer layer
__device__ __forceinline__ bool mesh_sphere(const BZDev& bz, const double* tp, const double* x) {
  const double2 c01 = reinterpret_cast<const double2*>(tp)[0];
  const double2 c23 = reinterpret_cast<const double2*>(tp)[1];
  const double d0 = x[0] - c01.x, d1 = x[1] - c01.y, d2v = x[2] - c23.x;
  const double nrm = sqrt(((0.0 + d0 * d0) + d1 * d1) + d2v * d2v);  // unsafe_might_contain (:254-256)
  return approx_eq(nrm, c23.y, bz.def_rel, bz.def_abs) || nrm < c23.y;
}
__device__ __forceinline__ bool mesh_contains(const BZDev& bz, const double* tp, const double* x, double* w) {
  tetgen_weights(tp, x, w);
  bool ok = true;  // unsafe_contains (:261-264)
#pragma unroll
  for (int j = 0; j < 4; ++j) ok &= (w[j] > 0.0 || approx_eq(w[j], 0.0, bz.def_rel, bz.def_abs));
  return ok;
}
// TetTri::locate as the reference walks it (triangulation_layers.hpp:401-417): the slow path behind fast_candidate
__device__ __noinline__ uint32_t mesh_search(const BZDev& bz, const MeshDev& t, const double* x, double* w, uint32_t& idx_out) {
  uint32_t idx = 0xffffffffu;
  for (uint32_t layer = 0; layer < t.n_layers; ++layer) {
    uint32_t found = 0xffffffffu;
    const uint32_t base = t.tet_offset[layer];
    // candidate list of this layer: all tetrahedra (layer 0) or connections[layer-1][idx]; scanned in blocks of 64 with the
    // cheap circumsphere test first and the four weights only for the survivors, in list order (first match wins)
    const uint32_t c = layer ? t.tet_offset[layer - 1] + idx : 0u;
    const uint32_t k0 = layer ? t.conn_offset[c] : 0u, k1 = layer ? t.conn_offset[c + 1] : t.tet_offset[1] - base;
    for (uint32_t kb = k0; kb < k1 && found == 0xffffffffu; kb += 64) {
      const uint32_t nblk = min(64u, k1 - kb);
      unsigned long long cand = 0ull;
      for (uint32_t k = 0; k < nblk; ++k) {
        const uint32_t ti = layer ? t.conn_index[kb + k] : kb + k;
        if (mesh_sphere(bz, t.tet_pack + (size_t)TET_PACK * (base + ti), x)) cand |= 1ull << k;
      }
      while (cand) {
        const uint32_t k = (uint32_t)(__ffsll((long long)cand) - 1);
        cand &= cand - 1ull;
        const uint32_t ti = layer ? t.conn_index[kb + k] : kb + k;
        if (mesh_contains(bz, t.tet_pack + (size_t)TET_PACK * (base + ti), x, w)) { found = ti; break; }
      }
    }
    if (found == 0xffffffffu) return B200_ST_NOT_FOUND;
    idx = found;
  }
  idx_out = idx;
  return 0u;
}

__device__ __forceinline__ uint32_t mesh_locate(const BZDev& bz, const GridDev& gd, const double* x, EmitOut& e, uint32_t& cell, int& tet) {
  const MeshDev& t = gd.me;
  e.n = 0;
  e.slots = 0;
  tet = -1;
  cell = 0xffffffffu;
  double w[4];
  uint32_t idx = 0xffffffffu;
  {  // fast path: a point strictly inside a tetrahedron of the finest layer (see fast_candidate)
    const uint32_t last = t.tet_offset[t.n_layers - 1];
    const int fast = fast_candidate(gd.bins, t.tet_pack + (size_t)TET_PACK * last, false, x, w);
    if (fast >= 0) {
      idx = (uint32_t)fast;
      emit_tet(e, t.tets + 4 * (size_t)(last + idx), w, bz.def_rel, bz.def_abs);
      cell = idx;
      tet = (int)idx;
      return e.n < 1 ? (uint32_t)B200_ST_NOT_FOUND : 0u;
    }
  }
  const uint32_t st = mesh_search(bz, t, x, w, idx);
  if (st) return st;
  emit_tet(e, t.tets + 4 * (size_t)(t.tet_offset[t.n_layers - 1] + idx), w, bz.def_rel, bz.def_abs);
  cell = idx;
  tet = (int)idx;
  return e.n < 1 ? (uint32_t)B200_ST_NOT_FOUND : 0u;
}

// in-order scan with the reference arithmetic (bz_move.cpp:262-285); taken by points within tolerance of a wedge
// plane and by Brillouin zones for which the sign-pattern lookup is not available
// (`qs`: the rounding bounds of the certified tests were derived for |q_i| <= 4 rlu, true inside any first Brillouin zone; the
// wedge rotation of an untranslated point, ir_moveinto_wedge, scales them with the point)
__device__ __noinline__ bool wedge_scan(const BZDev& bz, double* q, int& ridx, int& invridx, double qs) {
  for (int j = 0; j < bz.n_ops; ++j) {
    int verdict = 2;  // 0 outside, 1 inside, 2 ask the reference arithmetic
    if (bz.wedge_fast) {
      // (G* n_k).(R_j^T q) evaluated as (R_j G* n_k).q; certain unless within eps_wedge of the threshold
      verdict = 1;
      const double ew = bz.eps_wedge * qs;
      const double lo = -bz.cfg_abs * (1.0 + 4.0 * bz.cfg_rel) - ew, hi = -bz.cfg_abs + ew;
      for (int k = 0; k < bz.n_wedge; ++k) {
        const double d = (bz.wc[j][k][0] * q[0] + bz.wc[j][k][1] * q[1]) + bz.wc[j][k][2] * q[2];
        if (d < lo) { verdict = 0; break; }
        if (d < hi) verdict = 2;
      }
    }
    if (verdict == 0) continue;
    double qj[3];
    matvec(qj, bz.Rt[j], q);
    if (verdict == 1 || inside_wedge(bz, qj)) {
      q[0] = qj[0]; q[1] = qj[1]; q[2] = qj[2];
      invridx = j;
      ridx = bz.inverse_index[j];
      return true;
    }
  }
  return false;
}

// SPLIT (trellis only): the kernel stops after the node is found and parks the point for k_locate_in_node
template <int KIND, bool SPLIT = false>
__global__ void __launch_bounds__(128)
k_locate(const BZDev* __restrict__ bzg, GridDev gd, const double* __restrict__ Q, size_t n, uint32_t mode,
         double eps_w, double eps_o, LocateOut out, unsigned long long* __restrict__ fail_count) {
  const TrellisDev& tr = gd.tr;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BZDev& bz = *reinterpret_cast<BZDev*>(smem_raw);
  double* knots = reinterpret_cast<double*>(smem_raw + ((sizeof(BZDev) + 15) / 16) * 16);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(bzg);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&bz);
    for (int i = threadIdx.x; i < (int)(sizeof(BZDev) / 4); i += blockDim.x) dst[i] = src[i];
    const int nk = (KIND != B200_GRID_TRELLIS || (mode & MODE_NO_LOCATE)) ? 0 : tr.n_knots[0] + tr.n_knots[1] + tr.n_knots[2];
    for (int i = threadIdx.x; i < nk; i += blockDim.x) knots[i] = tr.knots[i];
  }
  __syncthreads();
  unsigned long long f_bz = 0, f_wedge = 0, f_find = 0;
  // (SPLIT) the rank returned by the bucket atomicAdd is stored one trip later, so that its latency overlaps the next point
  uint32_t pending_rank = 0;
  uint32_t* pending_at = nullptr;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double Qi[3] = {Q[3 * i], Q[3 * i + 1], Q[3 * i + 2]};
    {  // the point of the next trip on its way into L2 (the warps of a CTA are too few to hide a DRAM round trip; staging it in
       // shared memory by an asynchronous copy, as k_trellis_in_node_coop does, was measured slower here: 0.69 -> 0.73 ms)
      const size_t i_next = i + (size_t)gridDim.x * blockDim.x;
      if (i_next < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(Q + 3 * i_next));
    }
    double q[3];
    int tau[3] = {0, 0, 0};
    int ridx = bz.identity_index, invridx = bz.identity_index;
    uint32_t st = 0;
    if (mode & MODE_NO_MOVE) {
      q[0] = Qi[0]; q[1] = Qi[1]; q[2] = Qi[2];
      // BrillouinZone::isinside (bz.hpp:631-640): the conventional-lattice plane test moveinto re-checks its result with
      if ((mode & MODE_ISINSIDE) &&
          !inside_planes(bz, false, q, eps_o * fmax(1.0, 0.25 * fmax(fabs(q[0]), fmax(fabs(q[1]), fabs(q[2]))))))
        st = B200_ST_OUTSIDE_BZ;
    } else {
      if (mode & MODE_NO_TAU) {  // ir_moveinto_wedge (bz_move.cpp:299-356): the rotation search on Q itself
        q[0] = Qi[0]; q[1] = Qi[1]; q[2] = Qi[2];
      } else {
        st = moveinto_one(bz, eps_w, eps_o, Qi, q, tau);
      }
      if (mode & MODE_IR) {
        // ---- wedge rotation: bz_move.cpp:257-285 ----
        // fast path: signs of q on the distinct wedge-bounding planes -> operation index through a lookup table
        int pick = -1;
        const double qs = (mode & MODE_NO_TAU) ? fmax(1.0, 0.25 * fmax(fabs(q[0]), fmax(fabs(q[1]), fabs(q[2])))) : 1.0;
        if (bz.n_wplanes > 0) {
          unsigned mask = 0;
          bool ambiguous = false;
          const double band = bz.wband * qs;
          for (int p = 0; p < bz.n_wplanes; ++p) {
            const double d = (bz.wplane[p][0] * q[0] + bz.wplane[p][1] * q[1]) + bz.wplane[p][2] * q[2];
            ambiguous |= fabs(d) <= band;
            mask |= (d > 0.0 ? 1u : 0u) << p;
          }
          if (!ambiguous) {
            const int j = bz.wtable[mask];
            if (j != 0xff) pick = j;
          }
        }
        bool done = true;
        if (pick >= 0) {
          if (pick != bz.identity_index) {
            double qj[3];
            matvec(qj, bz.Rt[pick], q);  // exactly the reference's R_j^T q
            q[0] = qj[0]; q[1] = qj[1]; q[2] = qj[2];
            invridx = pick;
            ridx = bz.inverse_index[pick];
          }
        } else if (!inside_wedge(bz, q)) {
          done = wedge_scan(bz, q, ridx, invridx, qs);
        }
        if (!done) {
          st |= B200_ST_OUTSIDE_WEDGE;
          ridx = invridx = 0;
          // ir_moveinto_wedge: the reference leaves the zero-initialised output row of such a point and does not fail
          // (bz_move.cpp:348-354 re-tests the OUTPUT row, and the origin is inside every wedge)
          if (mode & MODE_NO_TAU) q[0] = q[1] = q[2] = 0.0;
        }
      }
    }
    double x[3];
    matvec(x, bz.to_xyz, q);  // array_lvec_methods.tpp:40-52
    out.q_ir[3 * i] = q[0]; out.q_ir[3 * i + 1] = q[1]; out.q_ir[3 * i + 2] = q[2];
    if (out.x_ir) { out.x_ir[3 * i] = x[0]; out.x_ir[3 * i + 1] = x[1]; out.x_ir[3 * i + 2] = x[2]; }
    if (out.tau) { out.tau[3 * i] = tau[0]; out.tau[3 * i + 1] = tau[1]; out.tau[3 * i + 2] = tau[2]; }
    out.ridx[i] = ridx;
    out.invridx[i] = invridx;
    uint32_t cell = 0xffffffffu;
    int tet = -1, n_emit = 0;
    if (SPLIT) {
      // two-kernel location: find the node (trellis) / spatial bin (nest, mesh), park the point, count it in that bucket;
      // k_locate_in_node finishes it
      if (KIND == B200_GRID_TRELLIS) {
        st |= trellis_find_node(bz, tr, knots, x, cell);
      } else {
        const BinDev& bn = gd.bins;
        int ib[3], nb[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double f = (x[d] - bn.lo[d]) * bn.inv[d];
          ib[d] = (f > 0.0 ? (f < (double)bn.n[d] ? (int)f : bn.n[d] - 1) : 0) >> out.bin_shift;  // (NaN -> 0)
          nb[d] = ((bn.n[d] - 1) >> out.bin_shift) + 1;
        }
        cell = (uint32_t)(ib[0] + nb[0] * (ib[1] + nb[1] * ib[2]));
      }
      ParkedPoint pp;
      pp.x[0] = x[0]; pp.x[1] = x[1]; pp.x[2] = x[2];
      pp.rot_st = (uint32_t)ridx | ((uint32_t)invridx << 8) | (st << 16);
      pp.cell = cell;
      st32(reinterpret_cast<double*>(out.parked + i), pp.x[0], pp.x[1], pp.x[2], __hiloint2double((int)pp.cell, (int)pp.rot_st));
      // third sector of the point's record (device_tables.cuh): q_ir, rotation indices, point index
      st32(out.weight + REC_DOUBLES * i + 8, q[0], q[1], q[2],
           __hiloint2double((int)(uint32_t)i, (int)((uint32_t)ridx | ((uint32_t)invridx << 16))));
      const uint32_t bucket = (st & B200_ST_NOT_FOUND) ? (KIND == B200_GRID_TRELLIS ? tr.n_nodes : bins_at_level(gd.bins, out.bin_shift)) : cell;
      out.key[i] = bucket;
      if (pending_at) *pending_at = pending_rank;
      pending_rank = atomicAdd(out.node_count + bucket, 1u);
      pending_at = out.rank + i;
      continue;
    }
    if (!SPLIT && !(mode & MODE_NO_LOCATE)) {
      EmitOut e;
      e.v = out.vertex + 8 * i;
      e.w = out.weight + REC_DOUBLES * i;
      if (KIND == B200_GRID_TRELLIS) st |= trellis_locate(bz, tr, knots, x, e, cell, tet);
      else if (KIND == B200_GRID_NEST) st |= nest_locate(bz, gd, x, e, cell, tet);
      else st |= mesh_locate(bz, gd, x, e, cell, tet);
      if (e.n == 0) {  // not found: defined contents for the row
        for (int j = 0; j < 8; ++j) {
          e.v[j] = 0xffffffffu;
          e.w[j] = 0.0;
        }
      }
      n_emit = e.n;
      out.cell[i] = cell;
      out.tet[i] = tet;
      out.n_vert[i] = e.n;
      out.slots[i] = e.slots;
    }
    out.status[i] = st;
    // the rest of the point's record (device_tables.cuh): q_ir, rotation indices, point index
    st32(out.weight + REC_DOUBLES * i + 8, q[0], q[1], q[2],
         __hiloint2double((int)(uint32_t)i, (int)((uint32_t)ridx | ((uint32_t)invridx << 16))));
    if (out.key) {
      // bucket for the cell-batched interpolation: only "generic" points (every corner of the cell carries weight, i.e.
      // the pivot is the cell's first emitted corner) share a bucket with their cell
      uint32_t key = gd.cells.n_cubes + gd.cells.n_tets;
      if (!(mode & MODE_NO_LOCATE) && !(st & (B200_ST_OUTSIDE_BZ | B200_ST_OUTSIDE_WEDGE | B200_ST_NOT_FOUND))) {
        if (tet >= 0 && n_emit == 4) key = gd.cells.n_cubes + (uint32_t)tet;
        else if (tet < 0 && n_emit == 8) key = gd.cells.node_index[cell];
      }
      key = key * out.sub + (uint32_t)invridx;
      out.key[i] = key;
      out.rank[i] = atomicAdd(out.cell_count + key, 1u);
    }
    f_bz += (st & B200_ST_OUTSIDE_BZ) != 0 && !(mode & MODE_ISINSIDE);  // (isinside reports, it does not fail)
    f_wedge += (st & B200_ST_OUTSIDE_WEDGE) != 0 && !(mode & MODE_NO_TAU);
    f_find += (st & B200_ST_NOT_FOUND) != 0;
  }
  if (pending_at) *pending_at = pending_rank;
  // fail_count[0..2]: outside first zone / outside wedge / not found (all zero on the normal path)
  if (f_bz) atomicAdd(fail_count + 0, f_bz);
  if (f_wedge) atomicAdd(fail_count + 1, f_wedge);
  if (f_find) atomicAdd(fail_count + 2, f_find);
}

// Second kernel of the split trellis location: the points arrive sorted by node, so the lanes of a warp work in the same
// node -- same branch (cube / triangulated), same tetrahedra, same trip counts, loads that broadcast.  Same arithmetic and
// the same outputs as the tail of k_locate.
template <int KIND>
__global__ void __launch_bounds__(128, KIND == B200_GRID_TRELLIS ? 5 : 4)
k_locate_in_node(const BZDev* __restrict__ bzg, GridDev gd, size_t n, uint32_t mode, LocateOut out, const uint32_t* __restrict__ order,
                 unsigned long long* __restrict__ fail_count) {
  const TrellisDev& tr = gd.tr;
  const BZDev& bz = *bzg;  // only the default tolerance pair is read
  unsigned long long f_bz = 0, f_wedge = 0, f_find = 0;
  // software pipeline over the grid-stride loop: the parked point of the next trip and the sort order of the trip after it are
  // copied asynchronously into the thread's staging slots in shared memory while this trip runs (nothing is held in registers
  // across the trip: the compiler spilled a prefetched register right behind its load, and the spill store waits for the load it
  // was meant to hide)
  __shared__ __align__(16) double s_park[128][4];
  __shared__ uint32_t s_idx[128];
  const uint32_t sp_addr = (uint32_t)__cvta_generic_to_shared(&s_park[threadIdx.x][0]);
  const uint32_t si_addr = (uint32_t)__cvta_generic_to_shared(&s_idx[threadIdx.x]);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t p_first = p;
  uint32_t i_cur = p < n ? order[p] : 0u, i_nxt = p + stride < n ? order[p + stride] : 0u;
  double2 pa = make_double2(0.0, 0.0), pb = pa;
  if (p < n) {
    const double2* src = reinterpret_cast<const double2*>(out.parked + i_cur);
    pa = src[0];
    pb = src[1];
  }
  for (; p < n; p += stride) {
    if (p != p_first) {  // what the previous trip requested
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      i_cur = i_nxt;
      pa = *reinterpret_cast<const double2*>(&s_park[threadIdx.x][0]);
      pb = *reinterpret_cast<const double2*>(&s_park[threadIdx.x][2]);
      i_nxt = s_idx[threadIdx.x];
    }
    if (p + stride < n) {
      const char* src = reinterpret_cast<const char*>(out.parked + i_nxt);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp_addr), "l"(src) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp_addr + 16u), "l"(src + 16) : "memory");
    }
    if (p + 2 * stride < n) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(si_addr), "l"(order + p + 2 * stride) : "memory");
    else s_idx[threadIdx.x] = 0u;
    asm volatile("cp.async.commit_group;" ::: "memory");
    const size_t i = i_cur;
    ParkedPoint pp;
    pp.x[0] = pa.x; pp.x[1] = pa.y; pp.x[2] = pb.x;
    pp.rot_st = (uint32_t)__double2loint(pb.y);
    pp.cell = (uint32_t)__double2hiint(pb.y);
    uint32_t st = pp.rot_st >> 16;
    uint32_t cell = pp.cell;  // trellis: the node; nest / mesh: the spatial bin (replaced by the containing tetrahedron below)
    const int invridx = (int)((pp.rot_st >> 8) & 0xffu);
    int tet = -1;
    __align__(16) uint32_t vtmp[8];  // the vertex list goes to memory only if somebody reads it (probe, general kernel)
    EmitOut e;
    e.v = out.lean ? vtmp : out.vertex + 8 * i;
    e.w = out.weight + REC_DOUBLES * i;
    e.n = 0;
    e.slots = 0;
    if (KIND == B200_GRID_TRELLIS) {
      if (!(st & B200_ST_NOT_FOUND)) st |= trellis_in_node(bz, tr, pp.x, cell, e, tet);
    } else if (KIND == B200_GRID_NEST) {
      st |= nest_locate(bz, gd, pp.x, e, cell, tet);
    } else {
      st |= mesh_locate(bz, gd, pp.x, e, cell, tet);
    }
    if (e.n == 0) {  // not found: defined contents for the row
      for (int j = 0; j < 8; ++j) {
        e.v[j] = 0xffffffffu;
        e.w[j] = 0.0;
      }
    }
    const int n_emit = e.n;
    const uint32_t general = gd.cells.n_cubes + gd.cells.n_tets;
    uint32_t key = general;
    if (!(st & (B200_ST_OUTSIDE_BZ | B200_ST_OUTSIDE_WEDGE | B200_ST_NOT_FOUND))) {
      if (tet >= 0 && n_emit == 4) key = gd.cells.n_cubes + (uint32_t)tet;
      else if (KIND == B200_GRID_TRELLIS && tet < 0 && n_emit == 8) key = gd.cells.node_index[cell];
    }
    const bool in_cell_bucket = key != general;
    key = key * out.sub + (uint32_t)invridx;
    {
      // The points of a warp sit in the same node: many share the sub-bucket.  One atomicAdd per distinct key of the warp
      // (the leader adds the group's population, the members take consecutive ranks) instead of 32 on the same address.
      const unsigned active = __activemask();
      const unsigned same = __match_any_sync(active, key);
      const int leader = __ffs(same) - 1;
      const unsigned lane = threadIdx.x & 31u;
      uint32_t first = 0;
      if ((int)lane == leader) first = atomicAdd(out.cell_count + key, (uint32_t)__popc(same));
      first = __shfl_sync(same, first, leader);
      out.key[p] = key;  // (sorted position, coalesced; the scatter goes through `order`)
      out.rank[p] = first + (uint32_t)__popc(same & ((1u << lane) - 1u));
    }
    if (!(out.lean && in_cell_bucket)) {  // per-point arrays only the probe and the general kernel read
      if (out.lean) {
        uint4* vo = reinterpret_cast<uint4*>(out.vertex + 8 * i);
        vo[0] = make_uint4(vtmp[0], vtmp[1], vtmp[2], vtmp[3]);
        vo[1] = make_uint4(vtmp[4], vtmp[5], vtmp[6], vtmp[7]);
      }
      out.cell[i] = cell;
      out.tet[i] = tet;
      out.n_vert[i] = n_emit;
      out.slots[i] = e.slots;
      out.status[i] = st;
    }
    (void)mode;
    f_bz += (st & B200_ST_OUTSIDE_BZ) != 0;
    f_wedge += (st & B200_ST_OUTSIDE_WEDGE) != 0;
    f_find += (st & B200_ST_NOT_FOUND) != 0;
  }
  if (f_bz) atomicAdd(fail_count + 0, f_bz);
  if (f_wedge) atomicAdd(fail_count + 1, f_wedge);
  if (f_find) atomicAdd(fail_count + 2, f_find);
}

// ---------------------------------------------------------------------------------------------------------------------
// Second kernel of the split TRELLIS location, warp-cooperative form (the default).
//
// The points arrive sorted by node, so the 32 points of a warp sit in one node (a node holds thousands of points; one warp in
// a hundred straddles two).  k_locate_in_node lets every lane walk the node's records on its own: each step is a dependent load
// from L2 (long-scoreboard stalls were 5.5 per issued instruction).  Here the warp copies the records of the node -- the cube's
// 8 corners, or the tetrahedra of a triangulated node in blocks of TET_BLOCK -- into its slice of shared memory once, coalesced,
// and the lanes scan them from there.  Same arithmetic in the same order as trellis_in_node: identical bits.
// The parked point of the warp's next tile is fetched a tile ahead (its sort order two tiles ahead).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TET_BLOCK = 32;                            // tetrahedra staged per round
constexpr int COOP_WARPS = 4;                            // warps per CTA
constexpr int COOP_RUN = 2;                              // consecutive tiles per warp (2 / 4 / 8 / 16 measured: 0.770 / 0.786 / 0.781 / 0.828 ms)
constexpr int COOP_SLICE = TET_BLOCK * TET_PACK;         // doubles per warp (4608 bytes; a cube record needs 28)

__global__ void __launch_bounds__(COOP_WARPS * 32, 4)
k_trellis_in_node_coop(const BZDev* __restrict__ bzg, GridDev gd, size_t n, LocateOut out, const uint32_t* __restrict__ order,
                       unsigned long long* __restrict__ fail_count) {
  __shared__ __align__(16) double s_rec[COOP_WARPS][COOP_SLICE];
  const TrellisDev& tr = gd.tr;
  const double rel = bzg->def_rel, abs_ = bzg->def_abs;
  const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  double* const rec = s_rec[wib];
  unsigned long long f_bz = 0, f_wedge = 0, f_find = 0;
  // A warp works on runs of COOP_RUN consecutive tiles (the tiles are in node order: within a run the node rarely changes, and
  // the node's header and the records staged in shared memory are reused), the runs are dealt round-robin to the warps.
  const size_t n_tiles = (n + 31) / 32, n_warps = (size_t)gridDim.x * COOP_WARPS, warp_id = (size_t)blockIdx.x * COOP_WARPS + wib;
  auto tile_of = [&](size_t s) { return ((s / COOP_RUN) * n_warps + warp_id) * COOP_RUN + (s % COOP_RUN); };
  if (tile_of(0) >= n_tiles) return;
  uint32_t c_node = 0xffffffffu, c_payload = 0, c_t0 = 0, c_t1 = 0;  // header of the node seen last (warp-uniform)
  bool c_cube = false;
  uint32_t rec_node = 0xffffffffu, rec_base = 0;                       // what the warp's slice of shared memory holds
  // Pipeline: the parked point of the warp's NEXT tile and the sort order of the tile after that are fetched while this tile is
  // worked on -- by asynchronous copies into the warp's staging slots in shared memory, not into registers: a loaded register
  // that stays live across the body was spilled by the compiler right behind its load, which waited for the load there
  // (14 % of the samples of the kernel on that one local store; profiles/README.md).
  __shared__ __align__(16) double s_park[COOP_WARPS][32][4];
  __shared__ uint32_t s_idx[COOP_WARPS][32];
  auto order_at = [&](size_t t) { const size_t p = t * 32 + lane; return t < n_tiles && p < n ? order[p] : 0xffffffffu; };
  auto parked_at = [&](uint32_t i, double2& a, double2& b) {
    if (i != 0xffffffffu) {
      const double2* src = reinterpret_cast<const double2*>(out.parked + i);
      a = src[0];
      b = src[1];
    }
  };
  const uint32_t sp_addr = (uint32_t)__cvta_generic_to_shared(&s_park[wib][lane][0]);
  const uint32_t si_addr = (uint32_t)__cvta_generic_to_shared(&s_idx[wib][lane]);
  uint32_t i_cur = order_at(tile_of(0)), i_nxt = order_at(tile_of(1));
  double2 pa = make_double2(0.0, 0.0), pb = pa;
  parked_at(i_cur, pa, pb);
  for (size_t seq = 0;; ++seq) {
    const size_t tile = tile_of(seq);
    if (seq % COOP_RUN == 0 && tile >= n_tiles) break;  // (a run that starts past the end; tiles past the end inside a run are empty)
    {
      if (i_nxt != 0xffffffffu) {
        const char* src = reinterpret_cast<const char*>(out.parked + i_nxt);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp_addr), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp_addr + 16u), "l"(src + 16) : "memory");
      }
      const size_t t2 = tile_of(seq + 2), p2 = t2 * 32 + lane;
      if (t2 < n_tiles && p2 < n) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(si_addr), "l"(order + p2) : "memory");
      else s_idx[wib][lane] = 0xffffffffu;
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const size_t p = tile * 32 + lane;
    const bool valid = i_cur != 0xffffffffu;
    const size_t i = valid ? i_cur : 0;
    const double x[3] = {pa.x, pa.y, pb.x};
    const uint32_t rot_st = (uint32_t)__double2loint(pb.y);
    const uint32_t cell = (uint32_t)__double2hiint(pb.y);
    uint32_t st = rot_st >> 16;
    const int invridx = (int)((rot_st >> 8) & 0xffu);
    int tet = -1, n_emit = 0;
    uint64_t slots = 0;
    double w[8];
    bool generic = false;   // every corner of the cell carries weight: the weights go out with two vector stores
    bool keep[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { w[j] = 0.0; keep[j] = false; }
    bool is_cube = false;
    // ---- node by node (normally one) ----------------------------------------------------------------------------------
    unsigned todo = __ballot_sync(FULL_MASK, valid && !(st & B200_ST_NOT_FOUND));
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const uint32_t node = __shfl_sync(FULL_MASK, cell, leader);
      const bool mine = valid && !(st & B200_ST_NOT_FOUND) && cell == node;
      todo &= ~__ballot_sync(FULL_MASK, mine);
      if (node != c_node) {  // (dependent loads from L2: worth skipping for the tiles of a run)
        c_node = node;
        c_payload = tr.node_index[node];
        c_cube = tr.node_type[node] == B200_NODE_CUBE;
        if (!c_cube) { c_t0 = tr.poly_offsets[c_payload]; c_t1 = tr.poly_offsets[c_payload + 1]; }
      }
      const uint32_t payload = c_payload;
      if (c_cube) {
        // corners (24 doubles) through shared memory; CubeNode::indices_weights (trellis_node.hpp:130-149)
        if (rec_node != node) {
          if (lane < 12) reinterpret_cast<double2*>(rec)[lane] = reinterpret_cast<const double2*>(tr.cube_pack + 24 * (size_t)payload)[lane];
          rec_node = node;
          __syncwarp();
        }
        if (mine) {
          is_cube = true;
          const double vol = (fabs(rec[0] - rec[21]) * fabs(rec[1] - rec[22])) * fabs(rec[2] - rec[23]);
          bool all = true;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            w[c] = ((fabs(x[0] - rec[3 * c]) * fabs(x[1] - rec[3 * c + 1])) * fabs(x[2] - rec[3 * c + 2])) / vol;
            keep[c] = !approx_eq(w[c], 0.0, rel, abs_) && w[c] > 0.0;  // w.is(gt, 0.)
            all &= keep[c];
          }
          generic = all;
          n_emit = 8;
          slots = 0x0001020304050607ull;
        }
        __syncwarp();
      } else {
        // PolyNode::indices_weights (trellis_node.hpp:273-308): first tetrahedron, in storage order, that contains the point
        const uint32_t t0 = c_t0, t1 = c_t1;
        double best = 0.0;
        uint32_t best_at = t0;
        bool have_best = false;
        int found = -1;
        for (uint32_t base = t0; base < t1; base += TET_BLOCK) {
          if (!__ballot_sync(FULL_MASK, mine && found < 0)) break;
          const uint32_t nblk = min((uint32_t)TET_BLOCK, t1 - base);
          if (rec_node != node || rec_base != base) {
            const double2* src = reinterpret_cast<const double2*>(tr.tet_pack + (size_t)TET_PACK * base);
            for (uint32_t c = lane; c < nblk * (TET_PACK / 2); c += 32) reinterpret_cast<double2*>(rec)[c] = src[c];
            rec_node = node;
            rec_base = base;
            __syncwarp();
          }
          if (mine && found < 0) {
            // circumsphere test of every tetrahedron of the block (tetrahedra_might_contain :349-364), then the weights of the
            // candidates in storage order (first accepted wins, :284-290); `best` keeps max_element's first-maximum rule
            unsigned cand = 0u;
            for (uint32_t k = 0; k < nblk; ++k) {
              const double* tp = rec + TET_PACK * k;
              const double v0 = tp[0] - x[0], v1 = tp[1] - x[1], v2 = tp[2] - x[2];
              const double d2 = ((0.0 + v0 * v0) + v1 * v1) + v2 * v2;
              if (d2 < tp[3] || approx_eq(d2, tp[3], rel, abs_)) {
                cand |= 1u << k;
              } else {
                const double mn = -d2;
                if (!have_best || mn > best || (mn == best && base + k < best_at)) {
                  best = mn;
                  best_at = base + k;
                  have_best = true;
                }
              }
            }
            while (cand && found < 0) {
              const uint32_t kk = (uint32_t)(__ffs((int)cand) - 1);
              cand &= cand - 1u;
              const double mn = tet_weights(rec + TET_PACK * kk, x, w, rel, abs_);
              if (mn >= 0.0) {
                found = (int)(base + kk);
              } else if (!have_best || mn > best || (mn == best && base + kk < best_at)) {
                best = mn;
                best_at = base + kk;
                have_best = true;
              }
            }
          }
          __syncwarp();
        }
        if (mine) {
          if (found < 0) {
            if (t1 == t0) {
              st |= B200_ST_NOT_FOUND;
            } else {
              st |= B200_ST_FALLBACK_TET;  // trellis_node.hpp:295-306
              found = (int)best_at;
              tet_weights(tr.tet_pack + (size_t)TET_PACK * best_at, x, w, rel, abs_);
            }
          }
          if (found >= 0) {
            tet = found;
            bool all = true;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              keep[j] = !approx_eq(w[j], 0.0, rel, abs_);
              all &= keep[j];
            }
            generic = all;
            n_emit = 4;
            slots = 0x03020100ull;
#pragma unroll
            for (int j = 4; j < 8; ++j) w[j] = 0.0;
          }
        }
      }
    }
    // ---- outputs (as k_locate_in_node) -----------------------------------------------------------------------------------
    if (valid) {
      double* const wrow = out.weight + REC_DOUBLES * i;
      __align__(16) uint32_t vtmp[8];
      bool wrote_vertices = false;
      if (generic) {
        st32(wrow, w[0], w[1], w[2], w[3]);
        st32(wrow + 4, w[4], w[5], w[6], w[7]);
      } else {
        // a point on a face / edge / vertex of its cell, or one that was not found: compact emission (rare)
        EmitOut e;
        e.v = vtmp;
        e.w = wrow;
        e.n = 0;
        e.slots = 0;
        if (n_emit) {
          uint32_t vi[8];
          if (is_cube) {
            const uint4* vip = reinterpret_cast<const uint4*>(tr.cube_vertices + 8 * (size_t)tr.node_index[cell]);
            const uint4 va = vip[0], vb = vip[1];
            vi[0] = va.x; vi[1] = va.y; vi[2] = va.z; vi[3] = va.w; vi[4] = vb.x; vi[5] = vb.y; vi[6] = vb.z; vi[7] = vb.w;
          } else {
            const uint4 v4 = *reinterpret_cast<const uint4*>(tr.tet_vertices + 4 * (size_t)tet);
            vi[0] = v4.x; vi[1] = v4.y; vi[2] = v4.z; vi[3] = v4.w;
          }
          double wc[8];  // (copies: the arrays handed to the out-of-line routine live in local memory, w and keep must not)
          bool kc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { wc[j] = w[j]; kc[j] = keep[j]; }
          emit_compact(e, vi, wc, kc, is_cube ? 8 : 4, is_cube);
        }
        if (e.n == 0) {
          for (int j = 0; j < 8; ++j) { vtmp[j] = 0xffffffffu; wrow[j] = 0.0; }
          if (n_emit) st |= B200_ST_NOT_FOUND;
        }
        n_emit = e.n;
        slots = e.slots;
        wrote_vertices = true;
      }
      const uint32_t general = gd.cells.n_cubes + gd.cells.n_tets;
      uint32_t key = general;
      if (generic && !(st & (B200_ST_OUTSIDE_BZ | B200_ST_OUTSIDE_WEDGE | B200_ST_NOT_FOUND)))
        key = tet >= 0 ? gd.cells.n_cubes + (uint32_t)tet : gd.cells.node_index[cell];
      const bool in_cell_bucket = key != general;
      key = key * out.sub + (uint32_t)invridx;
      {
        const unsigned active = __activemask();
        const unsigned same = __match_any_sync(active, key);
        const int lead = __ffs(same) - 1;
        uint32_t first = 0;
        if ((int)lane == lead) first = atomicAdd(out.cell_count + key, (uint32_t)__popc(same));
        first = __shfl_sync(same, first, lead);
        out.key[p] = key;
        out.rank[p] = first + (uint32_t)__popc(same & ((1u << lane) - 1u));
      }
      if (!(out.lean && in_cell_bucket)) {  // per-point arrays only the probe and the general kernel read
        if (!wrote_vertices) {
          if (is_cube) {
            const uint4* vip = reinterpret_cast<const uint4*>(tr.cube_vertices + 8 * (size_t)tr.node_index[cell]);
            const uint4 va = vip[0], vb = vip[1];
            vtmp[0] = vb.w; vtmp[1] = vb.z; vtmp[2] = vb.y; vtmp[3] = vb.x; vtmp[4] = va.w; vtmp[5] = va.z; vtmp[6] = va.y; vtmp[7] = va.x;
          } else {
            const uint4 v4 = *reinterpret_cast<const uint4*>(tr.tet_vertices + 4 * (size_t)tet);
            vtmp[0] = v4.x; vtmp[1] = v4.y; vtmp[2] = v4.z; vtmp[3] = v4.w;
            vtmp[4] = vtmp[5] = vtmp[6] = vtmp[7] = 0xffffffffu;
          }
        }
        uint4* vo = reinterpret_cast<uint4*>(out.vertex + 8 * i);
        vo[0] = make_uint4(vtmp[0], vtmp[1], vtmp[2], vtmp[3]);
        vo[1] = make_uint4(vtmp[4], vtmp[5], vtmp[6], vtmp[7]);
        out.cell[i] = cell;
        out.tet[i] = tet;
        out.n_vert[i] = n_emit;
        out.slots[i] = slots;
        out.status[i] = st;
      }
      f_bz += (st & B200_ST_OUTSIDE_BZ) != 0;
      f_wedge += (st & B200_ST_OUTSIDE_WEDGE) != 0;
      f_find += (st & B200_ST_NOT_FOUND) != 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    i_cur = i_nxt;
    if (i_cur != 0xffffffffu) {
      pa = *reinterpret_cast<const double2*>(&s_park[wib][lane][0]);
      pb = *reinterpret_cast<const double2*>(&s_park[wib][lane][2]);
    }
    i_nxt = s_idx[wib][lane];
  }
  if (f_bz) atomicAdd(fail_count + 0, f_bz);
  if (f_wedge) atomicAdd(fail_count + 1, f_wedge);
  if (f_find) atomicAdd(fail_count + 2, f_find);
}

cudaError_t launch_locate_in_node(const BZDev* bzg, const GridDev& gd, size_t n, uint32_t mode, const LocateOut& out, const uint32_t* order,
                                  unsigned long long* fail_count, int sm_count, cudaStream_t stream, bool coop) {
  if (n == 0) return cudaSuccess;
  if (coop && gd.kind == B200_GRID_TRELLIS) {
    const size_t tiles = (n + 31) / 32, per_cta = (size_t)COOP_WARPS * COOP_RUN, want_c = (tiles + per_cta - 1) / per_cta, cap_c = (size_t)sm_count * 16;
    k_trellis_in_node_coop<<<(unsigned)(want_c < cap_c ? want_c : cap_c), COOP_WARPS * 32, 0, stream>>>(bzg, gd, n, out, order, fail_count);
    return cudaGetLastError();
  }
  const size_t want = (n + 127) / 128, cap = (size_t)sm_count * 32;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  switch (gd.kind) {
    case B200_GRID_TRELLIS: k_locate_in_node<B200_GRID_TRELLIS><<<grid, 128, 0, stream>>>(bzg, gd, n, mode, out, order, fail_count); break;
    case B200_GRID_NEST: k_locate_in_node<B200_GRID_NEST><<<grid, 128, 0, stream>>>(bzg, gd, n, mode, out, order, fail_count); break;
    default: k_locate_in_node<B200_GRID_MESH><<<grid, 128, 0, stream>>>(bzg, gd, n, mode, out, order, fail_count); break;
  }
  return cudaGetLastError();
}

static size_t locate_smem_bytes(const GridDev& gd, uint32_t mode) {
  const TrellisDev& tr = gd.tr;
  size_t nk = (gd.kind != B200_GRID_TRELLIS || (mode & MODE_NO_LOCATE)) ? 0 : (size_t)(tr.n_knots[0] + tr.n_knots[1] + tr.n_knots[2]);
  return ((sizeof(BZDev) + 15) / 16) * 16 + nk * sizeof(double);
}

template <int KIND, bool SPLIT>
static cudaError_t launch_kind(const BZDev* bzg, const GridDev& gd, const double* Q, size_t n, uint32_t mode, double eps_w,
                               double eps_o, const LocateOut& out, unsigned long long* fail_count, int sm_count,
                               cudaStream_t stream) {
  const int threads = 128;
  const size_t smem = locate_smem_bytes(gd, mode);
  static bool attr_set_dev[MAX_DEVICES] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    cudaFuncSetAttribute(k_locate<KIND, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_set = true;
  }
  size_t want = (n + threads - 1) / threads;
  size_t cap = (size_t)sm_count * 16;  // grid-stride: a multiple of the SM count
  int blocks = (int)(want < cap ? want : cap);
  k_locate<KIND, SPLIT><<<blocks, threads, smem, stream>>>(bzg, gd, Q, n, mode, eps_w, eps_o, out, fail_count);
  return cudaGetLastError();
}

cudaError_t launch_locate(const BZDev* bzg, const GridDev& gd, const double* Q, size_t n, uint32_t mode, double eps_w,
                          double eps_o, const LocateOut& out, unsigned long long* fail_count, int sm_count,
                          cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  switch (gd.kind) {
    case B200_GRID_TRELLIS:
      if (mode & MODE_SPLIT_A) return launch_kind<B200_GRID_TRELLIS, true>(bzg, gd, Q, n, mode, eps_w, eps_o, out, fail_count, sm_count, stream);
      return launch_kind<B200_GRID_TRELLIS, false>(bzg, gd, Q, n, mode, eps_w, eps_o, out, fail_count, sm_count, stream);
    case B200_GRID_NEST:
      if (mode & MODE_SPLIT_A) return launch_kind<B200_GRID_NEST, true>(bzg, gd, Q, n, mode, eps_w, eps_o, out, fail_count, sm_count, stream);
      return launch_kind<B200_GRID_NEST, false>(bzg, gd, Q, n, mode, eps_w, eps_o, out, fail_count, sm_count, stream);
    default:
      if (mode & MODE_SPLIT_A) return launch_kind<B200_GRID_MESH, true>(bzg, gd, Q, n, mode, eps_w, eps_o, out, fail_count, sm_count, stream);
      return launch_kind<B200_GRID_MESH, false>(bzg, gd, Q, n, mode, eps_w, eps_o, out, fail_count, sm_count, stream);
  }
}

}  // namespace b200
