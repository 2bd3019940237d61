// brille_b200 bridge: flatten brille's host C++ objects into plain SoA tables.
// (header: included by the _bridge module, flatten.cpp, and by the pybind11 add-on module, accel/accel.cpp)
//
// This code is the *reference-side* half of the drop-in boundary:
// it is compiled against brille's own headers and walks the already-constructed host objects
// (BrillouinZone, BrillouinZoneTrellis3/Nest3/Mesh3, DualInterpolator, GammaTable) ONCE, emitting
// plain contiguous arrays -- the `b200_tables_t` of include/brille_b200.h -- that are then uploaded
// to the GPU through the C ABI.  Construction (lattice, symmetry, polyhedra, TetGen, fill, sort)
// stays brille's host C++; nothing here runs per Q point.
//
// Quantities that brille recomputes at the start of every `moveinto` call (plane points in the
// primitive lattice, normalised face normals, tau vectors: bz_move.cpp:118-140) are produced here by
// calling the very same brille functions, so the tables hold bit-identical numbers.
//
// Non-public members are reached without touching the reference sources:
//   * protected members through a derived "spy" class and a pointer-to-member
//   * private members through the explicit-instantiation access idiom (Access<Tag, &T::member>)
#pragma once
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <pybind11/numpy.h>
#include <pybind11/complex.h>
#include <complex>
#include <cstring>
#include <deque>
#include "bz_trellis.hpp"
#include "bz_nest.hpp"
#include "bz_mesh.hpp"

namespace py = pybind11;
using namespace brille;
using cplx = std::complex<double>;

// ---------------------------------------------------------------------------------------------
// access helpers
// ---------------------------------------------------------------------------------------------
struct BZSpy : BrillouinZone {
  static double ftol(const BrillouinZone& b) { return b.*(&BZSpy::float_tolerance); }
  static int atol(const BrillouinZone& b) { return b.*(&BZSpy::approx_tolerance); }
  static bool nomirror(const BrillouinZone& b) { return b.*(&BZSpy::no_ir_mirroring); }
  static bool isprim(const BrillouinZone& b) { return b.*(&BZSpy::is_primitive); }
  static const poly_t& first(const BrillouinZone& b) { return b.*(&BZSpy::_first); }
};
struct PolyNodeSpy : brille::PolyNode {
  static const std::vector<std::array<ind_t, 4>>& vi(const brille::PolyNode& n) { return n.*(&PolyNodeSpy::vi_t); }
  static const std::vector<std::array<double, 4>>& ci(const brille::PolyNode& n) { return n.*(&PolyNodeSpy::ci_t); }
  static const std::vector<double>& vol(const brille::PolyNode& n) { return n.*(&PolyNodeSpy::vol_t); }
};
struct CubeNodeSpy : brille::CubeNode {
  static const std::array<ind_t, 8>& vi(const brille::CubeNode& n) { return n.*(&CubeNodeSpy::vertex_indices); }
};
template <class T, class R, class S>
struct TrellisSpy : BrillouinZoneTrellis3<T, R, S> {
  using base = BrillouinZoneTrellis3<T, R, S>;
  static const typename base::knots_t& knots(const base& g) { return g.*(&TrellisSpy::knots_); }
  static const typename base::nodes_t& nodes(const base& g) { return g.*(&TrellisSpy::nodes_); }
};
template <class T>
struct InterpSpy : Interpolator<T> {
  static LengthUnit lenunit(const Interpolator<T>& i) { return i.*(&InterpSpy::lenunit_); }
  static const std::array<double, 3>& costmult(const Interpolator<T>& i) { return i.*(&InterpSpy::_costmult); }
  static const std::array<ind_t, 3>& funtype(const Interpolator<T>& i) { return i.*(&InterpSpy::_funtype); }
};
template <class T, class R>
struct DualSpy : DualInterpolator<T, R> {
  static const PermutationTable& table(const DualInterpolator<T, R>& d) { return d.*(&DualSpy::permutation_table_); }
  static PermutationTable& table(DualInterpolator<T, R>& d) { return d.*(&DualSpy::permutation_table_); }
};
struct PermSpy : PermutationTable {
  static const std::map<size_t, size_t>& map(const PermutationTable& p) { return p.*(&PermSpy::ijmap); }
  static const std::vector<std::vector<ind_t>>& perms(const PermutationTable& p) { return p.*(&PermSpy::permutations); }
  static std::vector<std::vector<ind_t>>& perms(PermutationTable& p) { return p.*(&PermSpy::permutations); }
  static size_t nidx(const PermutationTable& p) { return p.*(&PermSpy::IndexSize); }
};

// private-member access (legal: explicit instantiation may name private members)
template <class Tag, typename Tag::type M>
struct Access {
  friend typename Tag::type get(Tag) { return M; }
};
struct LeafCR { using type = std::array<double, 4> NestLeaf::*; friend type get(LeafCR); };
template struct Access<LeafCR, &NestLeaf::centre_radius>;
template <class T, class R>
struct NestRoot { using type = NestNode Nest<T, R, double, Array2>::*; friend type get(NestRoot); };
template struct Access<NestRoot<double, double>, &Nest<double, double, double, Array2>::root_>;
template struct Access<NestRoot<double, cplx>, &Nest<double, cplx, double, Array2>::root_>;
template struct Access<NestRoot<cplx, cplx>, &Nest<cplx, cplx, double, Array2>::root_>;
struct TetTriLayers { using type = std::vector<TetTriLayer> TetTri::*; friend type get(TetTriLayers); };
template struct Access<TetTriLayers, &TetTri::layers>;
struct TetTriConns { using type = std::vector<std::vector<std::vector<ind_t>>> TetTri::*; friend type get(TetTriConns); };
template struct Access<TetTriConns, &TetTri::connections>;
template <class T, class R>
struct MeshSpy : Mesh3<T, R, double, Array2> {
  static const TetTri& mesh_of(const Mesh3<T, R, double, Array2>& m) { return m.*(&MeshSpy::mesh); }
};

// ---------------------------------------------------------------------------------------------
// numpy helpers (always own a copy: tables must outlive the host objects)
// ---------------------------------------------------------------------------------------------
template <class T>
static py::array_t<T> np1(const std::vector<T>& v) {
  py::array_t<T> a(static_cast<py::ssize_t>(v.size()));
  if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(T));
  return a;
}
template <class T>
static py::array_t<T> np2(const std::vector<T>& v, size_t cols) {
  py::array_t<T> a({static_cast<py::ssize_t>(cols ? v.size() / cols : 0), static_cast<py::ssize_t>(cols)});
  if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(T));
  return a;
}
template <class T, class A>
static py::array_t<T> np_from_a2(const A& arr) {  // (n, m) Array2 / LVec -> contiguous numpy
  std::vector<T> v;
  v.reserve(static_cast<size_t>(arr.size(0)) * arr.size(1));
  for (ind_t i = 0; i < arr.size(0); ++i)
    for (ind_t j = 0; j < arr.size(1); ++j) v.push_back(static_cast<T>(arr.val(i, j)));
  return np2(v, arr.size(1));
}
template <class T, size_t N>
static py::array_t<T> np_arr(const std::array<T, N>& a) {
  return np1(std::vector<T>(a.begin(), a.end()));
}

// ---------------------------------------------------------------------------------------------
// BrillouinZone  ->  tables used by ir_moveinto   (bz_move.cpp:103-296)
// ---------------------------------------------------------------------------------------------
static py::dict flatten_bz(const BrillouinZone& bz) {
  using namespace brille::lattice;
  py::dict d;
  const auto outer = bz.get_lattice();
  const auto inner = bz.get_primitive_lattice();
  PrimitiveTransform PT(outer.bravais());
  // bz_move.cpp:111 (Q is always given in the outer lattice on this path)
  const bool transform_needed = PT.does_anything();
  if (transform_needed != BZSpy::isprim(bz))
    throw std::runtime_error(
        "brille_b200: BrillouinZone built with primitive=false on a centred lattice is not supported "
        "(the reference's own moveinto throws for it)");
  d["transform_needed"] = transform_needed ? 1 : 0;
  d["P6t"] = np_arr<int>(PT.get_6Pt());      // transform.hpp:175 (inverse_angstrom: 6*P^T, then /6)
  d["invPt"] = np_arr<int>(PT.get_invPt());  // transform.hpp:213
  const auto work = transform_needed ? outer.primitive() : outer;
  d["w_recip_metric"] = np_arr<double>(work.metric(LengthUnit::inverse_angstrom));
  d["w_real_metric"] = np_arr<double>(work.metric(LengthUnit::angstrom));
  d["w_recip_volume"] = work.volume(LengthUnit::inverse_angstrom);
  d["o_recip_metric"] = np_arr<double>(outer.metric(LengthUnit::inverse_angstrom));
  d["o_real_metric"] = np_arr<double>(outer.metric(LengthUnit::angstrom));
  d["o_recip_volume"] = outer.volume(LengthUnit::inverse_angstrom);
  d["to_xyz"] = np_arr<double>(outer.to_xyz(LengthUnit::inverse_angstrom));  // array_lvec_methods.tpp:40-52

  // first-Brillouin-zone planes: conventional (isinside re-check, bz.hpp:631-642) and working lattice
  const auto& first = BZSpy::first(bz);
  auto [a, b, c] = first.planes();
  d["ca"] = np_from_a2<double>(a);
  d["cb"] = np_from_a2<double>(b);
  d["cc"] = np_from_a2<double>(c);
  auto pa = parallel_transform_to_primitive(outer, a, 1);  // bz_move.cpp:124-126
  auto pb = parallel_transform_to_primitive(outer, b, 1);
  auto pc = parallel_transform_to_primitive(outer, c, 1);
  d["pa"] = np_from_a2<double>(pa);
  d["pb"] = np_from_a2<double>(pb);
  d["pc"] = np_from_a2<double>(pc);
  auto normals = bz.get_primitive_normals();  // bz_move.cpp:137-140
  normals = normals / norm(normals);
  auto taus = (2.0 * bz.get_primitive_points()).round();
  auto tau_lens = norm(taus);
  d["normals"] = np_from_a2<double>(normals);
  d["taus"] = np_from_a2<int>(taus);
  d["tau_lens"] = np_from_a2<double>(tau_lens).attr("reshape")(-1);

  // irreducible wedge (bz.hpp:757-763)
  auto wn = bz.get_ir_wedge_normals();
  d["wedge_normals"] = np_from_a2<double>(wn).attr("reshape")(-1, 3);
  d["no_ir_mirroring"] = BZSpy::nomirror(bz) ? 1 : 0;
  d["float_tolerance"] = BZSpy::ftol(bz);
  d["approx_tolerance"] = BZSpy::atol(bz);
  d["time_reversal"] = bz.add_time_reversal();

  // point group in the order ir_moveinto scans it (bz_move.cpp:257-285)
  PointSymmetry ps = bz.get_pointgroup_symmetry();
  std::vector<int> rot;
  std::vector<int> inv;
  for (size_t i = 0; i < ps.size(); ++i) {
    auto r = ps.get(i);
    rot.insert(rot.end(), r.begin(), r.end());
    inv.push_back(static_cast<int>(ps.get_inverse_index(i)));  // pointsymmetry.cpp:131-142
  }
  d["rotations"] = np2(rot, 9);
  d["inverse_index"] = np1(inv);
  d["identity_index"] = static_cast<int>(ps.find_identity_index());
  return d;
}

// ---------------------------------------------------------------------------------------------
// PermutationTable lookup identical to safe_get (permutation_table.hpp:193-198,214)
// ---------------------------------------------------------------------------------------------
struct PermLookup {
  const std::map<size_t, size_t>& m;
  size_t n;
  explicit PermLookup(const PermutationTable& t) : m(PermSpy::map(t)), n(PermSpy::nidx(t)) {}
  unsigned row(size_t i, size_t j) const {
    size_t key = (i == j) ? 0u : i * n + j;
    auto it = m.find(key);
    return (it != m.end() && it->second >= 1u) ? static_cast<unsigned>(it->second - 1u) : 0u;
  }
};

template <class T, class R>
static void flatten_perm_rows(const DualInterpolator<T, R>& data, py::dict& d, bool& any_nonidentity) {
  const auto& rows = PermSpy::perms(DualSpy<T, R>::table(data));
  std::vector<unsigned> flat;
  size_t m = rows.empty() ? 0 : rows[0].size();
  for (const auto& r : rows) flat.insert(flat.end(), r.begin(), r.end());
  d["perm_rows"] = np2(flat, m);
  any_nonidentity = rows.size() > 1;
}

// ---------------------------------------------------------------------------------------------
// interpolation data   (interpolatordual.hpp, interpolator.hpp, phonon.hpp)
// ---------------------------------------------------------------------------------------------
template <class T>
static void flatten_interp(const Interpolator<T>& in, const char* prefix, py::dict& d) {
  std::string p(prefix);
  const auto& a = in.data();  // Array2 (n_pt, branches*span)
  py::array_t<T> arr({static_cast<py::ssize_t>(a.size(0)), static_cast<py::ssize_t>(a.size(1))});
  T* dst = arr.mutable_data();
  for (ind_t i = 0; i < a.size(0); ++i)
    for (ind_t j = 0; j < a.size(1); ++j) dst[static_cast<size_t>(i) * a.size(1) + j] = a.val(i, j);
  d[(p + "_data").c_str()] = arr;
  auto sh = in.shape();
  d[(p + "_shape").c_str()] = std::vector<unsigned>(sh.begin(), sh.end());
  auto el = in.elements();
  d[(p + "_elements").c_str()] = np1(std::vector<unsigned>(el.begin(), el.end()));
  d[(p + "_rotlike").c_str()] = static_cast<int>(in.rotateslike());
  d[(p + "_lenunit").c_str()] = static_cast<int>(InterpSpy<T>::lenunit(in));
  d[(p + "_branches").c_str()] = in.branches();
  d[(p + "_span").c_str()] = in.branch_span();
}

template <class Grid>
static void flatten_gamma(const Grid& g, py::dict& d) {
  // identical arguments to bz_trellis.hpp:182-186
  auto bz = g.get_brillouinzone();
  auto cfg = g.approx_config();
  auto lat = bz.get_lattice();
  bool has_basis = lat.basis().size() > 0;
  PointSymmetry ps = bz.get_pointgroup_symmetry();
  const size_t nops = ps.size();
  // Cartesian rotation matrices used when LengthUnit::angstrom (interpolator.hpp:409-423)
  std::vector<double> rc(nops * 9);
  std::array<double, 9> t0;
  for (size_t j = 0; j < nops; ++j) {
    brille::utils::mul_mat_mat(t0.data(), 3u, lat.to_xyz(LengthUnit::angstrom).data(), ps.data(j));
    brille::utils::mul_mat_mat(&rc[9 * j], 3u, t0.data(), lat.from_xyz(LengthUnit::angstrom).data());
  }
  d["rot_cart"] = np2(rc, 9);
  if (!has_basis) {
    d["gamma_natoms"] = 0;
    return;
  }
  GammaTable gt(true, lat, bz.add_time_reversal(), cfg.template direct<double>(), cfg.digit());
  const size_t nat = lat.basis().size();
  std::vector<unsigned> f0(nat * nops), vi(nat * nops);
  for (size_t k = 0; k < nat; ++k)
    for (size_t r = 0; r < nops; ++r) {
      f0[k * nops + r] = gt.F0(k, r);
      vi[k * nops + r] = gt.vector_index(k, r);
    }
  d["gamma_natoms"] = nat;
  d["gamma_F0"] = np2(f0, nops);
  d["gamma_vidx"] = np2(vi, nops);
  d["gamma_vectors"] = np_from_a2<double>(gt.vectors());
}

template <class Grid>
static py::dict flatten_data_common(const Grid& g, py::dict d) {
  const auto& data = g.data();
  flatten_interp(data.values(), "values", d);
  flatten_interp(data.vectors(), "vectors", d);
  bool gamma_needed = RotatesLike::Gamma == data.vectors().rotateslike() ||
                      RotatesLike::Gamma == data.values().rotateslike();
  d["gamma_needed"] = gamma_needed ? 1 : 0;
  try {
    flatten_gamma(g, d);
  } catch (const std::exception& e) {
    if (gamma_needed) throw;
    d["gamma_natoms"] = 0;
  }
  return d;
}

// ---------------------------------------------------------------------------------------------
// PolyTrellis   (trellis_poly.hpp, trellis_node.hpp)
// ---------------------------------------------------------------------------------------------
template <class T, class R, class S>
static py::dict flatten_trellis(const BrillouinZoneTrellis3<T, R, S>& g) {
  using Spy = TrellisSpy<T, R, S>;
  py::dict d;
  d["kind"] = "trellis";
  d["bz"] = flatten_bz(g.get_brillouinzone());
  const auto& knots = Spy::knots(g);
  d["knots0"] = np1(knots[0]);
  d["knots1"] = np1(knots[1]);
  d["knots2"] = np1(knots[2]);
  const auto& nodes = Spy::nodes(g);
  const size_t nn = nodes.size();
  std::vector<uint8_t> ntype(nn);
  std::vector<unsigned> nidx(nn, 0xffffffffu);
  std::vector<unsigned> cubes, poff{0u}, tvi;
  std::vector<double> tci, tvol;
  unsigned ncube = 0, npoly = 0;
  for (size_t i = 0; i < nn; ++i) {
    auto t = nodes.type(static_cast<ind_t>(i));
    ntype[i] = static_cast<uint8_t>(t);  // enums.hpp:43
    if (NodeType::cube == t) {
      nidx[i] = ncube++;
      for (auto v : CubeNodeSpy::vi(nodes.cube_at(static_cast<ind_t>(i)))) cubes.push_back(v);
    } else if (NodeType::poly == t) {
      nidx[i] = npoly++;
      const auto& pn = nodes.poly_at(static_cast<ind_t>(i));
      const auto& vi = PolyNodeSpy::vi(pn);
      const auto& ci = PolyNodeSpy::ci(pn);
      const auto& vol = PolyNodeSpy::vol(pn);
      for (size_t k = 0; k < vi.size(); ++k) {
        for (int j = 0; j < 4; ++j) tvi.push_back(vi[k][j]);
        for (int j = 0; j < 4; ++j) tci.push_back(ci[k][j]);
        tvol.push_back(vol[k]);
      }
      poff.push_back(static_cast<unsigned>(tvol.size()));
    }
  }
  d["node_type"] = np1(ntype);
  d["node_index"] = np1(nidx);
  d["cube_vertices"] = np2(cubes, 8);
  d["poly_offsets"] = np1(poff);
  d["tet_vertices"] = np2(tvi, 4);
  d["tet_circum"] = np2(tci, 4);
  d["tet_volume"] = np1(tvol);
  d["vertices"] = np_from_a2<double>(g.vertices());
  auto cfg = g.approx_config();
  d["approx_digit"] = cfg.digit();
  d["approx_direct"] = cfg.template direct<double>();
  d["approx_reciprocal"] = cfg.template reciprocal<double>();
  return d;
}

template <class T, class R, class S>
static py::dict flatten_trellis_data(const BrillouinZoneTrellis3<T, R, S>& g) {
  py::dict d;
  flatten_data_common(g, d);
  bool any = false;
  flatten_perm_rows(g.data(), d, any);
  d["perm_nonidentity"] = any ? 1 : 0;
  {  // the vertex lists the per-cell pair tables are indexed by (also the input of the device-side sort())
    using Spy = TrellisSpy<T, R, S>;
    const auto& nodes = Spy::nodes(g);
    std::vector<unsigned> cv, tv;
    for (size_t i = 0; i < nodes.size(); ++i) {
      auto t = nodes.type(static_cast<ind_t>(i));
      if (NodeType::cube == t) {
        for (auto v : CubeNodeSpy::vi(nodes.cube_at(static_cast<ind_t>(i)))) cv.push_back(v);
      } else if (NodeType::poly == t) {
        for (const auto& vi : PolyNodeSpy::vi(nodes.poly_at(static_cast<ind_t>(i))))
          for (int a = 0; a < 4; ++a) tv.push_back(vi[a]);
      }
    }
    d["perm_cube_vertices"] = np2(cv, 8);
    d["perm_tet_vertices"] = np2(tv, 4);
  }
  if (any) {
    // per-cell pair -> permutation-row index, so the device needs no map lookups
    using Spy = TrellisSpy<T, R, S>;
    PermLookup look(DualSpy<T, R>::table(g.data()));
    const auto& nodes = Spy::nodes(g);
    std::vector<unsigned> cp, tp;
    for (size_t i = 0; i < nodes.size(); ++i) {
      auto t = nodes.type(static_cast<ind_t>(i));
      if (NodeType::cube == t) {
        const auto& vi = CubeNodeSpy::vi(nodes.cube_at(static_cast<ind_t>(i)));
        for (int a = 0; a < 8; ++a)
          for (int b = 0; b < 8; ++b) cp.push_back(look.row(vi[a], vi[b]));
      } else if (NodeType::poly == t) {
        for (const auto& vi : PolyNodeSpy::vi(nodes.poly_at(static_cast<ind_t>(i))))
          for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) tp.push_back(look.row(vi[a], vi[b]));
      }
    }
    d["cube_perm"] = np2(cp, 64);
    d["tet_perm"] = np2(tp, 16);
  }
  return d;
}

// ---------------------------------------------------------------------------------------------
// Nest (nest.hpp): the tree flattened breadth-first, children of a node contiguous and in storage order,
// i.e. exactly the order in which NestNode::indices_weights' deque visits them (nest.hpp:163-192)
// ---------------------------------------------------------------------------------------------
template <class T, class R>
static py::dict flatten_nest(const BrillouinZoneNest3<T, R, double>& g) {
  py::dict d;
  d["kind"] = "nest";
  d["bz"] = flatten_bz(g.get_brillouinzone());
  const Nest<T, R, double, Array2>& nest = g;
  const NestNode& root = nest.*get(NestRoot<T, R>());
  std::vector<unsigned> vi, cbeg, cend;
  std::vector<double> cr, vol;
  std::vector<uint8_t> leaf;
  std::deque<const NestNode*> work;
  auto push = [&](const NestNode& n, bool is_root) {
    const auto& lf = n.boundary();
    for (auto v : lf.vertices()) vi.push_back(v);
    const auto& c = lf.*get(LeafCR());
    for (auto x : c) cr.push_back(x);
    vol.push_back(lf.volume());
    leaf.push_back((!is_root && n.is_leaf()) ? 1 : 0);
    cbeg.push_back(0);
    cend.push_back(0);
  };
  push(root, true);  // node 0 = root (its own boundary is unused)
  work.push_back(&root);
  size_t at = 0;
  while (!work.empty()) {
    const NestNode* n = work.front();
    work.pop_front();
    cbeg[at] = static_cast<unsigned>(vol.size());
    for (const auto& b : n->branches()) {
      push(b, false);
      work.push_back(&b);
    }
    cend[at] = static_cast<unsigned>(vol.size());
    ++at;
  }
  d["node_vertices"] = np2(vi, 4);
  d["node_circum"] = np2(cr, 4);
  d["node_volume"] = np1(vol);
  d["node_is_leaf"] = np1(leaf);
  d["child_begin"] = np1(cbeg);
  d["child_end"] = np1(cend);
  d["vertices"] = np_from_a2<double>(g.all_vertices());
  auto cfg = g.approx_config();
  d["approx_digit"] = cfg.digit();
  d["approx_direct"] = cfg.template direct<double>();
  d["approx_reciprocal"] = cfg.template reciprocal<double>();
  return d;
}

// per-tetrahedron pair -> permutation row for tetrahedral grids (nest: per node, mesh: per finest-layer tetrahedron)
template <class T, class R>
static void flatten_tet_perms(const DualInterpolator<T, R>& data, const std::vector<unsigned>& tets, py::dict& d) {
  bool any = false;
  flatten_perm_rows(data, d, any);
  d["perm_nonidentity"] = any ? 1 : 0;
  d["perm_cube_vertices"] = np2(std::vector<unsigned>(), 8);
  d["perm_tet_vertices"] = np2(tets, 4);
  if (!any) return;
  PermLookup look(DualSpy<T, R>::table(data));
  std::vector<unsigned> tp;
  tp.reserve(tets.size() * 4);
  for (size_t t = 0; t < tets.size() / 4; ++t)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) tp.push_back(look.row(tets[4 * t + a], tets[4 * t + b]));
  d["tet_perm"] = np2(tp, 16);
  d["cube_perm"] = np2(std::vector<unsigned>(), 64);
}

template <class T, class R>
static py::dict flatten_nest_data(const BrillouinZoneNest3<T, R, double>& g) {
  py::dict d;
  flatten_data_common(g, d);
  py::dict s = flatten_nest(g);
  auto nv = s["node_vertices"].cast<py::array_t<unsigned>>();
  std::vector<unsigned> tets(nv.data(), nv.data() + nv.size());
  flatten_tet_perms(g.data(), tets, d);
  return d;
}

// ---------------------------------------------------------------------------------------------
// Mesh (mesh.hpp, triangulation_layers.hpp): layers of tetrahedral meshes + layer-to-layer candidate lists
// ---------------------------------------------------------------------------------------------
template <class T, class R>
static py::dict flatten_mesh(const BrillouinZoneMesh3<T, R, double>& g) {
  py::dict d;
  d["kind"] = "mesh";
  d["bz"] = flatten_bz(g.get_brillouinzone());
  const TetTri& tt = MeshSpy<T, R>::mesh_of(g);
  const auto& layers = tt.*get(TetTriLayers());
  const auto& conns = tt.*get(TetTriConns());
  std::vector<unsigned> tet_off{0u}, vert_off{0u}, tets, conn_off{0u}, conn_idx;
  std::vector<double> centres, radii, vol6, verts;
  for (size_t l = 0; l < layers.size(); ++l) {
    const auto& L = layers[l];
    const auto& vpt = L.get_vertices_per_tetrahedron();
    const auto& cc = L.get_circum_centres();
    const auto& rr = L.get_circum_radii();
    const auto& vp = L.get_vertex_positions();
    for (ind_t t = 0; t < L.number_of_tetrahedra(); ++t) {
      for (ind_t j = 0; j < 4; ++j) tets.push_back(vpt.val(t, j));
      for (ind_t j = 0; j < 3; ++j) centres.push_back(cc.val(t, j));
      radii.push_back(rr[t]);
      vol6.push_back(6.0 * L.volume(t));  // triangulation_layers.hpp:267
    }
    for (ind_t v = 0; v < L.number_of_vertices(); ++v)
      for (ind_t j = 0; j < 3; ++j) verts.push_back(vp.val(v, j));
    tet_off.push_back(static_cast<unsigned>(radii.size()));
    vert_off.push_back(static_cast<unsigned>(verts.size() / 3));
    if (l + 1 < layers.size()) {
      const auto& map = conns[l];
      for (const auto& lst : map) {
        for (auto x : lst) conn_idx.push_back(x);
        conn_off.push_back(static_cast<unsigned>(conn_idx.size()));
      }
    }
  }
  d["n_layers"] = layers.size();
  d["tet_offset"] = np1(tet_off);
  d["vert_offset"] = np1(vert_off);
  d["tets"] = np2(tets, 4);
  d["centres"] = np2(centres, 3);
  d["radii"] = np1(radii);
  d["vol6"] = np1(vol6);
  d["vertices"] = np2(verts, 3);
  d["conn_offset"] = np1(conn_off);
  d["conn_index"] = np1(conn_idx);
  auto cfg = g.approx_config();
  d["approx_digit"] = cfg.digit();
  d["approx_direct"] = cfg.template direct<double>();
  d["approx_reciprocal"] = cfg.template reciprocal<double>();
  return d;
}

template <class T, class R>
static py::dict flatten_mesh_data(const BrillouinZoneMesh3<T, R, double>& g) {
  py::dict d;
  flatten_data_common(g, d);
  const TetTri& tt = MeshSpy<T, R>::mesh_of(g);
  const auto& vpt = tt.get_vertices_per_tetrahedron();  // finest layer
  std::vector<unsigned> tets;
  for (ind_t t = 0; t < vpt.size(0); ++t)
    for (ind_t j = 0; j < 4; ++j) tets.push_back(vpt.val(t, j));
  flatten_tet_perms(g.data(), tets, d);
  return d;
}

// ---------------------------------------------------------------------------------------------
// sort()   (interpolatordual.hpp:398-434): the connected vertex pairs and the cost configuration
// ---------------------------------------------------------------------------------------------
// The pairs (i < j) of the permutation table in key order -- exactly the list DualInterpolator::sort() walks -- and what
// Interpolator::add_cost needs to know (interpolator_cost.tpp:18-58, interpolator.hpp:246-305).
template <class T, class R>
static py::dict sort_plan(const DualInterpolator<T, R>& data) {
  py::dict d;
  const PermutationTable& table = DualSpy<T, R>::table(data);
  const size_t no = PermSpy::nidx(table);
  std::vector<unsigned> pairs;
  for (const auto& kv : PermSpy::map(table)) {  // std::map: ascending keys, like std::set<size_t> keys()
    const size_t key = kv.first;
    const size_t i = key / no;
    if (i * (no + 1) < key) {
      pairs.push_back(static_cast<unsigned>(i));
      pairs.push_back(static_cast<unsigned>(key - i * no));
    }
  }
  d["pairs"] = np2(pairs, 2);
  d["n_vertices"] = static_cast<unsigned>(no);
  const auto& vc = InterpSpy<T>::costmult(data.values());
  const auto& wc = InterpSpy<R>::costmult(data.vectors());
  d["values_costmult"] = np1(std::vector<double>{vc[0], vc[1], vc[2]});
  d["vectors_costmult"] = np1(std::vector<double>{wc[0], wc[1], wc[2]});
  d["values_vector_cost"] = static_cast<int>(InterpSpy<T>::funtype(data.values())[0]);
  d["vectors_vector_cost"] = static_cast<int>(InterpSpy<R>::funtype(data.vectors())[0]);
  return d;
}
// the permutations the host table holds for the ordered pairs (i,j) and (j,i) (identity when unset): (n_pairs, 2, modes)
template <class T, class R>
static py::array_t<unsigned> pair_permutations(const DualInterpolator<T, R>& data, py::array_t<unsigned, py::array::c_style | py::array::forcecast> pairs) {
  const PermutationTable& table = DualSpy<T, R>::table(data);
  PermLookup look(table);
  const auto& rows = PermSpy::perms(table);
  const size_t m = rows.empty() ? 0 : rows[0].size();
  const py::ssize_t n = pairs.shape(0);
  py::array_t<unsigned> out({n, static_cast<py::ssize_t>(2), static_cast<py::ssize_t>(m)});
  auto p = pairs.unchecked<2>();
  unsigned* o = out.mutable_data();
  for (py::ssize_t k = 0; k < n; ++k) {
    const auto& a = rows[look.row(p(k, 0), p(k, 1))];
    const auto& b = rows[look.row(p(k, 1), p(k, 0))];
    for (size_t e = 0; e < m; ++e) {
      o[(2 * k) * m + e] = a[e];
      o[(2 * k + 1) * m + e] = b[e];
    }
  }
  return out;
}

