// brille_b200._accel -- brille's grid classes with the batched interpolation path on the GPU (pybind11 add-on module).
//
// This is the reference-side binding of the drop-in boundary: the lines INTEGRATION.md proposes for
// wrap/_common_grid.hpp:272-341,408-437 and wrap/_bz.cpp:378-520, compiled here against brille's own headers WITHOUT editing
// brille's sources.  For every grid class brille registers (BZ{Trellis,Nest,Mesh}Q{dd,dc,cc}: wrap/_trellis.hpp:29-101,
// _nest.hpp:29-69, _mesh.hpp:29-55) a C++ subclass is registered as a Python subclass of brille's own class.  It inherits
// everything -- constructors, properties, fill, node queries -- and overrides
//
//     ir_interpolate_at(Q, useparallel=False, threads=-1, do_not_move_points=False)      wrap/_common_grid.hpp:276-301
//     interpolate_at(Q, useparallel=False, threads=-1, do_not_move_points=False)          wrap/_common_grid.hpp:412-437 (trellis)
//
// with calls to the C ABI (include/brille_b200.h): the host object is flattened once (bridge/flatten.hpp) into the tables
// b200_grid_create / b200_grid_set_data copy to the device; fill / set_flags_weights run brille's host code and mark the
// device copy of the data stale; sort() (and the `sort` argument of the other two) solves the mode assignments on the device
// (b200_grid_sort_pairs) and writes them into brille's own permutation table.  The GIL is released around the C-ABI calls; outputs are fresh numpy arrays of the shapes and
// dtypes brille returns; non-zero return codes become RuntimeError with brille's own texts.
// BrillouinZone.isinside / moveinto / ir_moveinto / ir_moveinto_wedge (wrap/_bz.cpp:378-520) are replaced on brille's own class
// by versions that run the same device kernel (b200_moveinto) and return what brille returns (rotation MATRICES).
//
// brille_b200/dropin/_brille.py re-exports brille's module with the nine grid names bound to these classes, so that
// `import brille` (the shim package built by brille_b200/accel/build_package.sh) or brille_b200.install() gives user code and
// brille's own tests the accelerated classes under the unchanged names.
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/complex.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <array>
#include <complex>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "flatten.hpp"
#include "brille_b200.h"

namespace py = pybind11;
using namespace pybind11::literals;

namespace {

int default_device() {
  const char* e = std::getenv("BRILLE_B200_DEVICE");
  return e ? std::atoi(e) : 0;
}
[[noreturn]] void raise_last(int rc) {
  const char* m = b200_last_error();
  throw std::runtime_error(m && *m ? std::string(m) : "brille_b200 error " + std::to_string(rc));
}
void check(int rc) {
  if (rc != B200_OK) raise_last(rc);
}

// ---- bridge dictionaries -> C-ABI tables (the arrays stay owned by `keep` for the duration of the create / set_data call) ----
struct Keep {
  std::vector<py::object> objs;
  template <class T>
  const T* arr(py::handle h, size_t* count = nullptr) {
    auto a = py::array_t<T, py::array::c_style | py::array::forcecast>::ensure(py::reinterpret_borrow<py::object>(h));
    if (!a) throw std::runtime_error("brille_b200: table entry is not convertible to the expected array type");
    objs.push_back(a);
    if (count) *count = static_cast<size_t>(a.size());
    return a.size() ? a.data() : nullptr;
  }
};
template <class T, size_t N>
void fill_fixed(T (&dst)[N], py::handle h) {
  auto a = py::array_t<T, py::array::c_style | py::array::forcecast>::ensure(py::reinterpret_borrow<py::object>(h));
  if (!a || static_cast<size_t>(a.size()) != N) throw std::runtime_error("brille_b200: fixed-size table entry has the wrong length");
  for (size_t i = 0; i < N; ++i) dst[i] = a.data()[i];
}

b200_bz_tables_t pack_bz(const py::dict& d, Keep& k) {
  b200_bz_tables_t t{};
  t.transform_needed = d["transform_needed"].cast<int>();
  fill_fixed(t.P6t, d["P6t"]);
  fill_fixed(t.invPt, d["invPt"]);
  fill_fixed(t.w_recip_metric, d["w_recip_metric"]);
  fill_fixed(t.w_real_metric, d["w_real_metric"]);
  fill_fixed(t.o_recip_metric, d["o_recip_metric"]);
  fill_fixed(t.o_real_metric, d["o_real_metric"]);
  fill_fixed(t.to_xyz, d["to_xyz"]);
  t.w_recip_volume = d["w_recip_volume"].cast<double>();
  t.o_recip_volume = d["o_recip_volume"].cast<double>();
  size_t n = 0;
  t.pa = k.arr<double>(d["pa"], &n);
  t.n_faces = static_cast<int32_t>(n / 3);
  t.pb = k.arr<double>(d["pb"]);
  t.pc = k.arr<double>(d["pc"]);
  t.normals = k.arr<double>(d["normals"]);
  t.taus = k.arr<int32_t>(d["taus"]);
  t.tau_lens = k.arr<double>(d["tau_lens"]);
  t.ca = k.arr<double>(d["ca"]);
  t.cb = k.arr<double>(d["cb"]);
  t.cc = k.arr<double>(d["cc"]);
  t.wedge_normals = k.arr<double>(d["wedge_normals"], &n);
  t.n_wedge = static_cast<int32_t>(n / 3);
  t.no_ir_mirroring = d["no_ir_mirroring"].cast<int>();
  t.float_tolerance = d["float_tolerance"].cast<double>();
  t.approx_tolerance = d["approx_tolerance"].cast<int>();
  t.rotations = k.arr<int32_t>(d["rotations"], &n);
  t.n_ops = static_cast<int32_t>(n / 9);
  t.inverse_index = k.arr<int32_t>(d["inverse_index"]);
  t.identity_index = d["identity_index"].cast<int>();
  return t;
}

struct Structure {
  int kind = 0;
  b200_trellis_tables_t tr{};
  b200_nest_tables_t ne{};
  b200_mesh_tables_t me{};
  const void* ptr() const { return kind == B200_GRID_TRELLIS ? (const void*)&tr : kind == B200_GRID_NEST ? (const void*)&ne : (const void*)&me; }
};
Structure pack_structure(const py::dict& d, Keep& k) {
  Structure s;
  const std::string kind = d["kind"].cast<std::string>();
  size_t n = 0;
  if (kind == "trellis") {
    s.kind = B200_GRID_TRELLIS;
    auto& t = s.tr;
    const char* names[3] = {"knots0", "knots1", "knots2"};
    for (int i = 0; i < 3; ++i) {
      t.knots[i] = k.arr<double>(d[names[i]], &n);
      t.n_knots[i] = static_cast<int32_t>(n);
    }
    t.node_type = k.arr<uint8_t>(d["node_type"], &n);
    t.n_nodes = static_cast<uint32_t>(n);
    t.node_index = k.arr<uint32_t>(d["node_index"]);
    t.cube_vertices = k.arr<uint32_t>(d["cube_vertices"], &n);
    t.n_cubes = static_cast<uint32_t>(n / 8);
    t.poly_offsets = k.arr<uint32_t>(d["poly_offsets"], &n);
    t.n_polys = static_cast<uint32_t>(n ? n - 1 : 0);
    t.tet_vertices = k.arr<uint32_t>(d["tet_vertices"], &n);
    t.n_tets = static_cast<uint32_t>(n / 4);
    t.tet_circum = k.arr<double>(d["tet_circum"]);
    t.tet_volume = k.arr<double>(d["tet_volume"]);
    t.vertices = k.arr<double>(d["vertices"], &n);
    t.n_vertices = static_cast<uint32_t>(n / 3);
  } else if (kind == "nest") {
    s.kind = B200_GRID_NEST;
    auto& t = s.ne;
    t.node_vertices = k.arr<uint32_t>(d["node_vertices"]);
    t.node_circum = k.arr<double>(d["node_circum"]);
    t.node_volume = k.arr<double>(d["node_volume"], &n);
    t.n_nodes = static_cast<uint32_t>(n);
    t.node_is_leaf = k.arr<uint8_t>(d["node_is_leaf"]);
    t.child_begin = k.arr<uint32_t>(d["child_begin"]);
    t.child_end = k.arr<uint32_t>(d["child_end"]);
    t.vertices = k.arr<double>(d["vertices"], &n);
    t.n_vertices = static_cast<uint32_t>(n / 3);
    t.tolerance = d["approx_reciprocal"].cast<double>();
    t.digit = d["approx_digit"].cast<int>();
  } else {
    s.kind = B200_GRID_MESH;
    auto& t = s.me;
    t.n_layers = d["n_layers"].cast<uint32_t>();
    t.tet_offset = k.arr<uint32_t>(d["tet_offset"]);
    t.vert_offset = k.arr<uint32_t>(d["vert_offset"]);
    t.tets = k.arr<uint32_t>(d["tets"]);
    t.centres = k.arr<double>(d["centres"]);
    t.radii = k.arr<double>(d["radii"]);
    t.vol6 = k.arr<double>(d["vol6"]);
    t.vertices = k.arr<double>(d["vertices"]);
    t.conn_offset = k.arr<uint32_t>(d["conn_offset"]);
    t.conn_index = k.arr<uint32_t>(d["conn_index"]);
  }
  return s;
}

void pack_interp(b200_interp_desc_t& t, const py::dict& d, const std::string& p, Keep& k, size_t* n_rows) {
  py::object data = d[(p + "_data").c_str()];
  py::array raw = py::array::ensure(data);
  t.is_complex = raw.dtype().kind() == 'c' ? 1 : 0;
  size_t n = 0;
  if (t.is_complex) t.data = k.arr<std::complex<double>>(data, &n);
  else t.data = k.arr<double>(data, &n);
  t.branches = d[(p + "_branches").c_str()].cast<uint32_t>();
  fill_fixed(t.elements, d[(p + "_elements").c_str()]);
  t.rotates_like = d[(p + "_rotlike").c_str()].cast<int>();
  t.length_unit = d[(p + "_lenunit").c_str()].cast<int>();
  *n_rows = raw.ndim() >= 1 ? static_cast<size_t>(raw.shape(0)) : 0;
}
b200_data_tables_t pack_data(const py::dict& d, Keep& k, bool* filled) {
  b200_data_tables_t t{};
  size_t nv = 0, nw = 0, n = 0;
  pack_interp(t.values, d, "values", k, &nv);
  pack_interp(t.vectors, d, "vectors", k, &nw);
  t.n_vertices = static_cast<uint32_t>(nv);
  *filled = t.values.data != nullptr || t.vectors.data != nullptr;
  const bool sorted = d.contains("perm_nonidentity") && d["perm_nonidentity"].cast<int>() != 0;
  if (d.contains("perm_rows")) {
    t.perm_rows = k.arr<uint32_t>(d["perm_rows"], &n);
    t.n_perm_rows = t.values.branches ? static_cast<uint32_t>(n / t.values.branches) : 0;
  }
  if (sorted) {
    if (d.contains("cube_perm")) t.cube_perm = k.arr<uint32_t>(d["cube_perm"]);
    if (d.contains("tet_perm")) t.tet_perm = k.arr<uint32_t>(d["tet_perm"]);
  } else if (t.n_perm_rows > 1) {
    t.n_perm_rows = 1;
  }
  t.n_atoms = d.contains("gamma_natoms") ? d["gamma_natoms"].cast<uint32_t>() : 0u;
  if (t.n_atoms) {
    t.gamma_F0 = k.arr<uint32_t>(d["gamma_F0"]);
    t.gamma_vidx = k.arr<uint32_t>(d["gamma_vidx"]);
    t.gamma_vectors = k.arr<double>(d["gamma_vectors"], &n);
    t.n_gamma_vectors = static_cast<uint32_t>(n / 3);
  }
  if (d.contains("rot_cart")) t.rot_cart = k.arr<double>(d["rot_cart"]);
  return t;
}

// ---- one device-resident copy of a host grid ------------------------------------------------------------------------------
struct DeviceGrid {
  b200_grid_t* h = nullptr;
  int device = 0;
  bool data_current = false, filled = false;
  uint64_t launches_before = 0;
  ~DeviceGrid() {
    if (h) b200_grid_destroy(h);
  }
};

template <class T>
struct np_type { using type = T; };

// the accelerated subclass of one of brille's grid classes
template <class Base, class T, class R, int KIND>
struct Accel : Base {
  using Base::Base;
  explicit Accel(const Base& b) : Base(b) {}
  std::shared_ptr<DeviceGrid> dev;
  int device = default_device();

  py::dict structure_tables() const {
    if constexpr (KIND == B200_GRID_TRELLIS) return flatten_trellis(static_cast<const Base&>(*this));
    else if constexpr (KIND == B200_GRID_NEST) return flatten_nest(static_cast<const Base&>(*this));
    else return flatten_mesh(static_cast<const Base&>(*this));
  }
  py::dict data_tables() const {
    if constexpr (KIND == B200_GRID_TRELLIS) return flatten_trellis_data(static_cast<const Base&>(*this));
    else if constexpr (KIND == B200_GRID_NEST) return flatten_nest_data(static_cast<const Base&>(*this));
    else return flatten_mesh_data(static_cast<const Base&>(*this));
  }
  // flatten + upload what is missing or stale (GIL held: the flattening builds numpy arrays)
  DeviceGrid& ensure() {
    if (!dev) {
      auto g = std::make_shared<DeviceGrid>();
      g->device = device;
      Keep k;
      py::dict s = structure_tables();
      b200_bz_tables_t bz = pack_bz(s["bz"].cast<py::dict>(), k);
      Structure st = pack_structure(s, k);
      check(b200_grid_create(st.kind, &bz, st.ptr(), device, &g->h));
      dev = g;
    }
    if (!dev->data_current) {
      Keep k;
      py::dict d = data_tables();
      bool filled = false;
      b200_data_tables_t t = pack_data(d, k, &filled);
      if (filled) check(b200_grid_set_data(dev->h, &t));
      dev->filled = filled;
      dev->data_current = true;
    }
    return *dev;
  }
  void invalidate() {
    if (dev) dev->data_current = false;
  }
  // DualInterpolator::sort() (interpolatordual.hpp:398-434) with its parallel loop -- the cost matrix and the Jonker-Volgenant
  // assignment of every connected vertex pair -- on the device (b200_grid_sort_pairs), and determine_permutation_ij's
  // PermutationTable::overwrite(i, j, row) / (j, i, col) (:426-433) applied here in the order of the pairs, i.e. what the
  // reference does with one thread.  overwrite(i, j, vector) looks every permutation up by a linear scan of the list of distinct
  // permutations (permutation_table.hpp:173-183,215-226); the same list is kept here with a hash index next to it and the
  // pair is pointed at its entry with overwrite(i, j, index) (:164-170).  Real-valued eigenvectors stay with brille's host
  // sort() (the device refuses them: the reference's anti-phase of real data reads an uninitialised buffer, utilities.tpp:530).
  void sort_on_device() {
    if constexpr (!std::is_same<R, std::complex<double>>::value) {
      Base::sort();
      invalidate();
    } else {
      DeviceGrid& g = ensure();
      auto& dual = const_cast<std::remove_const_t<std::remove_reference_t<decltype(this->data())>>&>(this->data());
      py::dict plan = sort_plan(dual);
      auto pairs = plan["pairs"].cast<py::array_t<unsigned, py::array::c_style | py::array::forcecast>>();
      const size_t n_pairs = static_cast<size_t>(pairs.shape(0));
      if (!g.filled || n_pairs == 0) {  // nothing to assign: the reference's loop is empty too
        Base::sort();
        return;
      }
      b200_sort_config_t cfg{};
      fill_fixed(cfg.values_costmult, plan["values_costmult"]);
      fill_fixed(cfg.vectors_costmult, plan["vectors_costmult"]);
      cfg.values_vector_cost = plan["values_vector_cost"].cast<int>();
      cfg.vectors_vector_cost = plan["vectors_vector_cost"].cast<int>();
      const size_t modes = this->data().branches();
      std::vector<int32_t> row(n_pairs * modes), col(n_pairs * modes);
      int rc;
      {
        py::gil_scoped_release release;
        rc = b200_grid_sort_pairs(g.h, pairs.data(), n_pairs, &cfg, row.data(), col.data(), nullptr);
      }
      check(rc);
      brille::PermutationTable& table = DualSpy<T, R>::table(dual);
      auto& perms = PermSpy::perms(table);
      std::unordered_map<std::string, size_t> index;  // permutation bytes -> position of its FIRST occurrence in the list
      auto key_of = [modes](const brille::ind_t* p) { return std::string(reinterpret_cast<const char*>(p), modes * sizeof(brille::ind_t)); };
      for (size_t i = 0; i < perms.size(); ++i)
        if (perms[i].size() == modes) index.emplace(key_of(perms[i].data()), i);
      std::vector<brille::ind_t> v(modes);
      auto point_at = [&](size_t i, size_t j, const int32_t* src) {
        for (size_t e = 0; e < modes; ++e) v[e] = static_cast<brille::ind_t>(src[e]);
        auto found = index.emplace(key_of(v.data()), perms.size());
        if (found.second) perms.push_back(v);
        table.overwrite(i, j, found.first->second);
      };
      const unsigned* pr = pairs.data();
      for (size_t k = 0; k < n_pairs; ++k) {
        point_at(pr[2 * k], pr[2 * k + 1], row.data() + k * modes);
        point_at(pr[2 * k + 1], pr[2 * k], col.data() + k * modes);
      }
      invalidate();
    }
  }
  template <class X>
  static py::array_t<X> output(size_t n, const std::vector<unsigned>& stored_shape) {
    std::vector<py::ssize_t> sh{static_cast<py::ssize_t>(n)};
    for (size_t i = 1; i < stored_shape.size(); ++i) sh.push_back(static_cast<py::ssize_t>(stored_shape[i]));
    return py::array_t<X>(sh);
  }
  // ir != 0: ir_interpolate_at, else interpolate_at
  py::tuple run(py::array_t<double, py::array::c_style | py::array::forcecast> Q, bool no_move, int ir) {
    if (Q.ndim() != 2 || Q.shape(1) != 3) throw std::runtime_error("Interpolation requires one or more 3-vectors");
    DeviceGrid& g = ensure();
    if (!g.filled) throw std::runtime_error("The interpolation data must be filled before interpolating.");
    const size_t n = static_cast<size_t>(Q.shape(0));
    auto vs = this->data().values().shape();
    auto ws = this->data().vectors().shape();
    py::array_t<T> vals = output<T>(n, std::vector<unsigned>(vs.begin(), vs.end()));
    py::array_t<R> vecs = output<R>(n, std::vector<unsigned>(ws.begin(), ws.end()));
    int rc;
    {
      py::gil_scoped_release release;
      rc = (ir ? b200_ir_interpolate_at : b200_interpolate_at)(g.h, Q.data(), n, no_move ? B200_FLAG_NO_MOVE : 0u, vals.mutable_data(),
                                                              vecs.mutable_data(), nullptr);
    }
    check(rc);
    return py::make_tuple(vals, vecs);
  }

  // ---- device-resident consumers (the slot of the commented-out ir_interpolate_at_dw, wrap/_common_grid.hpp:343-405)
  using dvec = py::array_t<double, py::array::c_style | py::array::forcecast>;
  void set_structure_factor(py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> coef, py::object positions,
                            py::object q_transform, py::object debye_waller, bool conjugate) {
    DeviceGrid& g = ensure();
    b200_sf_config_t c{};
    c.n_atoms = static_cast<uint32_t>(coef.size());
    c.coef = reinterpret_cast<const double*>(coef.data());
    dvec pos, dw, qt;
    if (!positions.is_none()) {
      pos = dvec::ensure(positions);
      if (!pos || static_cast<size_t>(pos.size()) != 3u * c.n_atoms) throw std::runtime_error("positions must be (n_atoms, 3)");
      c.positions = pos.data();
    }
    if (!debye_waller.is_none()) {
      dw = dvec::ensure(debye_waller);
      if (!dw || static_cast<size_t>(dw.size()) != 9u * c.n_atoms) throw std::runtime_error("debye_waller must be (n_atoms, 3, 3)");
      c.debye_waller = dw.data();
    }
    for (int i = 0; i < 9; ++i) c.q_transform[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (!q_transform.is_none()) {
      qt = dvec::ensure(q_transform);
      if (!qt || qt.size() != 9) throw std::runtime_error("q_transform must be 3x3");
      for (int i = 0; i < 9; ++i) c.q_transform[i] = qt.data()[i];
    }
    c.conjugate = conjugate ? 1 : 0;
    check(b200_grid_set_structure_factor(g.h, &c));
  }
  DeviceGrid& ready() {
    DeviceGrid& g = ensure();
    if (!g.filled) throw std::runtime_error("The interpolation data must be filled before interpolating.");
    return g;
  }
  static void check_q(const dvec& Q) {
    if (Q.ndim() != 2 || Q.shape(1) != 3) throw std::runtime_error("Interpolation requires one or more 3-vectors");
  }
  py::tuple structure_factor(dvec Q, bool no_move) {
    check_q(Q);
    DeviceGrid& g = ready();
    const size_t n = static_cast<size_t>(Q.shape(0));
    auto vs = this->data().values().shape();
    py::array_t<T> vals = output<T>(n, std::vector<unsigned>(vs.begin(), vs.end()));
    py::array_t<double> sf({static_cast<py::ssize_t>(n), static_cast<py::ssize_t>(this->data().branches())});
    int rc;
    {
      py::gil_scoped_release release;
      rc = b200_ir_structure_factor(g.h, Q.data(), n, no_move ? B200_FLAG_NO_MOVE : 0u, vals.mutable_data(), sf.mutable_data());
    }
    check(rc);
    return py::make_tuple(vals, sf);
  }
  static b200_powder_config_t powder_config(std::array<double, 2> q_range, uint32_t n_qbins, std::array<double, 2> w_range, uint32_t n_wbins, int weight) {
    b200_powder_config_t c{};
    c.n_qbins = n_qbins;
    c.n_wbins = n_wbins;
    c.q_lo = q_range[0];
    c.q_hi = q_range[1];
    c.w_lo = w_range[0];
    c.w_hi = w_range[1];
    c.weight = weight;
    return c;
  }
  py::tuple powder_bin(dvec Q, std::array<double, 2> q_range, uint32_t n_qbins, std::array<double, 2> w_range, uint32_t n_wbins, int weight,
                       bool no_move) {
    check_q(Q);
    DeviceGrid& g = ready();
    const b200_powder_config_t c = powder_config(q_range, n_qbins, w_range, n_wbins, weight);
    py::array_t<double> hist({static_cast<py::ssize_t>(n_qbins), static_cast<py::ssize_t>(n_wbins)}), counts(static_cast<py::ssize_t>(n_qbins));
    std::fill_n(hist.mutable_data(), hist.size(), 0.0);
    std::fill_n(counts.mutable_data(), counts.size(), 0.0);
    int rc;
    {
      py::gil_scoped_release release;
      rc = b200_ir_powder_bin(g.h, Q.data(), static_cast<size_t>(Q.shape(0)), no_move ? B200_FLAG_NO_MOVE : 0u, &c, hist.mutable_data(),
                              counts.mutable_data());
    }
    check(rc);
    return py::make_tuple(hist, counts);
  }
  py::tuple powder_sweep(std::array<double, 2> q_range, uint32_t n_qbins, std::array<double, 2> w_range, uint32_t n_wbins, uint64_t n_dir,
                         uint64_t seed, int weight, py::object dir_range) {
    DeviceGrid& g = ready();
    const b200_powder_config_t c = powder_config(q_range, n_qbins, w_range, n_wbins, weight);
    uint64_t lo = 0, hi = n_dir;
    if (!dir_range.is_none()) {
      auto r = dir_range.cast<std::array<uint64_t, 2>>();
      lo = r[0];
      hi = r[1];
    }
    py::array_t<double> hist({static_cast<py::ssize_t>(n_qbins), static_cast<py::ssize_t>(n_wbins)}), counts(static_cast<py::ssize_t>(n_qbins));
    std::fill_n(hist.mutable_data(), hist.size(), 0.0);
    std::fill_n(counts.mutable_data(), counts.size(), 0.0);
    int rc;
    {
      py::gil_scoped_release release;
      rc = b200_ir_powder_sweep(g.h, &c, n_dir, seed, lo, hi, hist.mutable_data(), counts.mutable_data());
    }
    check(rc);
    return py::make_tuple(hist, counts);
  }
};

template <class Base, class T, class R, int KIND>
void declare(py::module& m, py::module& host, const char* name) {
  using A = Accel<Base, T, R, KIND>;
  py::object base = host.attr(name);
  py::class_<A, Base> cls(m, name, py::dynamic_attr(),
                          "brille's grid class of the same name (all of its interface is inherited) with ir_interpolate_at / interpolate_at "
                          "running on the GPU through the C ABI of brille_b200");
  if constexpr (KIND == B200_GRID_TRELLIS) {  // wrap/_trellis.hpp:37-38
    cls.def(py::init<brille::BrillouinZone, double, bool>(), "brillouin_zone"_a, "node_volume_fraction"_a = 0.1, "always_triangulate"_a = false);
    cls.def(py::init<brille::BrillouinZone, double, bool, brille::approx_float::Config>(), "brillouin_zone"_a, "node_volume_fraction"_a,
            "always_triangulate"_a, "approx_config"_a);
  } else if constexpr (KIND == B200_GRID_NEST) {  // wrap/_nest.hpp:37-38
    cls.def(py::init<brille::BrillouinZone, double, brille::ind_t>(), "brillouin_zone"_a, "max_volume"_a, "max_branchings"_a = 5);
    cls.def(py::init<brille::BrillouinZone, brille::ind_t, brille::ind_t>(), "brillouin_zone"_a, "number_density"_a, "max_branchings"_a = 5);
  } else {  // wrap/_mesh.hpp:36
    cls.def(py::init<brille::BrillouinZone, double, int, int>(), "brillouin_zone"_a, "max_size"_a = -1., "num_levels"_a = 3, "max_points"_a = -1);
  }
  cls.def(py::init([](const Base& b) { return std::make_unique<A>(b); }), "host_grid"_a, "wrap an existing brille grid object (shares its data)");
  // host methods that change what the device holds: run brille's own, then mark the device copy stale
  // (their trailing `sort` argument -- wrap/_common_grid.hpp:47,147,503: keyword or last positional -- is taken out and
  // honoured with the device sort afterwards)
  for (const char* meth : {"fill", "set_flags_weights"}) {
    py::object host_method = base.attr(meth);
    const bool is_fill = std::string(meth) == "fill";
    cls.attr(meth) = py::cpp_function(
        [host_method, is_fill](py::object self, py::args a, py::kwargs k) {
          bool sort = false;
          py::list pos;
          for (auto x : a) pos.append(x);
          py::dict kw;
          for (auto item : k)
            if (py::str(item.first).cast<std::string>() == "sort") sort = item.second.cast<bool>();
            else kw[item.first] = item.second;
          const size_t np = pos.size();
          if (np == 5 || (is_fill && np == 7)) {
            py::object last = pos[np - 1];
            if (py::isinstance<py::bool_>(last)) {
              sort = last.cast<bool>();
              pos.attr("pop")();
            }
          }
          py::object r = host_method(self, *py::tuple(pos), **kw);
          A& g = self.cast<A&>();
          g.invalidate();
          if (sort) g.sort_on_device();
          return r;
        },
        py::is_method(cls), py::name(meth), py::doc(py::str(host_method.attr("__doc__")).cast<std::string>().c_str()));
  }
  cls.def("sort", [](A& g) { g.sort_on_device(); },
          "Determine the equivalent-mode permutation of every connected pair of vertices (brille's sort(), wrap/_common_grid.hpp:486): "
          "cost matrices and assignments on the GPU, brille's permutation table updated with the result");
  const std::string doc = py::str(base.attr("ir_interpolate_at").attr("__doc__")).cast<std::string>();
  cls.def(
      "ir_interpolate_at",
      [](A& g, py::array_t<double, py::array::c_style | py::array::forcecast> Q, bool, int, bool no_move) { return g.run(Q, no_move, 1); }, "Q"_a,
      "useparallel"_a = false, "threads"_a = -1, "do_not_move_points"_a = false, doc.c_str());
  if constexpr (KIND == B200_GRID_TRELLIS)
    cls.def(
        "interpolate_at",
        [](A& g, py::array_t<double, py::array::c_style | py::array::forcecast> Q, bool, int, bool no_move) { return g.run(Q, no_move, 0); }, "Q"_a,
        "useparallel"_a = false, "threads"_a = -1, "do_not_move_points"_a = false);
  // what the GPU side adds
  cls.def_readwrite("device", &A::device, "index of the CUDA device the tables are uploaded to (before the first interpolation)");
  cls.def_property_readonly("gpu_launches", [](A& g) { return g.dev && g.dev->h ? b200_grid_launch_count(g.dev->h) : (uint64_t)0; },
                            "kernels launched on the GPU for this grid so far");
  cls.def_property_readonly("gpu_last_path", [](A& g) { return g.dev && g.dev->h ? b200_grid_last_path(g.dev->h) : 0u; });
  // device-resident consumers: the eigenvectors never leave the GPU (include/brille_b200.h: b200_grid_set_structure_factor ...)
  cls.def("set_structure_factor", &A::set_structure_factor, "coef"_a, "positions"_a = py::none(), "q_transform"_a = py::none(),
          "debye_waller"_a = py::none(), "conjugate"_a = true,
          "configure |sum_k coef_k e^{-qv.W_k.qv} e^{2 pi i Q.r_k} (qv . eps_k^[*])|^2 with qv = q_transform Q (one complex 3-vector per atom)");
  cls.def("ir_structure_factor", &A::structure_factor, "Q"_a, "do_not_move_points"_a = false,
          "(values, |F(Q, mode)|^2) of the interpolated, rotated eigenvectors, reduced on the device");
  cls.def("ir_powder_bin", &A::powder_bin, "Q"_a, "q_range"_a, "n_qbins"_a, "w_range"_a, "n_wbins"_a, "weight"_a = 0, "do_not_move_points"_a = false,
          "(hist, counts): |F|^2 of the given points binned on (|Q|, eigenvalue) on the device");
  cls.def("ir_powder_sweep", &A::powder_sweep, "q_range"_a, "n_qbins"_a, "w_range"_a, "n_wbins"_a, "n_dir"_a, "seed"_a = 0, "weight"_a = 0,
          "dir_range"_a = py::none(), "(hist, counts) of a powder sweep whose points are generated on the device: nothing per Q crosses PCIe");
  cls.def("host", [](const A& g) { return Base(static_cast<const Base&>(g)); }, "a plain brille object of the base class (the reference's CPU path)");
}

// ---- BrillouinZone methods on the device ------------------------------------------------------------------------------------
// A Brillouin zone has no grid: its tables go to the device with a one-node dummy trellis.  Handles are cached by the content of
// the tables (BrillouinZone objects are copied around by value).
struct BZCache {
  std::mutex mu;
  std::map<std::string, std::shared_ptr<DeviceGrid>> map;
};
BZCache& bz_cache() {
  static BZCache c;
  return c;
}
std::shared_ptr<DeviceGrid> bz_device(const brille::BrillouinZone& bz, py::dict* tables_out) {
  py::dict d = flatten_bz(bz);
  std::string key;
  for (auto item : d) {
    key += py::str(item.first).cast<std::string>();
    py::array a = py::array::ensure(item.second);
    if (a) {
      py::array c = py::array::ensure(a, py::array::c_style);
      key.append(static_cast<const char*>(c.data()), static_cast<size_t>(c.nbytes()));
    } else {
      key += py::str(item.second).cast<std::string>();
    }
  }
  if (tables_out) *tables_out = d;
  BZCache& c = bz_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.map.find(key);
  if (it != c.map.end()) return it->second;
  if (c.map.size() > 32) c.map.clear();
  Keep k;
  b200_bz_tables_t t = pack_bz(d, k);
  // dummy structure: one null node between two knots per axis
  const double knots[2] = {0.0, 1.0};
  const uint8_t ntype[1] = {B200_NODE_NULL};
  const uint32_t nidx[1] = {0xffffffffu}, poff[1] = {0u};
  const double vert[3] = {0.0, 0.0, 0.0};
  b200_trellis_tables_t tr{};
  for (int i = 0; i < 3; ++i) { tr.n_knots[i] = 2; tr.knots[i] = knots; }
  tr.n_nodes = 1; tr.node_type = ntype; tr.node_index = nidx; tr.poly_offsets = poff; tr.n_vertices = 1; tr.vertices = vert;
  auto g = std::make_shared<DeviceGrid>();
  g->device = default_device();
  check(b200_grid_create(B200_GRID_TRELLIS, &t, &tr, g->device, &g->h));
  c.map[key] = g;
  return g;
}
struct MoveOut {
  py::array_t<double> q;
  py::array_t<int> tau;
  std::vector<int32_t> ridx, invridx;
  std::vector<uint32_t> status;
};
MoveOut run_moveinto(const brille::BrillouinZone& bz, py::array_t<double, py::array::c_style | py::array::forcecast> Q, int ir, py::dict* tables) {
  if (Q.ndim() != 2 || Q.shape(1) != 3) throw std::runtime_error("The last dimension must have size 3");
  auto g = bz_device(bz, tables);
  const size_t n = static_cast<size_t>(Q.shape(0));
  MoveOut o;
  o.q = py::array_t<double>({static_cast<py::ssize_t>(n), static_cast<py::ssize_t>(3)});
  o.tau = py::array_t<int>({static_cast<py::ssize_t>(n), static_cast<py::ssize_t>(3)});
  o.ridx.assign(n, 0);
  o.invridx.assign(n, 0);
  o.status.assign(n, 0u);
  b200_probe_t p{};
  p.q_ir = o.q.mutable_data();
  p.tau = o.tau.mutable_data();
  p.ridx = o.ridx.data();
  p.invridx = o.invridx.data();
  p.status = o.status.data();
  int rc;
  {
    py::gil_scoped_release release;
    rc = b200_moveinto(g->h, Q.data(), n, ir, &p);
  }
  check(rc);
  return o;
}
py::array_t<int> matrices(const py::dict& tables, const std::vector<int32_t>& idx) {
  auto rot = py::array_t<int, py::array::c_style | py::array::forcecast>::ensure(tables["rotations"]);
  py::array_t<int> out({static_cast<py::ssize_t>(idx.size()), static_cast<py::ssize_t>(3), static_cast<py::ssize_t>(3)});
  int* o = out.mutable_data();
  for (size_t i = 0; i < idx.size(); ++i)
    for (int e = 0; e < 9; ++e) o[9 * i + e] = rot.data()[9 * (size_t)idx[i] + e];
  return out;
}
void patch_brillouinzone(py::module& host) {
  py::object cls = host.attr("BrillouinZone");
  using BZ = brille::BrillouinZone;
  using QA = py::array_t<double, py::array::c_style | py::array::forcecast>;
  auto keep_doc = [&](const char* name) { return py::str(cls.attr(name).attr("__doc__")).cast<std::string>(); };
  const std::string d0 = keep_doc("isinside"), d1 = keep_doc("moveinto"), d2 = keep_doc("ir_moveinto"), d3 = keep_doc("ir_moveinto_wedge");
  if (py::hasattr(cls, "host_isinside")) return;  // already patched
  cls.attr("host_isinside") = cls.attr("isinside");
  cls.attr("host_moveinto") = cls.attr("moveinto");
  cls.attr("host_ir_moveinto") = cls.attr("ir_moveinto");
  cls.attr("host_ir_moveinto_wedge") = cls.attr("ir_moveinto_wedge");
  cls.attr("isinside") = py::cpp_function(  // wrap/_bz.cpp:378-384
      [](const BZ& b, QA p) {
        MoveOut o = run_moveinto(b, p, 3, nullptr);
        py::array_t<bool> out(static_cast<py::ssize_t>(o.status.size()));
        for (size_t i = 0; i < o.status.size(); ++i) out.mutable_data()[i] = !(o.status[i] & B200_ST_OUTSIDE_BZ);
        return out;
      },
      py::is_method(cls), py::name("isinside"), "points"_a, py::doc(d0.c_str()));
  cls.attr("moveinto") = py::cpp_function(  // wrap/_bz.cpp:386-405
      [](const BZ& b, QA Q, int) {
        MoveOut o = run_moveinto(b, Q, 0, nullptr);
        return py::make_tuple(o.q, o.tau);
      },
      py::is_method(cls), py::name("moveinto"), "Q"_a, "threads"_a = 0, py::doc(d1.c_str()));
  cls.attr("ir_moveinto") = py::cpp_function(  // wrap/_bz.cpp:434-463
      [](const BZ& b, QA Q, int) {
        py::dict t;
        MoveOut o = run_moveinto(b, Q, 1, &t);
        return py::make_tuple(o.q, o.tau, matrices(t, o.ridx), matrices(t, o.invridx));
      },
      py::is_method(cls), py::name("ir_moveinto"), "Q"_a, "threads"_a = 0, py::doc(d2.c_str()));
  cls.attr("ir_moveinto_wedge") = py::cpp_function(  // wrap/_bz.cpp:498-520
      [](const BZ& b, QA Q, int) {
        py::dict t;
        MoveOut o = run_moveinto(b, Q, 2, &t);
        return py::make_tuple(o.q, matrices(t, o.ridx));
      },
      py::is_method(cls), py::name("ir_moveinto_wedge"), "Q"_a, "threads"_a = 0, py::doc(d3.c_str()));
}

}  // namespace

PYBIND11_MODULE(_accel, m) {
  m.doc() = "brille_b200: brille's grid classes with the interpolation path on the GPU (subclasses of brille's own classes)";
  // brille's module must be loaded first: its classes are the bases of the ones registered here
  py::module host = py::module::import("brille_b200.host").attr("get")().cast<py::module>();
  using D = double;
  using C = std::complex<double>;
  using namespace brille;
  declare<BrillouinZoneTrellis3<D, D, D>, D, D, B200_GRID_TRELLIS>(m, host, "BZTrellisQdd");
  declare<BrillouinZoneTrellis3<D, C, D>, D, C, B200_GRID_TRELLIS>(m, host, "BZTrellisQdc");
  declare<BrillouinZoneTrellis3<C, C, D>, C, C, B200_GRID_TRELLIS>(m, host, "BZTrellisQcc");
  declare<BrillouinZoneNest3<D, D, D>, D, D, B200_GRID_NEST>(m, host, "BZNestQdd");
  declare<BrillouinZoneNest3<D, C, D>, D, C, B200_GRID_NEST>(m, host, "BZNestQdc");
  declare<BrillouinZoneNest3<C, C, D>, C, C, B200_GRID_NEST>(m, host, "BZNestQcc");
  declare<BrillouinZoneMesh3<D, D, D>, D, D, B200_GRID_MESH>(m, host, "BZMeshQdd");
  declare<BrillouinZoneMesh3<D, C, D>, D, C, B200_GRID_MESH>(m, host, "BZMeshQdc");
  declare<BrillouinZoneMesh3<C, C, D>, C, C, B200_GRID_MESH>(m, host, "BZMeshQcc");
  m.def("patch_brillouinzone", [host]() mutable { patch_brillouinzone(host); },
        "replace BrillouinZone.isinside / moveinto / ir_moveinto / ir_moveinto_wedge of brille's class by the device versions "
        "(the originals stay available as host_isinside, ...)");
  m.def("unpatch_brillouinzone", [host]() {
    py::object cls = host.attr("BrillouinZone");
    if (!py::hasattr(cls, "host_isinside")) return;
    for (const char* n : {"isinside", "moveinto", "ir_moveinto", "ir_moveinto_wedge"}) {
      const std::string h = std::string("host_") + n;
      cls.attr(n) = cls.attr(h.c_str());
      py::delattr(cls, h.c_str());
    }
  }, "restore brille's own BrillouinZone methods");
  m.attr("GRID_CLASSES") = py::make_tuple("BZTrellisQdd", "BZTrellisQdc", "BZTrellisQcc", "BZNestQdd", "BZNestQdc", "BZNestQcc", "BZMeshQdd",
                                          "BZMeshQdc", "BZMeshQcc");
}
