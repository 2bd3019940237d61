// oracle/probe.cpp -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
//
// A small pybind11 module compiled against the reference headers (by oracle/build_ref.sh) that exposes
// what brille's Python API hides but its public C++ API offers, so parity tests can compare the
// *intermediate* decisions of the path, not just the final arrays:
//   * point-group operation INDICES chosen by BrillouinZone::ir_moveinto      (bz_move.cpp:165-296)
//   * trellis node index / (vertex, weight) lists for a Cartesian point       (trellis_poly.hpp:245-255,436)
//   * the same for Nest and Mesh grids                                        (nest.hpp:339-346, mesh.hpp)
// Objects created through the `_brille` module can be passed straight in: pybind11 shares the
// registered C++ types between extension modules built with the same compiler/ABI.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <pybind11/numpy.h>
#include <pybind11/complex.h>
#include <complex>
#include "bz_trellis.hpp"
#include "bz_nest.hpp"
#include "bz_mesh.hpp"

namespace py = pybind11;
using namespace brille;
using cplx = std::complex<double>;

static Array2<double> to_a2(const py::array_t<double, py::array::c_style | py::array::forcecast>& Q) {
  auto b = Q.request();
  if (b.ndim != 2 || b.shape[1] != 3) throw std::runtime_error("expected an (n,3) float64 array");
  auto n = static_cast<ind_t>(b.shape[0]);
  Array2<double> a(n, 3u);
  const double* p = static_cast<const double*>(b.ptr);
  for (ind_t i = 0; i < n; ++i)
    for (ind_t j = 0; j < 3; ++j) a.val(i, j) = p[3 * i + j];
  return a;
}

// (q_ir rlu, q_ir xyz, tau, Ridx, invRidx) for every Q (rlu of the conventional lattice)
static py::tuple ir_moveinto_idx(const BrillouinZone& bz, py::array_t<double, py::array::c_style | py::array::forcecast> Q, int threads) {
  using namespace brille::lattice;
  auto a = to_a2(Q);
  ind_t n = a.size(0);
  LVec<double> Qv(LengthUnit::inverse_angstrom, bz.get_lattice(), a);
  LVec<double> q(LengthUnit::inverse_angstrom, bz.get_lattice(), n);
  LVec<int> tau(LengthUnit::inverse_angstrom, bz.get_lattice(), n);
  std::vector<size_t> r, ir;
  bz.ir_moveinto(Qv, q, tau, r, ir, threads);
  auto xyz = q.xyz();
  py::array_t<double> qo({(py::ssize_t)n, (py::ssize_t)3}), xo({(py::ssize_t)n, (py::ssize_t)3});
  py::array_t<int> to({(py::ssize_t)n, (py::ssize_t)3});
  py::array_t<int> ro((py::ssize_t)n), iro((py::ssize_t)n);
  for (ind_t i = 0; i < n; ++i) {
    for (ind_t j = 0; j < 3; ++j) {
      qo.mutable_at(i, j) = q.val(i, j);
      xo.mutable_at(i, j) = xyz.val(i, j);
      to.mutable_at(i, j) = tau.val(i, j);
    }
    ro.mutable_at(i) = static_cast<int>(r[i]);
    iro.mutable_at(i) = static_cast<int>(ir[i]);
  }
  return py::make_tuple(qo, xo, to, ro, iro);
}

// (q rlu, tau) from BrillouinZone::moveinto
static py::tuple moveinto(const BrillouinZone& bz, py::array_t<double, py::array::c_style | py::array::forcecast> Q, int threads) {
  using namespace brille::lattice;
  auto a = to_a2(Q);
  ind_t n = a.size(0);
  LVec<double> Qv(LengthUnit::inverse_angstrom, bz.get_lattice(), a);
  LVec<double> q(LengthUnit::inverse_angstrom, bz.get_lattice(), n);
  LVec<int> tau(LengthUnit::inverse_angstrom, bz.get_lattice(), n);
  bz.moveinto(Qv, q, tau, threads);
  py::array_t<double> qo({(py::ssize_t)n, (py::ssize_t)3});
  py::array_t<int> to({(py::ssize_t)n, (py::ssize_t)3});
  for (ind_t i = 0; i < n; ++i)
    for (ind_t j = 0; j < 3; ++j) {
      qo.mutable_at(i, j) = q.val(i, j);
      to.mutable_at(i, j) = tau.val(i, j);
    }
  return py::make_tuple(qo, to);
}

// per point: number of vertices, vertex indices (padded with 0xffffffff) and weights (padded 0)
template <class G>
static py::tuple indices_weights(const G& g, py::array_t<double, py::array::c_style | py::array::forcecast> X) {
  auto a = to_a2(X);
  ind_t n = a.size(0);
  py::array_t<int> cnt((py::ssize_t)n);
  py::array_t<unsigned> idx({(py::ssize_t)n, (py::ssize_t)8});
  py::array_t<double> wgt({(py::ssize_t)n, (py::ssize_t)8});
  for (ind_t i = 0; i < n; ++i) {
    std::vector<std::pair<ind_t, double>> iw;
    try {
      iw = g.indices_weights(a.view(i));
    } catch (const std::exception&) {
      iw.clear();
    }
    cnt.mutable_at(i) = static_cast<int>(iw.size());
    for (size_t j = 0; j < 8; ++j) {
      idx.mutable_at(i, j) = j < iw.size() ? iw[j].first : 0xffffffffu;
      wgt.mutable_at(i, j) = j < iw.size() ? iw[j].second : 0.;
    }
  }
  return py::make_tuple(cnt, idx, wgt);
}

// Mesh3 keeps its TetTri protected and offers no per-point accessor: reach it through a derived class
template <class T, class R>
struct MeshProbe : Mesh3<T, R, double, Array2> {
  static const TetTri& mesh_of(const Mesh3<T, R, double, Array2>& m) { return m.*(&MeshProbe::mesh); }
};
template <class T, class R>
static py::tuple mesh_indices_weights(const BrillouinZoneMesh3<T, R, double>& g, py::array_t<double, py::array::c_style | py::array::forcecast> X) {
  auto a = to_a2(X);
  ind_t n = a.size(0);
  const TetTri& tt = MeshProbe<T, R>::mesh_of(g);
  py::array_t<int> cnt((py::ssize_t)n);
  py::array_t<unsigned> idx({(py::ssize_t)n, (py::ssize_t)8});
  py::array_t<double> wgt({(py::ssize_t)n, (py::ssize_t)8});
  for (ind_t i = 0; i < n; ++i) {
    std::vector<std::pair<ind_t, double>> iw;
    try {
      iw = tt.locate(a.view(i));
    } catch (const std::exception&) {
      iw.clear();
    }
    cnt.mutable_at(i) = static_cast<int>(iw.size());
    for (size_t j = 0; j < 8; ++j) {
      idx.mutable_at(i, j) = j < iw.size() ? iw[j].first : 0xffffffffu;
      wgt.mutable_at(i, j) = j < iw.size() ? iw[j].second : 0.;
    }
  }
  return py::make_tuple(cnt, idx, wgt);
}

template <class G>
static py::array_t<unsigned> node_index(const G& g, py::array_t<double, py::array::c_style | py::array::forcecast> X) {
  auto a = to_a2(X);
  ind_t n = a.size(0);
  py::array_t<unsigned> out((py::ssize_t)n);
  for (ind_t i = 0; i < n; ++i) out.mutable_at(i) = g.node_index(a.view(i));
  return out;
}

template <class G>
static std::vector<ind_t> permutation(const G& g, ind_t i, ind_t j) {
  return g.data().get_permutation(i, j);
}

template <class T, class R>
static void def_all(py::module& m) {
  using Tr = BrillouinZoneTrellis3<T, R, double>;
  using Ne = BrillouinZoneNest3<T, R, double>;
  m.def("indices_weights", &indices_weights<Tr>, py::arg("grid"), py::arg("x_xyz"));
  m.def("indices_weights", &indices_weights<Ne>, py::arg("grid"), py::arg("x_xyz"));
  m.def("node_index", &node_index<Tr>, py::arg("grid"), py::arg("x_xyz"));
  m.def("indices_weights", &mesh_indices_weights<T, R>, py::arg("grid"), py::arg("x_xyz"));
  m.def("permutation", &permutation<Tr>);
  m.def("permutation", &permutation<Ne>);
}

PYBIND11_MODULE(_probe, m) {
  m.doc() = "reference-internals probe for parity tests (test infrastructure)";
  m.def("ir_moveinto_idx", &ir_moveinto_idx, py::arg("bz"), py::arg("Q"), py::arg("threads") = 1);
  m.def("moveinto", &moveinto, py::arg("bz"), py::arg("Q"), py::arg("threads") = 1);
  def_all<double, double>(m);
  def_all<double, cplx>(m);
  def_all<cplx, cplx>(m);
}
