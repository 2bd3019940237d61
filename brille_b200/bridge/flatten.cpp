// brille_b200._bridge: the flattening of brille host objects (flatten.hpp) as a Python module -- used by the ctypes front end
// (brille_b200/grid.py), the tests and the fixture generator.
#include "flatten.hpp"

// ---------------------------------------------------------------------------------------------
// module
// ---------------------------------------------------------------------------------------------
template <class T, class R>
static void def_trellis(py::module& m) {
  using G = BrillouinZoneTrellis3<T, R, double>;
  m.def("flatten", [](const G& g) { return flatten_trellis(g); }, py::arg("grid"),
        "Structure tables (Brillouin zone + trellis) of a BZTrellisQ* object");
  m.def("flatten_data", [](const G& g) { return flatten_trellis_data(g); }, py::arg("grid"),
        "Data tables (values, vectors, permutations, gamma table) of a BZTrellisQ* object");
  m.def("sort_plan", [](const G& g) { return sort_plan(g.data()); }, py::arg("grid"),
        "Vertex pairs and cost configuration of DualInterpolator::sort()");
  m.def("pair_permutations", [](const G& g, py::array_t<unsigned, py::array::c_style | py::array::forcecast> p) { return pair_permutations(g.data(), p); },
        py::arg("grid"), py::arg("pairs"));
}

template <class T, class R>
static void def_nest_mesh(py::module& m) {
  using N = BrillouinZoneNest3<T, R, double>;
  using M = BrillouinZoneMesh3<T, R, double>;
  m.def("flatten", [](const N& g) { return flatten_nest(g); }, py::arg("grid"));
  m.def("flatten_data", [](const N& g) { return flatten_nest_data(g); }, py::arg("grid"));
  m.def("flatten", [](const M& g) { return flatten_mesh(g); }, py::arg("grid"));
  m.def("flatten_data", [](const M& g) { return flatten_mesh_data(g); }, py::arg("grid"));
  m.def("sort_plan", [](const N& g) { return sort_plan(g.data()); }, py::arg("grid"));
  m.def("sort_plan", [](const M& g) { return sort_plan(g.data()); }, py::arg("grid"));
  m.def("pair_permutations", [](const N& g, py::array_t<unsigned, py::array::c_style | py::array::forcecast> p) { return pair_permutations(g.data(), p); },
        py::arg("grid"), py::arg("pairs"));
  m.def("pair_permutations", [](const M& g, py::array_t<unsigned, py::array::c_style | py::array::forcecast> p) { return pair_permutations(g.data(), p); },
        py::arg("grid"), py::arg("pairs"));
}

PYBIND11_MODULE(_bridge, m) {
  m.doc() = "brille_b200 bridge: flatten brille host objects into SoA tables";
  m.def("flatten_bz", &flatten_bz, py::arg("bz"));
  def_trellis<double, double>(m);
  def_trellis<double, cplx>(m);
  def_trellis<cplx, cplx>(m);
  def_nest_mesh<double, double>(m);
  def_nest_mesh<double, cplx>(m);
  def_nest_mesh<cplx, cplx>(m);
}
