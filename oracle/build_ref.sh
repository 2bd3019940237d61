#!/bin/bash
# Build the UNMODIFIED reference (brille) Python module from the sources where they lie under
# $BRILLE_REFERENCE (default /root/reference) into oracle/_ref/ .  Test infrastructure only:
# the result is (1) the host C++ that constructs lattices / Brillouin zones / grids, (2) the parity
# oracle and (3) the CPU baseline of bench.py.  No reference source is copied into this repository;
# only compiled objects and the final shared objects land under oracle/_ref/ (git-ignored).
#
# The reference's own build system (cmake + conan + HighFive/HDF5 + Catch2) cannot run in this image
# (no network, no HDF5), so the 39 translation units are compiled directly with g++ and the no-op
# HighFive shim in oracle/shim/.  See DESIGN.md "Oracle".
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${BRILLE_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: reference sources not found at $REF (expected on the build container only)" >&2
  exit 3
fi
PY="${PYTHON:-python3}"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
PBINC="$($PY -c 'import pybind11;print(pybind11.get_include())')"
EXT="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
mkdir -p "$OUT/obj" "$OUT/gen" "$OUT/brille_host"
# version.hpp is configured by cmake in the reference build (version.hpp.in) -> fill the tokens
sed -e 's/@BRILLE_HASH@/oracle/;s/@BRILLE_BRANCH@/oracle/;s/@BRILLE_CONFIGURE_TIME@/none/' \
    -e 's/@BRILLE_SAFE_VERSION@/0.0.0/;s/@BRILLE_HOSTNAME@/oracle/;s/@BRILLE_VERSION@/0.0.0+oracle/' \
    "$REF/version.hpp.in" > "$OUT/gen/version.hpp"
CXXFLAGS="-std=c++17 -O3 -DNDEBUG -include cassert -fopenmp -fPIC -w -I$HERE/shim -I$OUT/gen -I$REF/src -I$REF/lib/tetgen -I$PYINC -I$PBINC"
SRCS=$(ls "$REF"/wrap/_*.cpp; \
       for f in approx_config bravais bz bz_move bz_wedge comparisons debug hall_symbol hdf_interface neighbours \
                polyhedron_faces pointgroup pointsymmetry process_id spg_database symmetry vertex_map_set; do echo "$REF/src/$f.cpp"; done; \
       echo "$REF/lib/tetgen/tetgen.cxx"; echo "$REF/lib/tetgen/predicates.cxx")
compile_one() {
  src="$1"; obj="$OUT/obj/$(basename "${src%.*}").o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then
    g++ $CXXFLAGS -c "$src" -o "$obj" || { echo "FAILED $src" >&2; exit 1; }
  fi
}
export -f compile_one; export OUT CXXFLAGS
echo "$SRCS" | xargs -P "$JOBS" -I{} bash -c 'compile_one {}'
LIBOBJS=$(ls "$OUT"/obj/*.o | grep -v '/_[a-z_]*\.o$' || true)
WRAPOBJS=$(ls "$OUT"/obj/_*.o | grep -v '_probe.o' )
g++ -shared -fopenmp -o "$OUT/brille_host/_brille$EXT" $WRAPOBJS $LIBOBJS
# probe: exposes internals Python lacks (node index, vertex/weight lists, rotation indices, gamma table)
if [ ! -f "$OUT/brille_host/_probe$EXT" ] || [ "$HERE/probe.cpp" -nt "$OUT/brille_host/_probe$EXT" ]; then
  g++ $CXXFLAGS -c "$HERE/probe.cpp" -o "$OUT/obj/_probe.o"
  g++ -shared -fopenmp -o "$OUT/brille_host/_probe$EXT" "$OUT/obj/_probe.o" $LIBOBJS
fi
echo "built $OUT/brille_host/_brille$EXT and _probe$EXT"
