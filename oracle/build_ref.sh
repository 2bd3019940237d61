#!/bin/bash
# TEST INFRASTRUCTURE.  The parity oracle "the reference itself" is brille's own host library, built unmodified by
# third_party/build_brille_host.sh (which this script runs first).  What is built HERE is only the probe: a small pybind11 module
# (oracle/probe.cpp, compiled against the reference headers, linked with the reference's library objects) that exposes what
# Python lacks -- operation indices of ir_moveinto, (vertex, weight) lists, node indices, the GammaTable -- through the
# reference's public C++ API.  Output: oracle/_ref/_probe*.so (git-ignored, travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/.." && pwd)"
REF="${BRILLE_REFERENCE:-/root/reference}"
TP="$ROOT/third_party"
[ -d "$REF/src" ] || { echo "build_ref.sh: reference sources not found at $REF (expected on the build container only)" >&2; exit 3; }
bash "$TP/build_brille_host.sh"
PY="${PYTHON:-python3}"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
PBINC="$($PY -c 'import pybind11;print(pybind11.get_include())')"
EXT="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
CXXFLAGS="-std=c++17 -O3 -DNDEBUG -include cassert -fopenmp -fPIC -w -I$TP/shim -I$TP/_build/gen -I$REF/src -I$REF/lib/tetgen -I$PYINC -I$PBINC"
LIBOBJS=$(ls "$TP"/_build/obj/*.o | grep -v '/_[a-z_0-9]*\.o$')
mkdir -p "$HERE/_ref"
TARGET="$HERE/_ref/_probe$EXT"
if [ ! -f "$TARGET" ] || [ "$HERE/probe.cpp" -nt "$TARGET" ]; then
  g++ $CXXFLAGS -c "$HERE/probe.cpp" -o "$HERE/_ref/probe.o"
  g++ -shared -fopenmp -o "$TARGET" "$HERE/_ref/probe.o" $LIBOBJS
fi
echo "built $TARGET"
