#!/bin/bash
# Compile the bridge (flatten.cpp) against brille's headers -> brille_b200/_bridge*.so
# Needs $BRILLE_REFERENCE (default /root/reference: brille's sources) and the objects of brille's host library built by
# third_party/build_brille_host.sh (the bridge links brille's library objects, exactly as brille's own _brille module does).
# On the GPU box the prebuilt .so is used.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${BRILLE_REFERENCE:-/root/reference}"
TP="$ROOT/third_party"
[ -d "$REF/src" ] || { echo "build_bridge.sh: no brille sources at $REF" >&2; exit 3; }
[ -d "$TP/_build/obj" ] || { echo "build_bridge.sh: run third_party/build_brille_host.sh first" >&2; exit 3; }
PY="${PYTHON:-python3}"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
PBINC="$($PY -c 'import pybind11;print(pybind11.get_include())')"
EXT="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
CXXFLAGS="-std=c++17 -O2 -DNDEBUG -include cassert -fopenmp -fPIC -w -I$TP/shim -I$TP/_build/gen -I$REF/src -I$REF/lib/tetgen -I$REF/wrap -I$ROOT/include -I$PYINC -I$PBINC"
LIBOBJS=$(ls "$TP"/_build/obj/*.o | grep -v '/_[a-z_0-9]*\.o$')
mkdir -p "$ROOT/brille_b200/_build"
TARGET="$ROOT/brille_b200/_bridge$EXT"
if [ ! -f "$TARGET" ] || [ "$HERE/flatten.cpp" -nt "$TARGET" ] || [ "$HERE/flatten.hpp" -nt "$TARGET" ]; then
  g++ $CXXFLAGS -c "$HERE/flatten.cpp" -o "$ROOT/brille_b200/_build/bridge.o"
  g++ -shared -fopenmp -o "$TARGET" "$ROOT/brille_b200/_build/bridge.o" $LIBOBJS
fi
echo "built $TARGET"
# the pybind11 add-on module: brille's grid classes with the interpolation path on the GPU (accel/accel.cpp)
ACCEL="$ROOT/brille_b200/_accel$EXT"
if [ -f "$ROOT/brille_b200/accel/accel.cpp" ]; then
  if [ ! -f "$ACCEL" ] || [ "$ROOT/brille_b200/accel/accel.cpp" -nt "$ACCEL" ] || [ "$HERE/flatten.hpp" -nt "$ACCEL" ] || [ "$ROOT/include/brille_b200.h" -nt "$ACCEL" ]; then
    g++ $CXXFLAGS -I"$HERE" -c "$ROOT/brille_b200/accel/accel.cpp" -o "$ROOT/brille_b200/_build/accel.o"
    g++ -shared -fopenmp -o "$ACCEL" "$ROOT/brille_b200/_build/accel.o" $LIBOBJS -L"$ROOT/brille_b200" -lbrille_b200 -Wl,-rpath,'$ORIGIN'
  fi
  echo "built $ACCEL"
fi
