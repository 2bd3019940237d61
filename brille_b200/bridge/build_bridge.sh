#!/bin/bash
# Compile the bridge (flatten.cpp) against the reference headers -> brille_b200/_bridge*.so
# Needs $BRILLE_REFERENCE (default /root/reference) and the object files of the host library built by
# oracle/build_ref.sh (the bridge links brille's non-wrapper objects, exactly as brille's own
# _brille module does).  On the GPU box the prebuilt .so is used.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${BRILLE_REFERENCE:-/root/reference}"
OUT="$ROOT/oracle/_ref"
[ -d "$REF/src" ] || { echo "build_bridge.sh: no reference sources at $REF" >&2; exit 3; }
[ -d "$OUT/obj" ] || { echo "build_bridge.sh: run oracle/build_ref.sh first" >&2; exit 3; }
PY="${PYTHON:-python3}"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
PBINC="$($PY -c 'import pybind11;print(pybind11.get_include())')"
EXT="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
CXXFLAGS="-std=c++17 -O2 -DNDEBUG -include cassert -fopenmp -fPIC -w -I$ROOT/oracle/shim -I$OUT/gen -I$REF/src -I$REF/lib/tetgen -I$PYINC -I$PBINC"
LIBOBJS=$(ls "$OUT"/obj/*.o | grep -v '/_[a-z_]*\.o$')
TARGET="$ROOT/brille_b200/_bridge$EXT"
if [ ! -f "$TARGET" ] || [ "$HERE/flatten.cpp" -nt "$TARGET" ]; then
  g++ $CXXFLAGS -c "$HERE/flatten.cpp" -o "$OUT/obj/_bridge.o"
  g++ -shared -fopenmp -o "$TARGET" "$OUT/obj/_bridge.o" $LIBOBJS
fi
echo "built $TARGET"
