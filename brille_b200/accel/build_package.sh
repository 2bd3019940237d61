#!/bin/bash
# Assemble the drop-in `brille` package: brille's own Python files (unmodified, copied from $BRILLE_REFERENCE/brille at build
# time into a git-ignored directory) around brille_b200/dropin/_brille.py, which binds the grid class names to the GPU
# subclasses of brille_b200._accel.  With brille_b200/dropin/site on sys.path, `import brille` is the accelerated brille.
# brille's own Python tests are copied next to it (also git-ignored) so that the GPU box, which has no /root/reference, can run
# them unmodified against the package (tests/test_dropin.py).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${BRILLE_REFERENCE:-/root/reference}"
[ -d "$REF/brille" ] || { echo "build_package.sh: no brille sources at $REF" >&2; exit 3; }
SITE="$ROOT/brille_b200/dropin/site"
rm -rf "$SITE"
mkdir -p "$SITE/brille" "$SITE/reference_tests"
cp "$REF"/brille/*.py "$SITE/brille/"
cp "$ROOT/brille_b200/dropin/_brille.py" "$SITE/brille/_brille.py"
cp "$REF"/wrap/tests/*.py "$REF"/wrap/tests/*.npz "$REF"/wrap/tests/*.json "$SITE/reference_tests/"
echo "assembled $SITE/brille (+ reference_tests)"
