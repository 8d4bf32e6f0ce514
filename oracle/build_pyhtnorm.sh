#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles the reference's OWN Cython wrapper (pyhtnorm/_htnorm.pyx, unmodified, read in
# place from /root/reference) against THIS repo's headers (include/, incl. the htnorm_distributions.h shim) and links
# it to htnorm_b200/libhtnorm_b200.so instead of the reference's C sources: the boundary-compatibility client of
# SURVEY 8(f) N2.  Output (git-ignored, travels to the GPU box with gpurun):
#   oracle/_ref/pyhtnorm_b200/pyhtnorm/_htnorm*.so   the extension module
#   oracle/_ref/pyhtnorm_b200/pyhtnorm/__init__.py    two re-exports (the reference's __init__ also imports a
#                                                      setuptools_scm-generated _version module that does not exist here)
#   oracle/_ref/pyhtnorm_b200/test_pyhtnorm.py        the reference's own pytest file, staged byte for byte so that the
#                                                      GPU box (no /root/reference there) can run it UNCHANGED
#   oracle/_ref/pyhtnorm_b200/conftest.py             np.alltrue -> np.all (removed in NumPy 2; reference test :48,93)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${HTNORM_REFERENCE:-/root/reference}"
OUT="$HERE/_ref/pyhtnorm_b200"
if [ ! -f "$REF/pyhtnorm/_htnorm.pyx" ]; then
  echo "build_pyhtnorm: $REF not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
if [ ! -f "$ROOT/htnorm_b200/libhtnorm_b200.so" ]; then
  echo "build_pyhtnorm: build htnorm_b200/libhtnorm_b200.so first" >&2
  exit 1
fi
mkdir -p "$OUT/pyhtnorm"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
python -m cython -3 "$REF/pyhtnorm/_htnorm.pyx" -o "$TMP/_htnorm.c" >/dev/null
PYINC="$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
NPINC="$(python -c 'import numpy; print(numpy.get_include())')"
EXT="$(python -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
gcc -O2 -fPIC -shared -w -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION -I"$ROOT/include" -I"$PYINC" -I"$NPINC" \
    "$TMP/_htnorm.c" -o "$OUT/pyhtnorm/_htnorm$EXT" \
    -L"$ROOT/htnorm_b200" -lhtnorm_b200 -Wl,-rpath,'$ORIGIN/../../../../htnorm_b200'
printf 'from ._htnorm import hyperplane_truncated_mvnorm, structured_precision_mvnorm  # noqa: F401\n' > "$OUT/pyhtnorm/__init__.py"
rm -f "$OUT/test_pyhtnorm.py"
cp "$REF/tests/test_pyhtnorm.py" "$OUT/test_pyhtnorm.py"
chmod u+w "$OUT/test_pyhtnorm.py"
cat > "$OUT/conftest.py" <<'PY'
# NumPy 2 removed np.alltrue, which the reference's tests call (tests/test_pyhtnorm.py:48,93): alias, nothing else.
import numpy as np
if not hasattr(np, "alltrue"):
    np.alltrue = np.all
PY
echo "build_pyhtnorm: wrote $OUT (reference wrapper linked to libhtnorm_b200.so)"
