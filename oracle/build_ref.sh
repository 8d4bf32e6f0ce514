#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference (zoj613/htnorm) from the sources where
# they lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
# Nothing is copied: gcc reads src/htnorm.c, src/htnorm_distributions.c, src/htnorm_rng.c in place.
# The reference's BLAS/LAPACK (un-vendored, unpinned: Makefile:17 `-lblas -llapack`) is satisfied by the
# OpenBLAS 0.3.31.dev that ships inside scipy's wheel in this image (LP64, symbols prefixed `scipy_`).
# Flags follow the reference's Python build (setup.py:24: -O2 -std=c99), i.e. IEEE semantics, no fast-math.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${HTNORM_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
SP="$(python -c 'import scipy, os; print(os.path.dirname(os.path.dirname(scipy.__file__)))')"
BLAS="$(ls "$SP"/scipy.libs/libscipy_openblas-*.so | head -1)"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
DEFS=""
for s in ddot daxpy dtrmv dtrsv dsymv dgemv dgemm dsymm dsyrk dpotrf dpotrs dpotri dposv; do
  DEFS="$DEFS -D${s}_=scipy_${s}_"
done
gcc -O2 -std=c99 -fPIC -shared $DEFS -I"$REF/include" -I"$REF/src" \
    "$REF/src/htnorm.c" "$REF/src/htnorm_distributions.c" "$REF/src/htnorm_rng.c" \
    -o "$OUT/libhtnorm_ref.so" "$BLAS" -Wl,-rpath,"$(dirname "$BLAS")" -lm
# column-major build (what the R binding compiles, src/Makevars:1) for the HTNORM_COLMAJOR parity row
gcc -O2 -std=c99 -fPIC -shared -DHTNORM_COLMAJOR $DEFS -I"$REF/include" -I"$REF/src" \
    "$REF/src/htnorm.c" "$REF/src/htnorm_distributions.c" "$REF/src/htnorm_rng.c" \
    -o "$OUT/libhtnorm_ref_colmajor.so" "$BLAS" -Wl,-rpath,"$(dirname "$BLAS")" -lm
echo "$BLAS" > "$OUT/BLAS_USED.txt"
echo "build_ref: wrote $OUT/libhtnorm_ref.so (+ colmajor) against $BLAS"
