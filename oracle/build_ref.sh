#!/bin/bash
# Build oracle/_ref/ from the reference's OWN sources where they lie under /root/reference (build
# container only; outputs are git-ignored but travel to the GPU box):
#   libfastgl.so            <- shenfun/legendre/fastgl/fastgl.cpp            (Gauss-Legendre nodes)
#   refcy/la*.so, Matvec*.so, transforms*.so
#                           <- shenfun/optimization/cython/{la,Matvec,transforms}.pyx, cythonised as
#                              they are (setup.py:43-65 does the same with numpy's headers): the
#                              reference's compiled line solvers / mat-vecs / recurrence transforms,
#                              used to pin the numpy oracle (tests/test_oracle_refcy.py) and as the
#                              `cpu_baseline.kind = "reference"` legs of the solver configs in bench.py.
# No reference source is copied into the repository.
set -e
cd "$(dirname "$0")"
REF=/root/reference/shenfun
if [ ! -f $REF/legendre/fastgl/fastgl.cpp ]; then echo "reference sources not present; keeping prebuilt oracle/_ref"; exit 0; fi
mkdir -p _ref _ref/refcy
g++ -O2 -fPIC -shared -I$REF/legendre/fastgl fastgl_capi.cpp $REF/legendre/fastgl/fastgl.cpp -o _ref/libfastgl.so
echo "built oracle/_ref/libfastgl.so"
PY=${PYTHON:-python}
NPI=$($PY -c "import numpy; print(numpy.get_include())")
PYI=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
SUF=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
TMP=$(mktemp -d)
pids=""
for m in la Matvec transforms; do
  src=$REF/optimization/cython/$m.pyx
  out=_ref/refcy/$m$SUF
  if [ -f $out ] && [ $out -nt $src ]; then continue; fi
  ( $PY -m cython -3 --cplus -I $REF/optimization/cython $src -o $TMP/$m.cpp > $TMP/$m.log 2>&1 && \
    g++ -O3 -fPIC -shared -std=c++17 -w -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION -I$NPI -I$PYI $TMP/$m.cpp -o $out && echo "built oracle/$out" ) &
  pids="$pids $!"
done
for p in $pids; do wait $p || { echo "cythonising a reference kernel failed"; cat $TMP/*.log | tail -20; exit 1; }; done
touch _ref/refcy/__init__.py
rm -rf $TMP
