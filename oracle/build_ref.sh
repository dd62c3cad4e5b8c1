#!/bin/bash
# Build oracle/_ref/ from the reference's own C++ sources (build container only).
set -e
cd "$(dirname "$0")"
REF=/root/reference/shenfun/legendre/fastgl
if [ ! -f $REF/fastgl.cpp ]; then echo "reference sources not present; keeping prebuilt oracle/_ref"; exit 0; fi
mkdir -p _ref
g++ -O2 -fPIC -shared -I$REF fastgl_capi.cpp $REF/fastgl.cpp -o _ref/libfastgl.so
echo "built oracle/_ref/libfastgl.so"
