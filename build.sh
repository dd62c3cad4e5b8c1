#!/bin/bash
# Build libshenb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
SRC=shenfun_b200/csrc
OUT=shenfun_b200/libshenb200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
mkdir -p build
objs=""
pids=""
for f in $SRC/*.cu; do
  o=build/$(basename ${f%.cu}).o
  objs="$objs $o"
  if [ ! -f $o ] || [ $f -nt $o ] || [ -n "$(find $SRC include -name '*.cuh' -newer $o -o -name '*.h' -newer $o)" ]; then
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS} -c $f -o $o &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $objs -lcudart
echo "built $OUT"
