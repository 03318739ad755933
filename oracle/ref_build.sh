#!/bin/sh
# oracle/ref_build.sh <reference root>
#
# Compiles the REFERENCE's own CUDA kernels for sm_100a into oracle/_ref/ (git-ignored,
# travels to the GPU box) so the tests can compare libb200fe bit for bit with them and
# time them on the same B200 ("same-box baseline").  TEST INFRASTRUCTURE ONLY.
#
# The reference cannot be built as a whole here: every benchmarkNN.cc includes
# <Kokkos_Core.hpp> and Kokkos is neither vendored nor installed (no network).  Its kernels
# are plain CUDA at the top of each file, above run_test(); this recipe cuts that kernel
# block out of the sources WHERE THEY LIE (line range found at build time: first
# `template` line .. line before run_test) into oracle/_ref/*.inc and compiles it inside
# the thin launch wrappers oracle/ref_wrap_*.cu.  Nothing from the reference is committed.
# `extern __shared__ T shared[]` inside the reference templates cannot be instantiated for
# float and double in one translation unit, hence one object per dtype.
set -e
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
HOSTCXX=/usr/bin/g++; [ -x "$HOSTCXX" ] || HOSTCXX=g++
mkdir -p "$OUT"

cut_kernels() { # <file.cc> <out.inc>
  first=$(grep -n '^template' "$1" | head -1 | cut -d: -f1)
  rt=$(grep -n 'void run_test' "$1" | head -1 | cut -d: -f1)
  last=$((rt - 1))
  # run_test is preceded by its own `template <typename T>` line
  if sed -n "${last}p" "$1" | grep -q '^template'; then last=$((last - 1)); fi
  sed -n "${first},${last}p" "$1" > "$2"
}
cut_kernels "$REF/benchmark01/benchmark01.cc" "$OUT/b01_kernels.inc"
cut_kernels "$REF/benchmark02/benchmark02.cc" "$OUT/b02_kernels.inc"
cut_kernels "$REF/benchmark03/benchmark03.cc" "$OUT/b03_kernels.inc"
cut_kernels "$REF/benchmark04/benchmark04.cc" "$OUT/b04_kernels.inc"
cut_kernels "$REF/benchmark05/benchmark05.cc" "$OUT/b05_kernels.inc"

FLAGS="-gencode arch=compute_100a,code=sm_100a -ccbin $HOSTCXX -O3 -std=c++17 --extended-lambda -lineinfo -Xcompiler -fPIC -I$OUT -I$REF"
objs=""
for t in double float; do
  for w in bwd vec; do
    $NVCC $FLAGS -DREF_T=$t -DREF_SUF=$( [ $t = double ] && echo f64 || echo f32 ) -c "$HERE/ref_wrap_$w.cu" -o "$OUT/ref_${w}_$t.o"
    objs="$objs $OUT/ref_${w}_$t.o"
  done
done
$NVCC -gencode arch=compute_100a,code=sm_100a -ccbin $HOSTCXX -shared -o "$OUT/libref_kernels.so" $objs
rm -f $objs "$OUT"/*.inc   # the cut-out kernel text is a build intermediate, not kept
echo "oracle/_ref: built libref_kernels.so from $REF"
