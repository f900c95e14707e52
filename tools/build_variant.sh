#!/bin/bash
# Build an A/B variant of the library: tools/build_variant.sh NAME "-DFLAG=1 ..." tu1.cu [tu2.cu ...]
# Only the named translation units are recompiled with the extra flags; the others are linked from the default build
# (run `make -C clifford-vae_b200/csrc` first).  Output: gpurun_variants/NAME.so (git-ignored, travels with gpurun);
# select it with CLIFFORD_B200_LIB=$PWD/gpurun_variants/NAME.so.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/clifford-vae_b200/csrc
NAME=$1; FLAGS=$2; shift 2
OUT=$ROOT/gpurun_variants; mkdir -p $OUT/obj_$NAME
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
OBJS=""
for f in $SRC/api_*.cu; do
  b=$(basename $f .cu)
  if [[ " $* " == *" $b.cu "* ]]; then
    ( cd $SRC && $NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $FLAGS -c $b.cu -o $OUT/obj_$NAME/$b.o ) &
    OBJS="$OBJS $OUT/obj_$NAME/$b.o"
  else
    OBJS="$OBJS $SRC/$b.o"
  fi
done
wait
$NVCC $ARCH -shared -o $OUT/$NAME.so $OBJS -lcudart
echo built $OUT/$NAME.so
