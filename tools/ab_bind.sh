#!/bin/bash
# A/B of bind builds inside ONE gpurun call: tools/ab_bind.sh lib1.so lib2.so ...  (the in-tree library first)
for lib in "" "$@"; do
  for variant in default staged; do
    echo "=== lib=${lib:-default} variant=$variant"
    if [ "$variant" = staged ]; then export CVB_BIND_VARIANT=staged; else unset CVB_BIND_VARIANT; fi
    if [ -n "$lib" ]; then CLIFFORD_B200_LIB=$PWD/$lib python tools/bench_ops.py bind; else python tools/bench_ops.py bind; fi
  done
done
unset CVB_BIND_VARIANT
