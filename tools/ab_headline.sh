#!/bin/bash
# A/B of library variants inside ONE gpurun call: tools/ab_headline.sh [what] lib1.so lib2.so ...  (the in-tree library first)
# what = bind | clifford : which tools/bench_ops.py section to run beside the headline bench legs
what=$1; shift
for lib in "" "$@"; do
  echo "=== lib=${lib:-default}"
  if [ -n "$lib" ]; then export CLIFFORD_B200_LIB=$PWD/$lib; else unset CLIFFORD_B200_LIB; fi
  python tools/bench_ops.py $what 2>&1 | grep -v "inject"
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-vae-step --no-other-configs | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('headline value %.4e  rsample_kl %.5f ms (%.3f)  bind %.5f ms (%.3f)  fused %.5f ms' % (d['value'], k['rsample_kl']['ms'], k['rsample_kl']['frac'], k['bind']['ms'], k['bind']['frac'], k['fused_one_launch_variant']['ms']))"
done
