#!/bin/bash
# Developer helper (GPU box): time every build/variants/libsrt_*.so (or the names given) on the Cornell box.
cd "$(dirname "$0")/.."
names=${@:-$(ls build/variants | sed 's/libsrt_\(.*\)\.so/\1/')}
for n in $names; do
  out=$(SRT_LIB_PATH=build/variants/libsrt_$n.so python scripts/quick_bench.py --integrator 1 --frames ${FRAMES:-32} --reps ${REPS:-3} --profile 0 ${EXTRA} 2>&1)
  r=$(echo "$out" | grep samples_per_s | python -c "
import sys,json
best=max((json.loads(l) for l in sys.stdin), key=lambda d: d['samples_per_s'])
print('%.1f M samples/s  %.2f ms  self_hit %.5f rays/sample %.4f' % (best['samples_per_s']/1e6, best['device_ms'], best['self_hit_frac'], best['rays_per_sample']))")
  m=$(echo "$out" | grep "mean rgb")
  echo "$n: $r | $m"
done
