#!/bin/bash
# Developer helper (GPU box): time every build/variants/libsrt_*.so (or the names given) on the Cornell box.
cd "$(dirname "$0")/.."
names=${@:-$(ls build/variants | sed 's/libsrt_\(.*\)\.so/\1/')}
for n in $names; do
  r=$(SRT_LIB_PATH=build/variants/libsrt_$n.so python scripts/quick_bench.py --integrator 1 --frames ${FRAMES:-32} --reps 2 --profile 0 ${EXTRA} 2>&1 | grep samples_per_s | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f M samples/s  %.2f ms' % (d['samples_per_s']/1e6, d['device_ms']))")
  echo "$n: $r"
done
