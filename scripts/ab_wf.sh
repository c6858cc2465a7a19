#!/bin/bash
# Developer helper (GPU box): wavefront integrator timing for the given variants (SCENE / ARG select the scene).
cd "$(dirname "$0")/.."
for n in "$@"; do
  r=$(SRT_LIB_PATH=build/variants/libsrt_$n.so python scripts/quick_bench.py --scene ${SCENE:-cornell} --arg ${ARG:-0} --integrator 0 --frames 16 --reps 2 --profile 1 2>&1 | grep samples_per_s | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f M samples/s  %.2f ms  stages %s' % (d['samples_per_s']/1e6, d['device_ms'], [round(x,1) for x in d['stage_ms']]))")
  echo "$n: $r"
done
