#!/bin/bash
# Developer helper (GPU box): time variants and collect the instruction-cache counters of k_resident.
cd "$(dirname "$0")/.."
for n in "$@"; do
  r=$(SRT_LIB_PATH=build/variants/libsrt_$n.so python scripts/quick_bench.py --integrator 1 --frames ${FRAMES:-32} --reps 2 --profile 0 --scene ${SCENE:-cornell} 2>&1 | grep samples_per_s | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f M samples/s' % (d['samples_per_s']/1e6))")
  m=$(SRT_LIB_PATH=build/variants/libsrt_$n.so ncu --metrics sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio --clock-control none -k regex:k_resident -s 1 -c 1 --csv python scripts/quick_bench.py --integrator 1 --frames 8 --reps 1 --profile 0 --scene ${SCENE:-cornell} 2>/dev/null | grep k_resident | python -c "
import sys,csv
for r in csv.reader(sys.stdin): print(r[-3].split('__')[-1][:38]+'='+r[-1], end='  ')
")
  echo "$n: $r  $m"
done
