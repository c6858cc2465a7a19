#!/bin/bash
# Developer helper (GPU box): per-kernel time and instruction-cache counters of the wavefront integrator.
cd "$(dirname "$0")/.."
python scripts/quick_bench.py --scene ${SCENE:-spheres} --arg ${ARG:-10000} --integrator 0 --frames 8 --reps 2 --profile 1 2>&1 | grep samples_per_s | tail -1 | cut -c1-120
python scripts/quick_bench.py --scene ${SCENE:-spheres} --arg ${ARG:-10000} --integrator 0 --frames 8 --reps 2 --profile 1 2>&1 | grep -o '"stage_ms.*'
ncu --metrics gpu__time_duration.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:"k_shade|k_extend" -s 30 -c 6 --csv python scripts/quick_bench.py --scene ${SCENE:-spheres} --arg ${ARG:-10000} --integrator 0 --frames 4 --reps 1 --profile 0 2>/dev/null | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
cur=None
for r in rows:
    key=(r[0],r[4][:40])
    if key!=cur: print(); print(r[4][:60], end=': '); cur=key
    print(r[-3].split('__')[-1][:30]+'='+r[-1], end='  ')
print()
"
