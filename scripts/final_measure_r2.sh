#!/bin/bash
# Round-2 measurement pass on the GPU box (1 GPU): tests, contract bench (both arms), launch list, ncu captures.
cd "$(dirname "$0")/.."
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2_final.log 2>&1; tail -2 gpurun_out/pytest_gpu_r2_final.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_reference_cpu.json 2> gpurun_out/r02_bench_reference_cpu.err
python bench.py --integrator 0 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_wavefront.json 2> gpurun_out/r02_bench_wavefront.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 -o gpurun_out/r02_k_resident_v12 -f python scripts/quick_bench.py --integrator 1 --frames 16 --reps 1 --profile 0 > gpurun_out/ncu_res_v12.log 2>&1
# BVH wavefront at steady state: find the first full-pool iteration of the measured render (the longest k_extend), then capture it
BVH_CMD="python scripts/quick_bench.py --scene spheres --arg 10000 --integrator 0 --frames 64 --reps 1 --profile 0"
SKIP=$(ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_shade|k_extend|k_shadow" -c 600 --csv $BVH_CMD 2>/dev/null | python -c "
import sys,csv
rows=[(int(r[0]), r[4], float(r[-1].replace(',',''))) for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
ext=[(v,i) for i,n,v in rows if 'k_extend' in n]
best=max(ext)[1]
print(best)")
echo "BVH capture starts at matched launch $SKIP"
ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_extend|k_shadow" -s $SKIP -c 4 -o gpurun_out/r02_wavefront_bvh_c4_v3 -f $BVH_CMD > gpurun_out/ncu_bvh_v3.log 2>&1
ncu --set full --clock-control none -k regex:k_resolve -c 1 -o gpurun_out/r02_k_resolve_v3 -f python scripts/quick_bench.py --integrator 1 --frames 2 --reps 1 --profile 0 > gpurun_out/ncu_resolve.log 2>&1
ls -la gpurun_out/r02_*v3.ncu-rep gpurun_out/r02_k_resident_v12.ncu-rep
