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
ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 -o gpurun_out/r02_k_resident_v11 -f python scripts/quick_bench.py --integrator 1 --frames 16 --reps 1 --profile 0 > gpurun_out/ncu_res_v11.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_extend|k_shadow" -s 120 -c 4 -o gpurun_out/r02_wavefront_bvh_c4_v2 -f python scripts/quick_bench.py --scene spheres --arg 10000 --integrator 0 --frames 32 --reps 1 --profile 0 > gpurun_out/ncu_bvh_v2.log 2>&1
ncu --set full --clock-control none -k regex:k_resolve -c 1 -o gpurun_out/r02_k_resolve_v1 -f python scripts/quick_bench.py --integrator 1 --frames 2 --reps 1 --profile 0 > gpurun_out/ncu_resolve.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
