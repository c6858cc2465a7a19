set -x
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_v5.log 2>&1; tail -2 gpurun_out/pytest_gpu_v5.log
python bench.py --steps 16 --warmup 3 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; tail -c 600 gpurun_out/bench_v5.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_v5.json 2> gpurun_out/bench_ref_v5.err
python scripts/config_sweep.py > gpurun_out/configs_v5.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v5.csv python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_bench_v5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 -o gpurun_out/res_v8 -f python scripts/quick_bench.py --integrator 1 --frames 16 --reps 1 --profile 0 > gpurun_out/ncu_res_v8.log 2>&1
ls -la gpurun_out/res_v8.ncu-rep
