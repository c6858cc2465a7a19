set -x
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_v6.log 2>&1; tail -2 gpurun_out/pytest_gpu_v6.log
python bench.py --steps 16 --warmup 3 > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; tail -c 600 gpurun_out/bench_v6.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_v6.json 2> gpurun_out/bench_ref_v6.err
python scripts/config_sweep.py > gpurun_out/configs_v6.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v6.csv python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_bench_v6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 -o gpurun_out/res_v9 -f python scripts/quick_bench.py --integrator 1 --frames 16 --reps 1 --profile 0 > gpurun_out/ncu_res_v9.log 2>&1
ls -la gpurun_out/res_v9.ncu-rep
