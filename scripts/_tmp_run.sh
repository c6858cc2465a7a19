qb() { python scripts/quick_bench.py "$@" 2>&1 | grep samples_per_s | python -c "
import sys,json
best=max((json.loads(l) for l in sys.stdin), key=lambda d: d['samples_per_s']); print('%.1f M samples/s %.2f ms stages %s' % (best['samples_per_s']/1e6, best['device_ms'], [round(x,2) for x in best['stage_ms']]))"; }
for v in df0 df1; do for leaf in 4 2 1; do echo -n "spheres10k $v leaf$leaf: "; SRT_BVH_LEAF=$leaf SRT_LIB_PATH=build/variants/libsrt_$v.so qb --scene spheres --arg 10000 --integrator 0 --frames 16 --reps 3 --profile 1; done; done
for v in df0 df1; do echo -n "default resident $v: "; SRT_LIB_PATH=build/variants/libsrt_$v.so qb --scene default --integrator 1 --frames 32 --reps 3 --profile 0; echo -n "prism resident $v: "; SRT_LIB_PATH=build/variants/libsrt_$v.so qb --scene prism --integrator 1 --frames 32 --reps 3 --profile 0; done
SRT_LIB_PATH=build/variants/libsrt_df0.so ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_extend" -s 90 -c 2 -o gpurun_out/wf_bvh_r2b -f python scripts/quick_bench.py --scene spheres --arg 10000 --integrator 0 --frames 16 --reps 1 --profile 0 > gpurun_out/ncu_wf_bvh_r2b.log 2>&1
ls -la gpurun_out/wf_bvh_r2b.ncu-rep
