#!/bin/bash
# round 2, GPU call G: launch-bound A/B (variant libraries selected through WRT_CUDA_LIB)
mkdir -p gpurun_out
P=$PWD/whittedstyle_raytracer_b200
for v in main f8 f10 f12 s4 s5; do
  lib=$P/libwrt_cuda_$v.so; [ $v = main ] && lib=$P/libwrt_cuda.so
  WRT_CUDA_LIB=$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline > "gpurun_out/r2g_var_${v}.json" 2>> gpurun_out/r2g_bench.err; echo "$v exit $?"
done
for v in main s4; do
  lib=$P/libwrt_cuda_$v.so; [ $v = main ] && lib=$P/libwrt_cuda.so
  WRT_CUDA_LIB=$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline --workload bunny_shadow_4k > "gpurun_out/r2g_hard_${v}.json" 2>> gpurun_out/r2g_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
    except Exception as e:
        print(f, 'ERR', e)
PY
