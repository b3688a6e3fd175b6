#!/bin/bash
# round 2, GPU call S: cooperative leaf tests A/B
mkdir -p gpurun_out
P=$PWD/whittedstyle_raytracer_b200
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2s_pytest.log
for v in main c0; do
  lib=$P/libwrt_cuda_$v.so; [ $v = main ] && lib=$P/libwrt_cuda.so
  for w in water_bunny_tex_soft_4k bunny_shadow_4k config; do
    WRT_CUDA_LIB=$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline --workload $w > "gpurun_out/r2s_${v}_${w}.json" 2>> gpurun_out/r2s_bench.err; echo "$v $w exit $?"
  done
  WRT_CUDA_LIB=$lib python tools/gpu_launch_times.py water_bunny_tex_soft_4k 8 > gpurun_out/r2s_lt_${v}_soft8.log 2>&1
  WRT_CUDA_LIB=$lib python tools/gpu_launch_times.py bunny_shadow_4k 8 > gpurun_out/r2s_lt_${v}_hard8.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], d['config'].get('image_checksum'), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
    except Exception as e:
        print(f, 'ERR', e)
PY
for v in main c0; do for n in soft8 hard8; do echo == $v $n; grep "launch family" gpurun_out/r2s_lt_${v}_$n.log | awk '{printf "%s:%s ", $4, $5}'; echo; tail -2 gpurun_out/r2s_lt_${v}_$n.log | head -1; done; done
