#!/bin/bash
# round 2, GPU call T: shortest list the filter takes on (WRT_FILTER_MIN) A/B
mkdir -p gpurun_out
P=$PWD/whittedstyle_raytracer_b200
for v in main fm1 fm2; do
  lib=$P/libwrt_cuda_$v.so; [ $v = main ] && lib=$P/libwrt_cuda.so
  WRT_CUDA_LIB=$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline > "gpurun_out/r2t_${v}_soft4k.json" 2>> gpurun_out/r2t_bench.err; echo "$v exit $?"
done
WRT_CUDA_LIB=$P/libwrt_cuda_fm1.so timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "soft or cull or fuzz" > gpurun_out/r2t_pytest_fm1.log 2>&1; tail -2 gpurun_out/r2t_pytest_fm1.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2t_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], d['config'].get('image_checksum'), d['config'].get('shadow_rays_traced'), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
    except Exception as e:
        print(f, 'ERR', e)
PY
