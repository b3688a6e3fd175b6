#!/bin/bash
# round 2, GPU call J: two-stage filter, edge-stage threshold A/B
mkdir -p gpurun_out
P=$PWD/whittedstyle_raytracer_b200
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "soft or cull or fuzz or 8k or 4k" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2j_pytest.log
for v in main e8 e999; do
  lib=$P/libwrt_cuda_$v.so; [ $v = main ] && lib=$P/libwrt_cuda.so
  WRT_CUDA_LIB=$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline > "gpurun_out/r2j_var_${v}.json" 2>> gpurun_out/r2j_bench.err; echo "$v exit $?"
done
timeout 300 python bench.py --steps 6 --warmup 3 --no-per-config --no-cpu-baseline --workload glass_bunny_soft_8k > gpurun_out/r2j_wl_8k.json 2>> gpurun_out/r2j_bench.err
python tools/gpu_rankshare.py > gpurun_out/r2j_share.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], d['config'].get('image_checksum'), d['config'].get('shadow_rays_traced'), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -4 gpurun_out/r2j_share.log
