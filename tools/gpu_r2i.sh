#!/bin/bash
# round 2, GPU call I: new prune rule, edge-plane filter, flattened filter kernel, compact ray work list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r2i_pytest.log
for w in water_bunny_tex_soft_4k bunny_shadow_4k config glass_bunny_soft_8k; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline --workload $w > gpurun_out/r2i_wl_$w.json 2>> gpurun_out/r2i_bench.err; echo "$w exit $?"
done
WRT_SOFT_FILTER=0 timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline > gpurun_out/r2i_var_nofilter.json 2>> gpurun_out/r2i_bench.err
python tools/gpu_rankshare.py > gpurun_out/r2i_share.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], d['config'].get('image_checksum'), d['config'].get('shadow_rays_traced'), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/r2i_share.log | tail -5
