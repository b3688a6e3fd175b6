#!/bin/bash
# round 2, GPU call F: restructured list-pruning kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/r2f_pytest.log
for v in "X=1" "WRT_SOFT_FILTER=0" "WRT_SOFT_FILTER=2"; do
  env $v timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline > "gpurun_out/r2f_var_${v}.json" 2>> gpurun_out/r2f_bench.err; echo "$v exit $?"
done
for v in "X=1"; do
  env $v python tools/gpu_rankshare.py > "gpurun_out/r2f_share_${v}.log" 2>&1
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-per-config --no-cpu-baseline --workload glass_bunny_soft_8k > "gpurun_out/r2f_wl_8k.json" 2>> gpurun_out/r2f_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], d['config']['shadow_rays_traced'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -n 4 gpurun_out/r2f_share_*.log
