#!/bin/bash
# round 2, GPU call C: refilling list kernel, 3 request queues, centroid-grid PLOC
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/r2c_pytest.log
for v in "X=1" "WRT_DEEP_SPLIT=8" "WRT_DEEP_SPLIT=2" "WRT_DEEP_SPLIT=4" "WRT_HOST_BVH=1" "WRT_TRACE_BLOCKS=8"; do
  env $v timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline > "gpurun_out/r2c_var_${v}.json" 2>> gpurun_out/r2c_bench.err; echo "$v exit $?"
done
for v in "X=1" "WRT_DEEP_SPLIT=8" "WRT_DEEP_SPLIT=2" "WRT_DEEP_SPLIT=5"; do
  env $v python tools/gpu_rankshare.py > "gpurun_out/r2c_share_${v}.log" 2>&1
done
for w in bunny_shadow_4k config; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline --workload $w > "gpurun_out/r2c_wl_${w}.json" 2>> gpurun_out/r2c_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -n 4 gpurun_out/r2c_share_*.log
