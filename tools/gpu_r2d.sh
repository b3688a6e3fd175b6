#!/bin/bash
# round 2, GPU call D: list kernel v3 (cooperative flush), 256-bit record loads A/B, queue split A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/r2d_pytest.log
P=$PWD/whittedstyle_raytracer_b200
for v in "X=1" "WRT_DEEP_SPLIT=8" "WRT_CUDA_LIB=$P/libwrt_cuda_noldg256.so" "WRT_HOST_BVH=1"; do
  n=$(echo $v | tr '/=' '__' | tail -c 40)
  env $v timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline > "gpurun_out/r2d_var_${n}.json" 2>> gpurun_out/r2d_bench.err; echo "$v exit $?"
done
for v in "X=1" "WRT_DEEP_SPLIT=8"; do
  env $v python tools/gpu_rankshare.py > "gpurun_out/r2d_share_${v}.log" 2>&1
done
for w in bunny_shadow_4k config; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline --workload $w > "gpurun_out/r2d_wl_${w}.json" 2>> gpurun_out/r2d_bench.err
  WRT_CUDA_LIB=$P/libwrt_cuda_noldg256.so timeout 300 python bench.py --steps 8 --warmup 3 --no-per-config --no-cpu-baseline --workload $w > "gpurun_out/r2d_wl_${w}_noldg256.json" 2>> gpurun_out/r2d_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -n 4 gpurun_out/r2d_share_*.log
