#!/bin/bash
# round 2, GPU call A: smoke, the whole GPU suite, the bench line, A/B of the new scheduling switches
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke exit $?"
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"
for v in "WRT_FUSE_FROM=1" "WRT_FUSE_FROM=0" "WRT_SHADE0_SEPARATE=0" "WRT_OVERLAP=0" "WRT_TRACE_BLOCKS=8" "WRT_SHAFT_LEVELS=8"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-per-config --no-cpu-baseline > "gpurun_out/r2a_var_${v}.json" 2>> gpurun_out/r2a_bench.err; echo "$v exit $?"
done
for v in "X=1" "WRT_FUSE_FROM=1" "WRT_OVERLAP=0"; do
  env $v python tools/gpu_rankshare.py > "gpurun_out/r2a_share_${v}.log" 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/r2a_share_*.log | tail -12
