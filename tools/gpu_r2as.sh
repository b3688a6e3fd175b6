#!/bin/bash
# round 2, GPU call AS: final checks of the committed build: GPU suite, smoke, rank share, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2as_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2as_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2as_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2as_smoke.log
python tools/gpu_rankshare.py > gpurun_out/r2as_rankshare.log 2>&1; cat gpurun_out/r2as_rankshare.log
python tools/gpu_launch_times.py water_bunny_tex_soft_4k 8 > gpurun_out/r2as_launch_times_soft_share8.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r2as_bench_short.json 2> gpurun_out/r2as_bench.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2as_bench_short.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'])"
