#!/bin/bash
# round 2, GPU call AN: final bench line (reads the final exec metrics), launch list of the bench command, GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2an_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2an_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2an_bench.json 2> gpurun_out/r2an_bench.err; echo "bench exit $?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r2an_bench_short.json 2>> gpurun_out/r2an_bench.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2an_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r2an_ncu_bench.log 2>&1
echo "launch list exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2an_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
for k,v in d.get('per_config',{}).items():
    print('   ', k, v.get('error') or (round(v['ms_per_step'],3), round(v['e2e']['ms_per_step'],3), v['launches_per_frame'], round(v['value'])))
PY
