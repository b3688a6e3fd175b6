#!/bin/bash
# round 2, multi-GPU check (run with gpurun --gpus N): multi-GPU tests, torchrun bench at N, wrt --gpus N wall clock
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "multi_gpu or rank or tile or drop_in" > gpurun_out/r2w_pytest_n$N.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2w_pytest_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2w_bench_n$N.json 2> gpurun_out/r2w_bench_n$N.err; echo "bench N=$N exit $?"
timeout 300 python tools/gpu_multi_bench.py water_bunny_tex_soft_4k 10 > gpurun_out/r2w_multi_n$N.json 2> gpurun_out/r2w_multi_n$N.err; echo "multi exit $?"
python - "$N" <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2w_bench_n{n}.json').read().strip().splitlines()[-1])
print('N',d['n_gpus'],'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'checksum',d['config']['image_checksum'],'per-rank',d.get('per_rank_render_ms'))
print(open(f'gpurun_out/r2w_multi_n{n}.json').read()[:600])
PY
