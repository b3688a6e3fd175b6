#!/bin/bash
# round 2, GPU call B: device BVH build + packed upload; GPU suite; bench; executed-work counters (ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r2b_pytest.log
WRT_VERBOSE=1 python tools/gpu_one_frame.py water_bunny_tex_soft_4k 3 > gpurun_out/r2b_verbose.log 2>&1; tail -4 gpurun_out/r2b_verbose.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"
WRT_HOST_BVH=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-per-config --no-cpu-baseline > gpurun_out/r2b_hostbvh.json 2>> gpurun_out/r2b_bench.err; echo "hostbvh exit $?"
M="gpu__time_duration.sum,sm__cycles_elapsed.avg,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed.avg.per_cycle_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active"
for w in water_bunny_tex_soft_4k bunny_shadow_4k; do
  python tools/gpu_one_frame.py $w 2 > gpurun_out/r2b_plain_$w.log 2>&1 &&
  timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2b_exec_$w.csv python tools/gpu_one_frame.py $w 2 > gpurun_out/r2b_ncu_$w.log 2>&1
  echo "ncu $w exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
        for k,v in d.get('per_config',{}).items():
            print('   ', k, v.get('error') or (round(v['ms_per_step'],3), round(v['e2e']['ms_per_step'],3), v['launches_per_frame']))
    except Exception as e:
        print(f, 'ERR', e)
PY
