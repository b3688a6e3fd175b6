#!/bin/bash
# round 2, GPU call X (evidence, re-run after the container was replaced): GPU suite, bench line (all configs, cpu_baseline),
# reference arm, ncu launch list of the bench command, per-launch event times, executed-work counters of four workloads,
# ncu --set full of the traversal / list kernels (summarised on the box: the report is too large to travel back)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2x_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2x_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_reference.json 2>> gpurun_out/r2x_bench.err; echo "reference arm exit $?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r2x_bench_short.json 2>> gpurun_out/r2x_bench.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2x_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r2x_ncu_bench.log 2>&1
echo "launch list exit $?"
python tools/gpu_launch_times.py water_bunny_tex_soft_4k 1 > gpurun_out/r2x_launch_times_soft_n1.log 2>&1
python tools/gpu_launch_times.py water_bunny_tex_soft_4k 8 > gpurun_out/r2x_launch_times_soft_share8.log 2>&1
python tools/gpu_launch_times.py bunny_shadow_4k 1 > gpurun_out/r2x_launch_times_hard_n1.log 2>&1
python tools/gpu_rankshare.py > gpurun_out/r2x_rankshare.log 2>&1
M="gpu__time_duration.sum,sm__cycles_elapsed.avg,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed.avg.per_cycle_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active"
for w in water_bunny_tex_soft_4k bunny_shadow_4k f4_directional_4k f4_spheres_1k_4k; do
  python tools/gpu_one_frame.py $w 2 > gpurun_out/r2x_plain_$w.log 2>&1 &&
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2x_exec_$w.csv python tools/gpu_one_frame.py $w 2 > gpurun_out/r2x_ncu_$w.log 2>&1
  echo "ncu $w exit $?"
done
python tools/gpu_one_frame.py water_bunny_tex_soft_4k 2 > gpurun_out/r2x_plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_soft_list_rays|k_soft_lists|k_soft_filter|k_trace_closest|k_surface_spawn" -s 27 -c 27 -o /tmp/r2x_full python tools/gpu_one_frame.py water_bunny_tex_soft_4k 2 > gpurun_out/r2x_ncu_full.log 2>&1
echo "ncu full exit $?"
python tools/summarize_ncu.py /tmp/r2x_full.ncu-rep gpurun_out/r2x_full_summary.json > /dev/null 2> gpurun_out/r2x_summarize.err; echo "summary exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['launches_per_frame'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v})
for k,v in d.get('per_config',{}).items():
    print('   ', k, v.get('error') or (round(v['ms_per_step'],3), round(v['e2e']['ms_per_step'],3), v['launches_per_frame'], round(v['value'])))
PY
cat gpurun_out/r2x_bench_reference.json | cut -c1-300
du -sh gpurun_out
