#!/bin/bash
# round 2, GPU call AF: source-level ncu capture of level 0's surface stage and of the combine/resolve kernel; knob combos
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_surface_spawn -s 9 -c 1 \
     -o gpurun_out/r2af_surface0 python tools/gpu_one_frame.py water_bunny_tex_soft_4k 2 > gpurun_out/r2af_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_combine_resolve -s 1 -c 1 \
     -o gpurun_out/r2af_combine python tools/gpu_one_frame.py water_bunny_tex_soft_4k 2 > gpurun_out/r2af_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python tools/gpu_share_sweep.py water_bunny_tex_soft_4k 8 1 > gpurun_out/r2af_sweep.log 2>&1
