#!/bin/bash
# round 2, GPU call AP: timeline of the frame as it really runs (two streams), whole frame and one rank's 1/8 share
mkdir -p gpurun_out
WRT_TIMING_OVERLAP=1 python tools/gpu_launch_times.py water_bunny_tex_soft_4k 8 > gpurun_out/r2ap_timeline_share8.log 2>&1
WRT_TIMING_OVERLAP=1 python tools/gpu_launch_times.py water_bunny_tex_soft_4k 1 > gpurun_out/r2ap_timeline_n1.log 2>&1
cat gpurun_out/r2ap_timeline_share8.log
