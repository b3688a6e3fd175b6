#!/bin/bash
# round 2, GPU call AO: deep request queues 2 / 4 / 8 at one rank's share of an 8 / 4 / 2 / 1-rank frame
mkdir -p gpurun_out
export SWEEP_KNOBS="auto:;queues2:WRT_DEEP_QUEUES=2;queues4:WRT_DEEP_QUEUES=4;queues8:WRT_DEEP_QUEUES=8;queues1:WRT_DEEP_QUEUES=1"
for w in water_bunny_tex_soft_4k bunny_shadow_4k; do
  echo "== $w"; timeout 300 python tools/gpu_share_sweep.py $w 8 4 2 1
done 2>&1 | tee gpurun_out/r2ao_sweep.log
echo "== config"; timeout 300 python tools/gpu_share_sweep.py config 1 2>&1 | tee -a gpurun_out/r2ao_sweep.log
