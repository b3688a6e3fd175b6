#!/bin/bash
mkdir -p gpurun_out
export SWEEP_KNOBS="default:;side5:WRT_SIDE_BLOCKS=5;side6:WRT_SIDE_BLOCKS=6;side4:WRT_SIDE_BLOCKS=4"
for w in bunny_shadow_4k glass_bunny_soft_8k f4_spheres_1k_4k f4_directional_4k config; do
  echo "== $w"; timeout 300 python tools/gpu_share_sweep.py $w 8 1
done 2>&1 | tee gpurun_out/r2ag_sweep.log
