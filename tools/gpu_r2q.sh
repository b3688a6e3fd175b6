#!/bin/bash
# round 2, GPU call Q: deep-queue split at 1/8 and full frame
mkdir -p gpurun_out
for sp in 8 4 3 5; do
  echo "== WRT_DEEP_SPLIT=$sp"
  WRT_DEEP_SPLIT=$sp python tools/gpu_rankshare.py 2>&1 | head -2
done
for sp in 4 3; do
  echo "== hard WRT_DEEP_SPLIT=$sp"
  WRT_DEEP_SPLIT=$sp python tools/gpu_launch_times.py bunny_shadow_4k 8 2>/dev/null | head -1
done
echo "== hard default"; python tools/gpu_launch_times.py bunny_shadow_4k 8 2>/dev/null | head -1
