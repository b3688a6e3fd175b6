#!/bin/bash
# Development aid: time every library under whittedstyle_raytracer_b200/variants/ (and the in-tree build) on the two 4K
# workloads: tools/gpu_variant_time.py per library.  Output: gpurun_out/ab_<tag>.log
TAG=${1:-ab}
mkdir -p gpurun_out
for lib in whittedstyle_raytracer_b200/libwrt_cuda.so whittedstyle_raytracer_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  WRT_CUDA_LIB=$lib timeout 300 python tools/gpu_variant_time.py 2>&1 | tee -a gpurun_out/ab_$TAG.log
done
