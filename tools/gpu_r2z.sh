#!/bin/bash
# round 2, GPU call Z: source-level ncu capture of ONE deep closest-hit launch (level 7 of the second frame), in-tree build and
# the deferred-leaf variant
mkdir -p gpurun_out
V=whittedstyle_raytracer_b200/variants
for tag in base defer1; do
  lib=whittedstyle_raytracer_b200/libwrt_cuda.so; [ $tag = defer1 ] && lib=$V/libwrt_defer1.so
  WRT_CUDA_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_closest -s 16 -c 1 \
     -o gpurun_out/r2z_trace7_$tag python tools/gpu_one_frame.py water_bunny_tex_soft_4k 2 > gpurun_out/r2z_ncu_$tag.log 2>&1
  echo "ncu $tag exit $?"
done
ls -la gpurun_out/*.ncu-rep
