#!/bin/bash
# round 2, GPU call Y: A/B of the deferred-leaf closest-hit walk (WRT_LEAF_DEFER=1: closest hit, 2: + hard shadows) and the
# two-phase list rays against the in-tree build; launch knobs at 1/8 share
mkdir -p gpurun_out
V=whittedstyle_raytracer_b200/variants
for lib in whittedstyle_raytracer_b200/libwrt_cuda.so $V/libwrt_defer1.so $V/libwrt_defer2.so $V/libwrt_twophase.so $V/libwrt_both.so; do
  [ -f "$lib" ] || continue
  WRT_CUDA_LIB=$lib timeout 300 python tools/gpu_variant_time.py 2>&1 | tee -a gpurun_out/r2y_variants.log
done
timeout 600 python tools/gpu_share_sweep.py water_bunny_tex_soft_4k 8 1 2>&1 | tee gpurun_out/r2y_sweep.log
