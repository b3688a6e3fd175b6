#!/bin/bash
mkdir -p gpurun_out
for lib in whittedstyle_raytracer_b200/libwrt_cuda.so whittedstyle_raytracer_b200/variants/*.so; do
  WRT_CUDA_LIB=$lib timeout 300 python tools/gpu_variant_time.py f4_spheres_1k_4k 2>&1 | grep -v "world 8" | tee -a gpurun_out/r2ah.log
done
