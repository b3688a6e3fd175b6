#!/bin/bash
# round 2, GPU call AA: A/B of the simplified deferred-leaf bookkeeping; source-level ncu capture of queue 1's three
# soft-shadow kernels (second frame)
mkdir -p gpurun_out
V=whittedstyle_raytracer_b200/variants
for lib in whittedstyle_raytracer_b200/libwrt_cuda.so $V/libwrt_defer2.so $V/libwrt_defer3.so $V/libwrt_defer4.so $V/libwrt_defer5.so; do
  [ -f "$lib" ] || continue
  WRT_CUDA_LIB=$lib timeout 300 python tools/gpu_variant_time.py 2>&1 | tee -a gpurun_out/r2aa_variants.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_soft_ -s 12 -c 3 \
     -o gpurun_out/r2aa_softq1 python tools/gpu_one_frame.py water_bunny_tex_soft_4k 2 > gpurun_out/r2aa_ncu.log 2>&1
echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
