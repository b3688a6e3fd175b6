#!/bin/bash
# Development aid: rebuild libwrt_cuda.so with extra nvcc flags on the GPU box and time the two 4K workloads.
# usage: tools/variant_sweep.sh "<nvcc flags>|<ENV=..,ENV=..>" ...
cd "$(dirname "$0")/.."
for spec in "$@"; do
  v="${spec%%|*}"; e="${spec#*|}"; [ "$e" = "$spec" ] && e=""
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -ccbin /usr/bin/g++ $v \
       -o whittedstyle_raytracer_b200/libwrt_cuda.so whittedstyle_raytracer_b200/csrc/cuda/wrt_cuda.cu 2>&1 | grep -E " error"
  if [ -n "$e" ]; then echo "variant [$v | $e]: $(python tools/gpu_sweep.py $e 2>&1 | tail -1)"; else echo "variant [$v]: $(python tools/gpu_sweep.py 2>&1 | head -1)"; fi
done
