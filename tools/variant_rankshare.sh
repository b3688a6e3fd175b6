#!/bin/bash
# Development aid: rebuild libwrt_cuda.so with extra nvcc flags on the GPU box; time full frame and one rank's 1/8 share.
cd "$(dirname "$0")/.."
for v in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared -ccbin /usr/bin/g++ $v \
       -o whittedstyle_raytracer_b200/libwrt_cuda.so whittedstyle_raytracer_b200/csrc/cuda/wrt_cuda.cu 2>&1 | grep -E " error"
  echo "variant [$v]: $(python tools/gpu_rankshare.py 2>&1 | head -2 | tr '\n' ' ')"
done
