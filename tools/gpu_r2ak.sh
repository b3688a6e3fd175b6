#!/bin/bash
# round 2, GPU call AK: parity tests incl. the wavefront closest-hit entry; source-level captures of the final kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2ak_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r2ak_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace_closest -s 16 -c 1 \
   -o gpurun_out/r2ak_trace7 python tools/gpu_one_frame.py water_bunny_tex_soft_4k 2 > gpurun_out/r2ak_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_soft_ -s 12 -c 3 \
   -o gpurun_out/r2ak_softq1 python tools/gpu_one_frame.py water_bunny_tex_soft_4k 2 > gpurun_out/r2ak_ncu2.log 2>&1
ls -la gpurun_out/r2ak*.ncu-rep
