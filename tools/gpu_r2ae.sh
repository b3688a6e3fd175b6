#!/bin/bash
# round 2, GPU call AE: GPU suite on the new default build; variants; launch knobs
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2ae_pytest.log
bash tools/gpu_ab.sh r2ae > /dev/null 2>&1
timeout 600 python tools/gpu_share_sweep.py water_bunny_tex_soft_4k 8 1 > gpurun_out/r2ae_sweep.log 2>&1
