#!/bin/bash
# round 2, GPU call AW (the last seconds of the budget): hard-shadow association with the second, collecting walk
mkdir -p gpurun_out
timeout 45 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "association or strategy_queries or million or device_built or image_matches or random_scene or debug_bounds" > gpurun_out/r2aw_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2aw_pytest.log
timeout 30 python tools/gpu_variant_time.py bunny_shadow_4k gla_bunny_tex_4k 2>&1 | grep -v "world 8" | tee gpurun_out/r2aw_times.log
