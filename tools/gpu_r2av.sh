#!/bin/bash
# round 2, GPU call AV (last of the budget): the hard-shadow association with the path codes fetched at hit time
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "association or strategy_queries or million or device_built or image_matches or random_scene or 4k_hard or debug_bounds" > gpurun_out/r2av_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2av_pytest.log
timeout 40 python tools/gpu_variant_time.py bunny_shadow_4k gla_bunny_tex_4k 2>&1 | grep -v "world 8" | tee gpurun_out/r2av_times.log
