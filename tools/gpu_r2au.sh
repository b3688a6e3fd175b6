#!/bin/bash
# round 2, GPU call AU (last of the budget): the tests the hard-shadow association touches; hard-shadow frame times
mkdir -p gpurun_out
timeout 85 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "association or strategy_queries or million or device_built or image_matches or random_scene or 4k_hard or debug_bounds or edge_scenes or reupload" > gpurun_out/r2au_pytest.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r2au_pytest.log
timeout 40 python tools/gpu_variant_time.py bunny_shadow_4k gla_bunny_tex_4k 2>&1 | grep -v "world 8" | tee gpurun_out/r2au_times.log
