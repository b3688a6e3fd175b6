#!/bin/bash
# round 2, GPU call AX (the last 13 seconds of the budget, no Python): the shipped k_shadow_hard (every translucent blocker kept on
# the one walk) through the drop-in executable, against PPMs the UNMODIFIED reference executable wrote in the build container
mkdir -p gpurun_out
cd tools/r2ax_case
for n in glass_bunny glass_row hard_bunny; do
  timeout 5 ../../whittedstyle_raytracer_b200/wrt $n.txt > /dev/null 2>&1; rc=$?
  echo "$n rc=$rc differing_bytes=$(cmp -l $n.ppm expected_$n.ppm 2>&1 | wc -l) of $(stat -c %s expected_$n.ppm)"
done 2>&1 | tee ../../gpurun_out/r2ax_cli.log
